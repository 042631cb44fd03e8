"""Zero-shot path's speaker encoder: HDF5 reader, feature front end, the CUDA ResCNN vs the numpy oracle.

Parity is UNPINNED for this row (no TensorFlow here to run the reference model); the oracle is anchored on the
`model_config` JSON Keras stored in the reference's own checkpoint (test_oracle_graph_matches_checkpoint_config)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cmtts_b200 import speaker_encoder as SE  # noqa: E402
from cmtts_b200.h5lite import H5File  # noqa: E402
from oracle import deepspeaker_oracle as DO  # noqa: E402

CKPT_REL = "deepspeaker/pretrained_models/ResCNN_triplet_training_checkpoint_265.h5"


def ckpt_path():
    for root in (os.path.join(ROOT, "oracle", "_ref"), "/root/reference"):
        p = os.path.join(root, CKPT_REL)
        if os.path.isfile(p):
            return p
    return None


needs_ckpt = pytest.mark.skipif(ckpt_path() is None, reason="reference DeepSpeaker checkpoint not staged")


def voice_like(seconds=2.5, sr=SE.SAMPLE_RATE, seed=0):
    from cmtts_b200.synthetic import make_voice_like
    return make_voice_like(seconds, sr, seed)


# ---------------------------------------------------------------------------------------------------- CPU
@needs_ckpt
def test_h5_reader_reads_reference_checkpoint():
    f = H5File(ckpt_path())
    assert f.keys("/") == ["model_weights", "optimizer_weights"]
    w = SE.read_keras_weights(ckpt_path())
    cin = 1
    for i, name in enumerate(SE.conv_layer_names()):
        ks, co = (5 if i % 7 == 0 else 3), SE.STAGE_FILTERS[i // 7]
        assert w[f"{name}/kernel:0"].shape == (ks, ks, cin, co) and w[f"{name}/kernel:0"].dtype == np.float32
        assert w[f"{name}/bias:0"].shape == (co,)
        for t in ("gamma", "beta", "moving_mean", "moving_variance"):
            assert w[f"{name}_bn/{t}:0"].shape == (co,)
        assert (w[f"{name}_bn/moving_variance:0"] >= 0).all()
        cin = co
    assert w["affine/kernel:0"].shape == (2048, 512) and w["affine/bias:0"].shape == (512,)
    assert sum(a.size for a in w.values()) == 24_185_728          # 28 convs + BN statistics + Dense
    assert all(np.isfinite(a).all() for a in w.values())


@needs_ckpt
def test_oracle_graph_matches_checkpoint_config():
    """the layer graph both the oracle and the CUDA path implement == the model_config Keras wrote into the checkpoint"""
    raw = open(ckpt_path(), "rb").read()
    i = raw.index(b'{"class_name": "Model"')
    cfg, _ = json.JSONDecoder().raw_decode(raw[i:i + 400000].decode("utf-8", "ignore"))
    layers = cfg["config"]["layers"]
    by_name = {l["name"]: l for l in layers}
    assert by_name["input"]["config"]["batch_input_shape"] == [None, SE.NUM_FRAMES, SE.NUM_FBANKS, 1]
    convs = [l for l in layers if l["class_name"] == "Conv2D"]
    assert [l["name"] for l in convs] == SE.conv_layer_names()
    for i, l in enumerate(convs):
        c = l["config"]
        first = i % 7 == 0
        assert c["kernel_size"] == ([5, 5] if first else [3, 3]) and c["strides"] == ([2, 2] if first else [1, 1])
        assert c["padding"] == "same" and c["use_bias"] and c["activation"] == "linear" and c["filters"] == SE.STAGE_FILTERS[i // 7]
        assert c["data_format"] == "channels_last" and c["dilation_rate"] == [1, 1]
    bns = [l for l in layers if l["class_name"] == "BatchNormalization"]
    assert len(bns) == 28 and all(abs(l["config"]["epsilon"] - SE.BN_EPS) < 1e-12 and l["config"]["axis"] == [3] for l in bns)
    assert abs(DO.BN_EPS - SE.BN_EPS) == 0
    assert by_name["affine"]["config"]["units"] == 512 and by_name["affine"]["config"]["activation"] == "linear"
    assert by_name["reshape"]["config"]["target_shape"] == [-1, 2048]
    # the order inside an identity block: conv, BN, clip, conv, BN, clip, add, clip (conv_models.py:83-108)
    names = [l["name"] for l in layers]
    k = names.index("res1_0_branch_2a")
    kinds = [by_name[n]["class_name"] for n in names[k:k + 8]]
    assert kinds == ["Conv2D", "BatchNormalization", "Lambda", "Conv2D", "BatchNormalization", "Lambda", "Add", "Lambda"]
    add = by_name[names[k + 6]]
    assert {n[0] for n in add["inbound_nodes"][0]} == {names[k + 5], names[k - 1]}       # clipped 2b output + the block input
    assert by_name["ln"]["class_name"] == "Lambda" and by_name["average"]["class_name"] == "Lambda"


def test_fbank_front_end():
    sr = SE.SAMPLE_RATE
    assert SE.calculate_nfft(sr, SE.WIN_LENGTH / sr) == 1024
    fb = SE.mel_filterbank(64, 1024, sr)
    assert fb.shape == (64, 513) and (fb >= 0).all() and fb.max() <= 1.0
    assert (fb.sum(axis=1) > 0).all()                              # no empty filter at this resolution
    peaks = fb.argmax(axis=1)
    assert (np.diff(peaks) > 0).all()                              # centre bins strictly increase
    x = voice_like(1.0)
    e = SE.fbank(x, sr, 64, 1024)
    flen, fstep = 551, 221                                          # round-half-up of 551.25 / 220.5 samples
    assert e.shape == (1 + int(np.ceil((x.size - flen) / fstep)), 64) and (e > 0).all()
    # a pure tone lands in the filter whose triangle covers its bin
    tone = np.sin(2 * np.pi * 1000.0 * np.arange(sr) / sr)
    et = SE.fbank(tone, sr, 64, 1024, preemph=0.0)
    k = int(round(1000.0 * 1024 / sr))
    assert fb[et[5].argmax(), k] > 0
    m = SE.read_mfcc(x, sr, SE.WIN_LENGTH)
    assert m.dtype == np.float32 and m.shape[1] == 64
    np.testing.assert_allclose(m.mean(axis=1), 0.0, atol=1e-5)
    np.testing.assert_allclose(m.std(axis=1), 1.0, atol=1e-4)
    # the 95th-percentile trim drops the quiet head and tail
    assert m.shape[0] < e.shape[0]


def test_sample_from_mfcc():
    m = np.arange(200 * 64, dtype=np.float32).reshape(200, 64)
    s = SE.sample_from_mfcc(m, 160, offset=7)
    assert s.shape == (160, 64, 1) and np.array_equal(s[..., 0], m[7:167])
    r = SE.sample_from_mfcc(m, 160)
    assert r.shape == (160, 64, 1) and any(np.array_equal(r[..., 0], m[o:o + 160]) for o in range(41))
    p = SE.sample_from_mfcc(m[:50], 160)
    assert p.shape == (160, 64, 1) and np.array_equal(p[:50, :, 0], m[:50]) and not p[50:].any()


def test_fold_weights_is_the_oracles_batchnorm():
    w = DO.synthetic_keras_weights(seed=3)
    table = SE.fold_weights(w)
    assert len(table) == 3 * 28 + 2
    rng = np.random.default_rng(0)
    for i, name in enumerate(SE.conv_layer_names()):
        co = table[3 * i].shape[-1]
        acc = rng.standard_normal((5, co))                          # conv output WITHOUT bias
        want = DO.batchnorm(acc + w[f"{name}/bias:0"].astype(np.float64), w, f"{name}_bn")
        got = acc * table[3 * i + 1].astype(np.float64) + table[3 * i + 2].astype(np.float64)
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)


def test_oracle_same_padding_rule():
    # TensorFlow SAME: total = max((ceil(n / s) - 1) s + k - n, 0), the odd element after
    assert DO._same_pads(160, 5, 2) == (80, 1, 2) and DO._same_pads(64, 5, 2) == (32, 1, 2)
    assert DO._same_pads(37, 5, 2) == (19, 2, 2) and DO._same_pads(10, 3, 1) == (10, 1, 1)
    x = np.zeros((1, 5, 4, 1)); x[0, 0, 0, 0] = 1.0
    k = np.arange(25, dtype=np.float32).reshape(5, 5, 1, 1)
    y = DO.conv2d_same(x, k, np.zeros(1, np.float32), 2)
    assert y.shape == (1, 3, 2, 1)
    assert y[0, 0, 0, 0] == k[2, 1, 0, 0]                           # H = 5: pads (2, 2); W = 4: pads (1, 2)


# ---------------------------------------------------------------------------------------------------- GPU
def _run_gpu(w, x):
    model = SE.DeepSpeakerModel("cuda:0").set_keras_weights(w)
    out = model.predict_tensor(torch.from_numpy(x))
    torch.cuda.synchronize()
    return out.cpu().numpy().astype(np.float64)


@pytest.mark.gpu
@pytest.mark.parametrize("B,T", [(2, 160), (1, 37), (3, 100)])
def test_rescnn_vs_oracle_synthetic_weights(B, T):
    w = DO.synthetic_keras_weights(seed=1)
    rng = np.random.default_rng(B * 1000 + T)
    x = rng.standard_normal((B, T, 64)).astype(np.float32)
    want = DO.rescnn_forward(x, w)
    got = _run_gpu(w, x)
    assert got.shape == (B, 512)
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)
    err = np.abs(got - want).max()
    print(f"ResCNN synthetic weights B={B} T={T}: embedding max-abs error vs oracle {err:.2e}")
    assert err <= 2e-5                                              # fp32 FFMA vs float64 (embedding entries are O(0.04))
    # utterances do not interact: row b of a batch == the single-utterance run
    one = _run_gpu(w, x[:1])
    assert np.array_equal(one[0], got[0])


@pytest.mark.gpu
@needs_ckpt
def test_rescnn_reference_checkpoint_vs_oracle(tmp_path):
    from scipy.io import wavfile
    w = SE.read_keras_weights(ckpt_path())
    audio = voice_like(2.5)
    mfcc = SE.read_mfcc(audio, SE.SAMPLE_RATE, SE.WIN_LENGTH)
    x = np.stack([SE.sample_from_mfcc(mfcc, SE.NUM_FRAMES, offset=o)[..., 0] for o in (0, 11)])
    want, stages = DO.rescnn_forward(x, w, return_stages=True)
    got = _run_gpu(w, x)
    err = np.abs(got - want).max()
    cos = float((got[0] * want[0]).sum())
    print(f"ResCNN reference checkpoint: embedding max-abs error vs oracle {err:.2e}, cosine {cos:.7f}; "
          f"stage-4 activations up to {max(float(s.max()) for s in stages):.2f} (clip 20)")
    assert err <= 2e-5 and cos > 0.999999
    # through the call the zero-shot scripts make (synthesize_zeroshot_lj.py:93-97), from a WAV file
    p = str(tmp_path / "ref.wav")
    wavfile.write(p, SE.SAMPLE_RATE, (audio * 32767).astype(np.int16))
    emb = SE.get_deep_speaker_emb(filepath=p, batch_size=3, device="cuda:0", ckpt_path=ckpt_path(), offset=0)
    assert emb.shape == (3, 512) and emb.is_cuda and torch.equal(emb[0], emb[2])
    a16 = SE.load_audio(p)
    x16 = SE.sample_from_mfcc(SE.read_mfcc(a16, SE.SAMPLE_RATE, SE.WIN_LENGTH), SE.NUM_FRAMES, offset=0)[None, ..., 0]
    np.testing.assert_allclose(emb[0].cpu().numpy(), DO.rescnn_forward(x16, w)[0], atol=2e-5)
    # the reference-shaped wrappers
    model = SE.build_model(ckpt_path(), "cuda:0")
    e1 = SE.predict_embedding(model, a16, SE.SAMPLE_RATE, SE.WIN_LENGTH, cuda=True, offset=0)
    assert e1.shape == (1, 512) and np.allclose(e1[0], emb[0].cpu().numpy(), atol=1e-6)
    with pytest.raises(Exception):
        SE.predict_embedding(model, a16, cuda=False)
