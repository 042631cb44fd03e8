"""Output stage (cmtts_b200/output.py): file naming / async writer on the CPU, and on the GPU the whole
`synth_samples` step (mel -> device int16 -> cropped WAV files) against the oracle."""
import argparse
import os

import numpy as np
import pytest
import torch

from cmtts_b200 import output


def test_output_names_follow_the_reference():
    a = argparse.Namespace(mode="single", speaker_id="p225", teacher_forced=False)
    assert output.output_name("utt", a, True, "wav") == "utt_p225.wav"          # utils/tools.py:603-605
    assert output.output_name("utt", a, False, "wav") == "utt.wav"
    b = argparse.Namespace(mode="batch", speaker_id="p225", teacher_forced=True)
    assert output.output_name("utt", b, True, "png") == "utt_teacher_forced.png"


def test_async_writer_round_trip(tmp_path):
    from scipy.io import wavfile
    rng = np.random.default_rng(0)
    wavs = [rng.integers(-32768, 32767, size=n, dtype=np.int16) for n in (1, 256, 7000)]
    with output.AsyncWavWriter(workers=3) as w:
        for i, x in enumerate(wavs):
            w.submit(str(tmp_path / "sub" / f"{i}.wav"), 22050, x)
        x0 = wavs[0].copy()
        wavs[0][:] = 0          # the writer owns a copy: later mutation by the caller must not leak into the file
    for i, x in enumerate([x0] + wavs[1:]):
        rate, got = wavfile.read(str(tmp_path / "sub" / f"{i}.wav"))
        assert rate == 22050 and got.dtype == np.int16 and np.array_equal(got, x)
    with pytest.raises(TypeError):
        with output.AsyncWavWriter() as w:
            w.submit(str(tmp_path / "f.wav"), 22050, np.zeros(4, dtype=np.float32))


@pytest.mark.gpu
def test_synth_samples_writes_cropped_int16_wavs(tmp_path):
    from scipy.io import wavfile

    from cmtts_b200 import synthetic
    from cmtts_b200.config import ModelSpec
    from cmtts_b200.synthesize import Pipeline
    from gpu_util import DEV, Replay, draw_noise
    from oracle import cmtts_oracle as O

    spec = ModelSpec.preset("VCTK")
    sd = synthetic.make_acoustic_state_dict(spec, seed=2)
    ck = synthetic.make_hifigan_checkpoint(spec.hifigan, seed=7)
    pipe = Pipeline(spec, sd, ck["generator"], DEV)
    batch = synthetic.make_batch(spec, 3, 6, 14, seed=8)
    W, Wf = O.Weights(sd), O.Weights(synthetic.fold_weight_norm(ck["generator"]))
    with torch.no_grad():
        pre = O.dpen(W, spec, **batch)
    L = pre["cond"].shape[1]
    noise = draw_noise(5, (3, 1, L, 80), 3)
    it = iter(noise)
    with torch.no_grad():
        _, _, i16, _ = O.synthesize(W, Wf, spec, batch, 2, lambda s: next(it))
    mel, out = pipe.acoustic(batch["texts"], batch["src_lens"], batch["spker_embeds"], 2, Replay(noise))
    preds = [None] * 12
    preds[0], preds[10], preds[11] = mel, batch["src_lens"], out["mel_lens"]
    args = argparse.Namespace(mode="batch", speaker_id="p225", teacher_forced=False, restore_step=1234, model="naive")
    pre_cfg = {"preprocessing": {"stft": {"hop_length": spec.hop_length},
                                 "audio": {"max_wav_value": spec.max_wav_value, "sampling_rate": spec.sampling_rate}}}
    names = ["a", "b", "c"]
    with output.AsyncWavWriter(2) as w:
        paths = output.synth_samples(args, [names], preds, pipe.vocoder, {"multi_speaker": True}, pre_cfg, str(tmp_path), writer=w)
    assert [os.path.relpath(p, tmp_path) for p in paths] == [os.path.join("1234", n + ".wav") for n in names]
    for p, ref, n in zip(paths, i16, out["mel_lens"].cpu().tolist()):
        rate, got = wavfile.read(p)
        assert rate == spec.sampling_rate and got.dtype == np.int16
        assert got.shape == ref.shape == (n * spec.hop_length,)                       # utils/model.py:201-203
        assert np.abs(got.astype(np.int32) - ref.astype(np.int32)).max() <= 132       # fp16 vocoder: 4e-3 * 32768
