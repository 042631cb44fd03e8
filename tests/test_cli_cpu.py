"""The command line of cmtts_b200.synthesize (the reference's `python synthesize.py ...`, synthesize.py:227-397):
flags, argument checks, batch assembly for both modes, and the loud failure without a CUDA device."""
import json

import numpy as np
import pytest
import torch
import yaml

from cmtts_b200 import _lib, frontend as F, synthesize as S


def _workspace(tmp_path, multi_speaker=False):
    pp = tmp_path / "pre"
    (pp / "spker_embed").mkdir(parents=True)
    (pp / "speakers.json").write_text(json.dumps({"LJSpeech": 0, "p225": 0, "p226": 1}))
    np.save(pp / "spker_embed" / "p226-spker_embed.npy", np.ones((1, 512), np.float32))
    lex = tmp_path / "lexicon.txt"
    lex.write_text("HELLO  HH AH0 L OW1\nWORLD  W ER1 L D\n")
    cfg = tmp_path / "config" / "Toy"
    cfg.mkdir(parents=True)
    pre = {"dataset": "Toy", "path": {"preprocessed_path": str(pp), "lexicon_path": str(lex)},
           "preprocessing": {"text": {"text_cleaners": ["english_cleaners"], "language": "en"},
                             "speaker_embedder": "DeepSpeaker" if multi_speaker else "none", "pitch": {"pitch_type": "cwt"}}}
    model = {"multi_speaker": multi_speaker, "vocoder": {"model": "HiFi-GAN", "speaker": "universal"}}
    train = {"path": {"result_path": str(tmp_path / "result")}}
    for name, d in (("preprocess", pre), ("model", model), ("train", train)):
        (cfg / f"{name}.yaml").write_text(yaml.safe_dump(d))
    src = tmp_path / "val.txt"
    src.write_text("a|p225|{HH AH0 L OW1}|hello\nb|p226|{W ER1 L D}|world\nc|p225|{HH AH0}|he\n")
    return str(tmp_path / "config"), str(src)


def _argv(cfg_dir, *extra):
    return ["--restore_step", "300000", "--dataset", "Toy", "--model_path", "/nonexistent", "--config_dir", cfg_dir, *extra]


def test_flags_and_defaults_follow_the_reference(tmp_path):
    cfg_dir, _ = _workspace(tmp_path)
    a = S.build_arg_parser().parse_args(_argv(cfg_dir, "--mode", "single", "--text", "hello"))
    assert (a.T, a.model, a.speaker_id, a.path_tag, a.batch_size) == (1, "naive", "p225", "", 8)
    assert (a.pitch_control, a.energy_control, a.duration_control) == (1.0, 1.0, 1.0)
    S.check_args(a)
    with pytest.raises(SystemExit):                        # T outside {1, 2, 4}: no sampler plan (synthesize.py:106-146)
        S.build_arg_parser().parse_args(_argv(cfg_dir, "--mode", "single", "--text", "x", "--T", "3"))


@pytest.mark.parametrize("extra,exc", [
    (("--mode", "batch"), ValueError),                                   # batch needs --source
    (("--mode", "batch", "--source", "s", "--text", "t"), ValueError),
    (("--mode", "batch", "--teacher_forced"), NotImplementedError),
    (("--mode", "single"), ValueError),                                  # single needs --text
    (("--mode", "single", "--text", "t", "--source", "s"), ValueError),
])
def test_argument_checks(tmp_path, extra, exc):
    cfg_dir, _ = _workspace(tmp_path)
    with pytest.raises(exc):
        S.check_args(S.build_arg_parser().parse_args(_argv(cfg_dir, *extra)))


def test_single_and_batch_mode_batches(tmp_path):
    cfg_dir, src = _workspace(tmp_path, multi_speaker=True)
    pre, model, _ = S.get_configs_of("Toy", cfg_dir)
    a = S.build_arg_parser().parse_args(_argv(cfg_dir, "--mode", "single", "--text", "hello world", "--speaker_id", "p226"))
    (ids, raw, speakers, texts, lens, max_len, emb), = S.prepare_batches(a, pre, model)
    assert ids == ["hello world"] and speakers.tolist() == [1] and emb.shape == (1, 512)
    assert texts[0].tolist() == F.text_to_sequence("{HH AH0 L OW1 W ER1 L D}", ["english_cleaners"]) and max_len == lens[0]
    b = S.build_arg_parser().parse_args(_argv(cfg_dir, "--mode", "batch", "--source", src, "--batch_size", "2"))
    batches = S.prepare_batches(b, pre, {"multi_speaker": False})
    assert [x[0] for x in batches] == [["a", "b"], ["c"]]
    dev = S.to_device(batches[0], "cpu")
    assert dev[3].dtype == torch.int64 and dev[3].shape == (2, 4) and dev[-1] is None


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the failure mode of a host without a GPU")
def test_main_fails_loudly_without_cuda(tmp_path):
    cfg_dir, _ = _workspace(tmp_path)
    with pytest.raises(_lib.CmttsError, match="no CPU fallback"):
        S.main(_argv(cfg_dir, "--mode", "single", "--text", "hello world", "--T", "4"))
