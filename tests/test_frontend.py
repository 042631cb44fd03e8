"""Text front-end (cmtts_b200/frontend.py) against golden vectors generated from the unmodified reference
(oracle/make_frontend_golden.py -> tests/golden/frontend.json), plus batch-collation behaviour."""
import hashlib
import json
import os

import numpy as np
import pytest

from cmtts_b200 import frontend as F
from cmtts_b200.config import ModelSpec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLDEN) as f:
        return json.load(f)


def test_symbol_table_matches_reference(gold):
    assert len(F.symbols) == gold["n_symbols"]
    assert hashlib.sha256("\x00".join(F.symbols).encode()).hexdigest() == gold["symbols_sha256"]
    assert F.sil_phonemes_ids() == gold["sil_ids"]
    # the embedding has one more row than there are symbols (model/modules.py:117)
    assert ModelSpec.preset("LJSpeech").vocab == len(F.symbols) + 1


def test_text_to_sequence_golden(gold):
    for e in gold["text_to_sequence"]:
        assert F.text_to_sequence(e["text"], gold["cleaners"]) == e["ids"], e["text"]
    for e in gold["sequence_to_text"]:
        assert F.sequence_to_text(e["ids"]) == e["text"]


def test_unknown_cleaner_raises():
    with pytest.raises(Exception, match="Unknown cleaner"):
        F.text_to_sequence("abc", ["no_such_cleaner"])


def _lexicon(tmp_path):
    # a slice of the CMU-style lexicon format: WORD  PH PH PH (first occurrence wins, case-insensitive)
    lines = ["HELLO  HH AH0 L OW1", "hello  HH EH1 L OW0", "WORLD\tW ER1 L D", "IS  IH1 Z", "IT  IH1 T", "SO  S OW1",
             "YES  Y EH1 S", "WELL  W EH1 L"]
    p = tmp_path / "lexicon.txt"
    p.write_text("\n".join(lines) + "\n")
    return str(p)


def test_preprocess_english_small_lexicon(tmp_path):
    cfg = {"path": {"lexicon_path": _lexicon(tmp_path)},
           "preprocessing": {"text": {"text_cleaners": ["english_cleaners"], "language": "en"}}}
    lex = F.read_lexicon(cfg["path"]["lexicon_path"])
    assert lex["hello"] == ["HH", "AH0", "L", "OW1"]               # first occurrence wins
    assert F.read_lexicon(cfg["path"]["lexicon_path"]) is lex       # cached, not re-read per sentence
    assert F.english_phonemes("Hello, world!", lex) == "{HH AH0 L OW1 sp W ER1 L D}"
    assert F.english_phonemes("well - is it so? yes; it is", lex) == \
        "{W EH1 L sp IH1 Z IH1 T S OW1 sp Y EH1 S sp IH1 T IH1 Z}"
    ids = F.preprocess_english("hello world", cfg)
    assert ids.tolist() == F.text_to_sequence("{HH AH0 L OW1 W ER1 L D}", ["english_cleaners"])
    if F._default_g2p() is None:      # no grapheme-to-phoneme model installed: loud failure, no silent guess
        with pytest.raises(KeyError, match="zyzzyva"):
            F.preprocess_english("hello zyzzyva", cfg)
    # an explicit g2p callable is honoured for out-of-lexicon tokens (it also sees the separators, like g2p_en)
    ids2 = F.preprocess_english("hello zyzzyva", cfg, g2p=lambda w: ["Z", "IH1", " ", "V", "AH0"] if w.strip() else [])
    assert ids2.tolist() == F.text_to_sequence("{HH AH0 L OW1 Z IH1 V AH0}", ["english_cleaners"])


@pytest.mark.reference
def test_preprocess_english_golden_full_lexicon(gold):
    from oracle import ref_shim
    pre, _, _ = ref_shim.load_configs("LJSpeech")
    cfg = {"path": {"lexicon_path": os.path.join(ref_shim.REFERENCE_ROOT, pre["path"]["lexicon_path"])},
           "preprocessing": pre["preprocessing"]}
    for e in gold["preprocess_english"]:
        assert F.preprocess_english(e["text"], cfg).tolist() == e["ids"], e["text"]


def test_dataset_and_collate(tmp_path):
    pp = tmp_path / "pre"
    (pp / "spker_embed").mkdir(parents=True)
    (pp / "speakers.json").write_text(json.dumps({"p225": 0, "p226": 1}))
    rng = np.random.default_rng(0)
    for s in ("p225", "p226"):
        np.save(pp / "spker_embed" / f"{s}-spker_embed.npy", rng.standard_normal((1, 512)).astype(np.float32))
    src = tmp_path / "val.txt"
    src.write_text("a|p225|{HH AH0 L OW1}|hello\nb|p226|{W ER1 L D sp HH AH0 L OW1}|world hello\nc|p225|{AH0}|a\n")
    pre = {"path": {"preprocessed_path": str(pp)},
           "preprocessing": {"text": {"text_cleaners": ["english_cleaners"], "language": "en"}, "speaker_embedder": "DeepSpeaker"}}
    ds = F.TextDataset(str(src), pre, {"multi_speaker": True})
    assert len(ds) == 3
    ids, raw, speakers, texts, lens, max_len, emb = ds.collate_fn([ds[i] for i in range(3)])
    assert ids == ["a", "b", "c"] and raw == ["hello", "world hello", "a"]
    assert speakers.tolist() == [0, 1, 0] and lens.tolist() == [4, 9, 1] and max_len == 9
    assert texts.shape == (3, 9) and texts.dtype == np.int64 and texts[0, 4:].tolist() == [0] * 5   # zero padding, pad_1D
    assert emb.shape == (3, 512) and emb.dtype == np.float32
    # length bucketing only reorders
    ds2 = F.TextDataset(str(src), pre, {"multi_speaker": True}, sort_by_length=True)
    assert [ds2[i][0] for i in range(3)] == ["c", "a", "b"]
    assert [b[0] for b in ds2.batches(2)] == [["c", "a"], ["b"]]
    # single speaker: no embeddings, speaker id from the map
    ds3 = F.TextDataset(str(src), pre, {"multi_speaker": False})
    assert ds3.collate_fn([ds3[0]])[-1] is None
