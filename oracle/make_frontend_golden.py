"""Golden vectors for the text front-end, generated from the UNMODIFIED reference (test infrastructure).

    python -m oracle.make_frontend_golden        # build container only (/root/reference mounted)

Writes tests/golden/frontend.json: the reference's symbol table (digest + size), id sequences of
`text_to_sequence` (text/__init__.py:15-44) for phoneme strings / raw text, and `preprocess_english`
(synthesize.py:168-192) outputs for sentences whose words are all in the shipped lexicon (so that the
neural g2p model, absent here, is never consulted).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

SENTENCES = [
    "hello world",
    "The quick brown fox jumps over the lazy dog.",
    "Printing, in the only sense with which we are at present concerned, differs from most arts!",
    "well - is it so? yes; it is",
]
PHONEME_TEXTS = [
    "{HH AH0 L OW1 sp W ER1 L D}",
    "{P R IH1 N T IH0 NG sp IH0 N DH AH0 OW1 N L IY0 S EH1 N S}",
    "Turn left on {HH AW1 S T AH0 N} Street.",
    "{sil AA1 spn ZH}",
    "{zh ang1 rr uei5}",
    "plain text, no braces",
    "{NOT_A_PHONE AH0}",
    "",
]


def main():
    # `unidecode` is absent from this image; on the ASCII fixtures used here it is the identity, so a stand-in
    # module with exactly that behaviour (and a loud failure otherwise) is registered before the shim's sink stub
    import types
    if "unidecode" not in sys.modules:
        try:
            import unidecode  # noqa: F401
        except ImportError:
            m = types.ModuleType("unidecode")

            def _ascii_only(t):
                assert t.isascii(), "fixture text must be ASCII (unidecode is not installed)"
                return t
            m.unidecode = _ascii_only
            sys.modules["unidecode"] = m
    ref_shim.install()
    import re
    from string import punctuation

    import numpy as np
    import text as ref_text
    from text.symbols import symbols as ref_symbols

    pre, _, _ = ref_shim.load_configs("LJSpeech")
    lex_path = os.path.join(ref_shim.REFERENCE_ROOT, pre["path"]["lexicon_path"])
    cleaners = pre["preprocessing"]["text"]["text_cleaners"]

    # synthesize.py cannot be imported here (argparse + absent g2p_en at module scope is fine, but it also pulls
    # the whole training stack); its two functions are driven through their own source, unmodified, instead
    src = open(os.path.join(ref_shim.REFERENCE_ROOT, "synthesize.py")).read()
    start, end = src.index("def read_lexicon"), src.index("def synthesize_cm")

    class G2pNever:
        def __call__(self, w):
            # what g2p_en returns for separators; real words must be in the lexicon for these fixtures
            assert w.strip() == "" or all(c in ",;.-?!+" for c in w), f"fixture word {w!r} is not in the lexicon"
            return [c for c in w]

    ns = {"re": re, "np": np, "punctuation": punctuation, "G2p": G2pNever, "text_to_sequence": ref_text.text_to_sequence}
    exec(compile(src[start:end], "reference synthesize.py", "exec"), ns)
    pre_abs = {"path": {"lexicon_path": lex_path}, "preprocessing": pre["preprocessing"]}

    import contextlib
    import io
    out = {
        "n_symbols": len(ref_symbols),
        "symbols_sha256": hashlib.sha256("\x00".join(ref_symbols).encode()).hexdigest(),
        "sil_ids": ref_text.sil_phonemes_ids(),
        "cleaners": cleaners,
        "text_to_sequence": [{"text": t, "ids": ref_text.text_to_sequence(t, cleaners)} for t in PHONEME_TEXTS],
        "preprocess_english": [],
    }
    for s in SENTENCES:
        with contextlib.redirect_stdout(io.StringIO()):
            ids = ns["preprocess_english"](s, pre_abs)
        out["preprocess_english"].append({"text": s, "ids": [int(i) for i in ids]})
    out["sequence_to_text"] = [{"ids": e["ids"], "text": ref_text.sequence_to_text(e["ids"])} for e in out["text_to_sequence"][:3]]
    dst = os.path.join(ROOT, "tests", "golden", "frontend.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", dst, "n_symbols", out["n_symbols"])


if __name__ == "__main__":
    main()
