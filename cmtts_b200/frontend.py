"""Text front-end and batch collation of the synthesis path — host code, CPU only.

Mirrors, with the same names and argument meaning:
  * the symbol inventory and `text_to_sequence`       (reference text/symbols.py:9-29, text/__init__.py:15-79)
  * `read_lexicon`, `preprocess_english`                (reference synthesize.py:155-192)
  * `TextDataset` + `collate_fn`, `pad_1D`              (reference dataset.py:237-296, utils/tools.py:744-758)
  * the `--mode single` batch                           (reference synthesize.py:373-394)

Differences that do not change results: the lexicon is parsed once and cached (the reference re-reads
the 5.6 MB file for every sentence, synthesize.py:170); batches can be length-bucketed
(`TextDataset(..., sort_by_length=True)`), which only changes the padding — results depend on the padded
lengths exactly as the reference's do.  Out-of-lexicon words need a grapheme-to-phoneme
model (`g2p_en` in the reference, synthesize.py:172-179); it is used when importable or when a callable
is passed, otherwise `preprocess_english` raises `KeyError` naming the word — no silent guess.

The symbol ids are part of the checkpoint contract (`embed_tokens.weight` has len(symbols) + 1 rows,
model/modules.py:117): the inventory below is generated, and pinned against the reference's table by
tests/test_frontend.py (digest + golden id sequences).
"""
from __future__ import annotations

import json
import os
import re
from string import punctuation
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

# ------------------------------------------------------------------------------------------------
# symbols (ids are positions in this list)
# ------------------------------------------------------------------------------------------------
_PAD = "_"
_SPECIAL = "-"
_PUNCTUATION = "!'(),.:;? "
_LETTERS = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"
_SILENCES = ["@sp", "@spn", "@sil"]

# ARPAbet: the CMUdict phone set in sorted order; vowels come bare and with stress 0/1/2
_ARPA_VOWELS = "AA AE AH AO AW AY EH ER EY IH IY OW OY UH UW".split()
_ARPA_CONSONANTS = "B CH D DH F G HH JH K L M N NG P R S SH T TH V W Y Z ZH".split()
# pinyin inventory of the multilingual symbol table: initials, finals x tones 1-5, plus "rr"
_PY_INITIALS = "b c ch d f g h j k l m n p q r s sh t w x y z zh".split()
_PY_FINALS = ("a ai an ang ao e ei en eng er i ia ian iang iao ie ii iii in ing iong iou o ong ou u ua uai uan uang "
              "uei uen uo v van ve vn").split()


def _build_symbols() -> List[str]:
    arpa = sorted(_ARPA_CONSONANTS + [v + s for v in _ARPA_VOWELS for s in ("", "0", "1", "2")])
    pinyin = _PY_INITIALS + [f + str(t) for f in _PY_FINALS for t in range(1, 6)] + ["rr"]
    return ([_PAD] + list(_SPECIAL) + list(_PUNCTUATION) + list(_LETTERS) + ["@" + s for s in arpa]
            + ["@" + s for s in pinyin] + _SILENCES)


symbols: List[str] = _build_symbols()
_symbol_to_id: Dict[str, int] = {s: i for i, s in enumerate(symbols)}
_id_to_symbol: Dict[int, str] = {i: s for i, s in enumerate(symbols)}
valid_arpabet = frozenset(s[1:] for s in symbols if s.startswith("@"))

# ------------------------------------------------------------------------------------------------
# cleaners (reference text/cleaners.py:62-89)
# ------------------------------------------------------------------------------------------------
_whitespace_re = re.compile(r"\s+")
_ABBREVIATIONS = [(re.compile(r"\b%s\." % a, re.IGNORECASE), b) for a, b in [
    ("mrs", "misess"), ("mr", "mister"), ("dr", "doctor"), ("st", "saint"), ("co", "company"), ("jr", "junior"),
    ("maj", "major"), ("gen", "general"), ("drs", "doctors"), ("rev", "reverend"), ("lt", "lieutenant"),
    ("hon", "honorable"), ("sgt", "sergeant"), ("capt", "captain"), ("esq", "esquire"), ("ltd", "limited"),
    ("col", "colonel"), ("ft", "fort")]]


def basic_cleaners(text: str) -> str:
    return _whitespace_re.sub(" ", text.lower())


def transliteration_cleaners(text: str) -> str:
    return basic_cleaners(_to_ascii(text))


def english_cleaners(text: str) -> str:
    """ascii -> lowercase -> numbers -> abbreviations -> whitespace (text/cleaners.py:82-89)."""
    text = _to_ascii(text).lower()
    text = _expand_numbers(text)
    for rx, rep in _ABBREVIATIONS:
        text = rx.sub(rep, text)
    return _whitespace_re.sub(" ", text)


def _to_ascii(text: str) -> str:
    if text.isascii():
        return text          # unidecode is the identity on ASCII
    try:
        from unidecode import unidecode
    except ImportError as e:   # pragma: no cover - depends on the environment
        raise RuntimeError("non-ASCII text needs the `unidecode` package (reference text/cleaners.py:18)") from e
    return unidecode(text)


def _expand_numbers(text: str) -> str:
    if not any(ch.isdigit() for ch in text):
        return text          # normalize_numbers only rewrites digit groups (text/numbers.py:66-73)
    try:
        import inflect  # noqa: F401
    except ImportError as e:   # pragma: no cover - depends on the environment
        raise RuntimeError("digits in raw text need the `inflect` package (reference text/numbers.py:3); "
                           "phoneme input in {curly braces} does not") from e
    return _normalize_numbers_inflect(text)


def _normalize_numbers_inflect(text: str) -> str:   # pragma: no cover - needs inflect
    """text/numbers.py:13-73 restated: commas, pounds, dollars, decimals, ordinals, cardinals (years read in pairs)."""
    import inflect
    eng = inflect.engine()

    def dollars(m):
        parts = m.group(1).split(".")
        if len(parts) > 2:
            return m.group(1) + " dollars"
        d = int(parts[0]) if parts[0] else 0
        c = int(parts[1]) if len(parts) > 1 and parts[1] else 0
        du, cu = ("dollar" if d == 1 else "dollars"), ("cent" if c == 1 else "cents")
        if d and c:
            return f"{d} {du}, {c} {cu}"
        if d:
            return f"{d} {du}"
        if c:
            return f"{c} {cu}"
        return "zero dollars"

    def number(m):
        n = int(m.group(0))
        if 1000 < n < 3000:
            if n == 2000:
                return "two thousand"
            if 2000 < n < 2010:
                return "two thousand " + eng.number_to_words(n % 100)
            if n % 100 == 0:
                return eng.number_to_words(n // 100) + " hundred"
            return eng.number_to_words(n, andword="", zero="oh", group=2).replace(", ", " ")
        return eng.number_to_words(n, andword="")

    text = re.sub(r"([0-9][0-9\,]+[0-9])", lambda m: m.group(1).replace(",", ""), text)
    text = re.sub(r"£([0-9\,]*[0-9]+)", r"\1 pounds", text)
    text = re.sub(r"\$([0-9\.\,]*[0-9]+)", dollars, text)
    text = re.sub(r"([0-9]+\.[0-9]+)", lambda m: m.group(1).replace(".", " point "), text)
    text = re.sub(r"[0-9]+(st|nd|rd|th)", lambda m: eng.number_to_words(m.group(0)), text)
    return re.sub(r"[0-9]+", number, text)


_CLEANERS = {"basic_cleaners": basic_cleaners, "transliteration_cleaners": transliteration_cleaners,
             "english_cleaners": english_cleaners}

# ------------------------------------------------------------------------------------------------
# text -> ids (reference text/__init__.py)
# ------------------------------------------------------------------------------------------------
_curly_re = re.compile(r"(.*?)\{(.+?)\}(.*)")


def _clean_text(text: str, cleaner_names: Sequence[str]) -> str:
    for name in cleaner_names:
        if name not in _CLEANERS:
            raise Exception("Unknown cleaner: %s" % name)     # text/__init__.py:64
        text = _CLEANERS[name](text)
    return text


def _keep(s: str) -> bool:
    return s in _symbol_to_id and s != "_" and s != "~"


def _symbols_to_sequence(syms: Iterable[str]) -> List[int]:
    return [_symbol_to_id[s] for s in syms if _keep(s)]


def text_to_sequence(text: str, cleaner_names: Sequence[str]) -> List[int]:
    """Ids of `text`; ARPAbet in {curly braces} is looked up as "@PHONE", the rest letter by letter after
    cleaning; unknown symbols are dropped silently, like the reference does (text/__init__.py:15-44, :69-79)."""
    seq: List[int] = []
    while len(text):
        m = _curly_re.match(text)
        if not m:
            seq += _symbols_to_sequence(_clean_text(text, cleaner_names))
            break
        seq += _symbols_to_sequence(_clean_text(m.group(1), cleaner_names))
        seq += _symbols_to_sequence("@" + s for s in m.group(2).split())
        text = m.group(3)
    return seq


def sequence_to_text(sequence: Iterable[int]) -> str:
    out = ""
    for i in sequence:
        s = _id_to_symbol.get(int(i))
        if s is None:
            continue
        out += "{%s}" % s[1:] if len(s) > 1 and s[0] == "@" else s
    return out.replace("}{", " ")


def sil_phonemes_ids() -> List[int]:
    return [_symbol_to_id[s] for s in _SILENCES]


# ------------------------------------------------------------------------------------------------
# lexicon + English preprocessing (reference synthesize.py:155-192)
# ------------------------------------------------------------------------------------------------
_LEXICON_CACHE: Dict[Tuple[str, float], Dict[str, List[str]]] = {}


def read_lexicon(lex_path: str) -> Dict[str, List[str]]:
    """word -> phones; first occurrence of a (lower-cased) word wins.  Cached per (path, mtime)."""
    key = (os.path.abspath(lex_path), os.path.getmtime(lex_path))
    lex = _LEXICON_CACHE.get(key)
    if lex is None:
        lex = {}
        with open(lex_path) as f:
            for line in f:
                parts = re.split(r"\s+", line.strip("\n"))
                w = parts[0].lower()
                if w not in lex:
                    lex[w] = parts[1:]
        _LEXICON_CACHE[key] = lex
    return lex


def _default_g2p() -> Optional[Callable[[str], List[str]]]:
    try:
        from g2p_en import G2p   # the reference's out-of-lexicon model (synthesize.py:16, :172)
    except ImportError:
        return None
    return G2p()


def english_phonemes(text: str, lexicon: Dict[str, List[str]], g2p: Optional[Callable[[str], List[str]]] = None) -> str:
    """The "{PH PH sp PH}" phoneme string of a sentence (synthesize.py:168-183)."""
    text = text.rstrip(punctuation)
    phones: List[str] = []
    for w in re.split(r"([,;.\-\?\!\s+])", text):
        if w.lower() in lexicon:
            phones += lexicon[w.lower()]
        else:
            if g2p is None:
                g2p = _default_g2p()
            if g2p is None:
                if w.strip() == "" or all(ch in ",;.-?!+" for ch in w):
                    # what G2p returns for separators: the token itself (punctuation) or nothing (white space)
                    phones += [p for p in w if p != " "]
                    continue
                raise KeyError(f"'{w}' is not in the lexicon and no grapheme-to-phoneme model is available "
                               f"(install g2p_en or pass g2p=)")
            phones += [p for p in g2p(w) if p != " "]
    s = "{" + "}{".join(phones) + "}"
    s = re.sub(r"\{[^\w\s]?\}", "{sp}", s)      # punctuation and empty tokens become short pauses
    return s.replace("}{", " ")


def preprocess_english(text: str, preprocess_config: dict, g2p=None, verbose: bool = False) -> np.ndarray:
    """Sentence -> int64 id array (synthesize.py:168-192)."""
    lexicon = read_lexicon(preprocess_config["path"]["lexicon_path"])
    phones = english_phonemes(text, lexicon, g2p)
    if verbose:
        print("Raw Text Sequence: {}".format(text.rstrip(punctuation)))
        print("Phoneme Sequence: {}".format(phones))
    return np.array(text_to_sequence(phones, preprocess_config["preprocessing"]["text"]["text_cleaners"]))


# ------------------------------------------------------------------------------------------------
# batches (reference dataset.py:237-296, utils/tools.py:744-758, synthesize.py:373-394)
# ------------------------------------------------------------------------------------------------
def pad_1D(inputs: Sequence[np.ndarray], PAD: int = 0) -> np.ndarray:
    max_len = max((len(x) for x in inputs), default=0)
    out = np.full((len(inputs), max_len), PAD, dtype=np.int64)
    for i, x in enumerate(inputs):
        out[i, : len(x)] = x
    return out


class TextDataset:
    """`basename|speaker|{phonemes}|raw text` lines -> items (basename, speaker_id, phone ids, raw_text,
    spker_embed (1,512) | None); `collate_fn` builds the reference's inference 7-tuple
    (ids, raw_texts, speakers, texts, src_lens, max_src_len, spker_embeds)."""

    def __init__(self, filepath: str, preprocess_config: dict, model_config: dict, sort_by_length: bool = False):
        self.cleaners = preprocess_config["preprocessing"]["text"]["text_cleaners"]
        self.preprocessed_path = preprocess_config["path"]["preprocessed_path"]
        self.load_spker_embed = bool(model_config["multi_speaker"]) and \
            preprocess_config["preprocessing"]["speaker_embedder"] != "none"
        self.basename, self.speaker, self.text, self.raw_text = self.process_meta(filepath)
        with open(os.path.join(self.preprocessed_path, "speakers.json")) as f:
            self.speaker_map = json.load(f)
        self._phones = [np.array(text_to_sequence(t, self.cleaners), dtype=np.int64) for t in self.text]
        self.order = list(range(len(self.text)))
        if sort_by_length:      # neighbours in a batch have similar lengths -> less padding
            self.order.sort(key=lambda i: len(self._phones[i]))
        self._embed_cache: Dict[str, np.ndarray] = {}

    def __len__(self) -> int:
        return len(self.text)

    def __getitem__(self, idx: int):
        i = self.order[idx]
        spk = self.speaker[i]
        emb = None
        if self.load_spker_embed:
            emb = self._embed_cache.get(spk)
            if emb is None:
                emb = np.load(os.path.join(self.preprocessed_path, "spker_embed", "{}-spker_embed.npy".format(spk)))
                self._embed_cache[spk] = emb
        return (self.basename[i], self.speaker_map[spk], self._phones[i], self.raw_text[i], emb)

    @staticmethod
    def process_meta(filename: str):
        name, speaker, text, raw = [], [], [], []
        with open(filename, "r", encoding="utf-8") as f:
            for line in f.readlines():
                n, s, t, r = line.strip("\n").split("|")
                name.append(n); speaker.append(s); text.append(t); raw.append(r)
        return name, speaker, text, raw

    def collate_fn(self, data):
        ids = [d[0] for d in data]
        speakers = np.array([d[1] for d in data])
        texts = [d[2] for d in data]
        raw_texts = [d[3] for d in data]
        text_lens = np.array([t.shape[0] for t in texts])
        spker_embeds = np.concatenate([d[4] for d in data], axis=0) if self.load_spker_embed else None
        return ids, raw_texts, speakers, pad_1D(texts), text_lens, max(text_lens), spker_embeds

    def batches(self, batch_size: int = 8):
        """The reference iterates `DataLoader(dataset, batch_size=8, collate_fn=...)` (synthesize.py:366-370)."""
        for s in range(0, len(self), batch_size):
            yield self.collate_fn([self[i] for i in range(s, min(s + batch_size, len(self)))])


def single_batch(text: str, speaker_id: str, preprocess_config: dict, model_config: dict, g2p=None):
    """The `--mode single` batch (synthesize.py:373-394); raw text is cut to 100 characters for the ids only."""
    ids = raw_texts = [text[:100]]
    multi = bool(model_config["multi_speaker"])
    load_embed = multi and preprocess_config["preprocessing"]["speaker_embedder"] != "none"
    pp = preprocess_config["path"]["preprocessed_path"]
    with open(os.path.join(pp, "speakers.json")) as f:
        speaker_map = json.load(f)
    speakers = np.array([speaker_map[speaker_id]]) if multi else np.array([0])
    emb = np.load(os.path.join(pp, "spker_embed", "{}-spker_embed.npy".format(speaker_id))) if load_embed else None
    lang = preprocess_config["preprocessing"]["text"]["language"]
    if lang != "en":
        raise NotImplementedError(lang)     # synthesize.py:389-390
    texts = np.array([preprocess_english(text, preprocess_config, g2p)])
    text_lens = np.array([len(texts[0])])
    return (ids, raw_texts, speakers, texts, text_lens, max(text_lens), emb)
