"""Tokeniser of the lexical store (reference rag/retrieval/bm25.py:34-70) and the
EN/IT language tag (rag/utils/lang_detect.py:17-27).

Host string work: the kernels only ever see term ids.
"""
from __future__ import annotations

import re
from typing import FrozenSet, List, Optional

# letter runs, basic latin + latin-1 accented letters (x D7 and x F7 are signs, not letters)
_LETTERS = re.compile("[A-Za-zÀ-ÖØ-öø-ÿ]+")

STOPWORDS_EN: FrozenSet[str] = frozenset("""
a an the and or but if then else for to of in on at by with from as is are was were be been being it its
this that these those i you he she we they them his her their my your our me us not no yes do does did
doing can could should would may might will shall about into over under again further there here when
where why how what which who whom
""".split())

STOPWORDS_IT: FrozenSet[str] = frozenset("""
un uno una le la il lo gli i l e o ma se allora altrimenti per di a da in su con come è era sono siamo
siete fui fu furono essere stato questo questa questi queste quello quella quelli quelle ciò cio io tu
lui lei noi voi loro mio mia tuo tua suo sua nostro vostro non no si sia fare fa fatto posso può puo
puoi possono dovrebbe potrebbe sarà sara sarebbe saremmo sarete siano che perché perche quando dove
cosa quale chi
""".split())


def stopwords_for(lang_hint: Optional[str]) -> FrozenSet[str]:
    """Italian for tags starting with "it"; English for "en*" and for anything unknown."""
    return STOPWORDS_IT if (lang_hint or "").lower().startswith("it") else STOPWORDS_EN


def tokenize(text: str, lang_hint: Optional[str] = None) -> List[str]:
    """Lower-cased letter runs minus stopwords and one-letter tokens; repeats are kept
    (they carry term frequency on the document side and multiplicity on the query side)."""
    stop = stopwords_for(lang_hint)
    out = []
    for run in _LETTERS.findall(text or ""):
        tok = run.lower()
        if len(tok) > 1 and tok not in stop:
            out.append(tok)
    return out


_detect = None


def detect_lang_tag(text: str) -> str:
    """'en' or 'it'.  Uses langdetect (seeded, as the reference does,
    rag/utils/lang_detect.py:10-27) when it is installed.  Without it the reference cannot even
    be imported; here the tag comes from a stopword vote between the two lists (English on ties
    or empty input) and a RuntimeWarning says so once, because the vote can pick another stopword
    list than langdetect would for a text that carries no metadata language."""
    global _detect
    if _detect is None:
        try:
            from langdetect import DetectorFactory, detect  # type: ignore
            DetectorFactory.seed = 42
            _detect = detect
        except Exception:
            _detect = False
            import warnings
            warnings.warn("langdetect is not installed: classmate_rag_b200 decides 'en' / 'it' by a stopword vote, which "
                          "can differ from the reference's langdetect for texts without a metadata language",
                          RuntimeWarning, stacklevel=2)
    if _detect:
        try:
            lang = _detect(text or "")
            return lang if lang in ("en", "it") else "en"
        except Exception:
            return "en"
    words = [w.lower() for w in _LETTERS.findall(text or "")]
    it = sum(w in STOPWORDS_IT and w not in STOPWORDS_EN for w in words)
    en = sum(w in STOPWORDS_EN and w not in STOPWORDS_IT for w in words)
    return "it" if it > en else "en"
