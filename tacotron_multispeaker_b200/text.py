"""Chinese/pinyin symbol front-end with the behaviour of the reference's
``text_to_sequence2`` (reference ``text/__init__.py:51-62,106-120``,
``text/symbols.py:20-23``).

The symbol table is ``['_', '~'] + json.load(normal.json)`` (7352 entries for
the reference's ``datasets/normal.json``).  The JSON asset is data of the
reference repository and is not vendored here: point ``load_symbols`` (or the
``TACO_SYMBOLS_JSON`` environment variable) at it; like the reference, the
default lookup is ``./datasets/normal.json`` relative to the working
directory.  Only the table SIZE matters to the GPU path
(``hparams.num_symbols``, reference ``models/tacotron.py:40``).

Reference quirks kept on purpose:
 * later duplicates win in the symbol->id dict (``text/__init__.py:11``);
 * unknown symbols are silently dropped (``:108-113``);
 * text inside ``{...}`` is split on spaces into multi-character symbols
   (``:58,119-120``);
 * the appended EOS id comes from the *English* table, where ``'~'`` is id 1
   (``:61``); ``Synthesizer.synthesize`` strips it again
   (``synthesizer.py:39``).
"""
from __future__ import annotations

import json
import os
import re
from typing import Dict, List, Optional

_pad, _eos = "_", "~"
_curly_re = re.compile(r"(.*?)\{(.+?)\}(.*)")
EOS_ID = 1  # _symbol_to_id['~'] of the English table (symbols = [_pad, _eos] + ...)

symbols2: Optional[List[str]] = None
_symbol_to_id2: Dict[str, int] = {}
_id_to_symbol2: Dict[int, str] = {}


def load_symbols(path_or_list=None) -> List[str]:
    """Install the symbol table from a JSON file (a list of strings) or a list."""
    global symbols2, _symbol_to_id2, _id_to_symbol2
    if path_or_list is None:
        path_or_list = os.environ.get("TACO_SYMBOLS_JSON", "./datasets/normal.json")
    if isinstance(path_or_list, (list, tuple)):
        chars = list(path_or_list)
    else:
        with open(path_or_list, "r", encoding="utf-8") as f:
            chars = json.load(f)
    symbols2 = [_pad, _eos] + chars
    _symbol_to_id2 = {s: i for i, s in enumerate(symbols2)}
    _id_to_symbol2 = {i: s for i, s in enumerate(symbols2)}
    return symbols2


def _require_table():
    if symbols2 is None:
        try:
            load_symbols()
        except OSError as e:
            raise RuntimeError(
                "no symbol table: call tacotron_multispeaker_b200.text.load_symbols(path to the "
                "reference's datasets/normal.json) or set TACO_SYMBOLS_JSON") from e


def _symbols_to_sequence2(symbols) -> List[int]:
    out = []
    for s in symbols:
        i = _symbol_to_id2.get(s)
        if i is not None:
            out.append(i)
    return out


def text_to_sequence2(text: str, cleaner_names=None) -> List[int]:
    """ids of the characters of ``text`` (+ EOS id 1).  ``cleaner_names`` is
    accepted and ignored, as in the reference."""
    _require_table()
    sequence: List[int] = []
    while len(text):
        m = _curly_re.match(text)
        if not m:
            sequence += _symbols_to_sequence2(text)
            break
        sequence += _symbols_to_sequence2(m.group(1))
        sequence += _symbols_to_sequence2(m.group(2).split())
        text = m.group(3)
    sequence.append(EOS_ID)
    return sequence


def sequence_to_text2(sequence) -> str:
    _require_table()
    return "".join(_id_to_symbol2[i] for i in sequence if i in _id_to_symbol2)
