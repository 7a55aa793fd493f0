"""Hyper-parameters of the synthesis path.

Mirrors the fields of the reference's ``hparams`` singleton that the forward
path reads (reference ``hparams.py:5-53``; read sites ``models/tacotron.py:43,
49,83,88,94,97,100-101``) and its ``parse("k=v,k2=v2")`` override syntax
(reference ``eval.py:81``).  The TF ``HParams`` container itself is not
reproduced: this is a plain dataclass.

Fork defaults are kept (``outputs_per_step=1``, ``max_iters=2000``); the
benchmark overrides them to the upstream ``r=5, max_iters=200`` that
BASELINE.json quotes.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass

# Number of rows of the text embedding table: ``len(symbols2)`` at reference
# ``models/tacotron.py:40`` = ['_', '~'] + datasets/normal.json (7350 entries).
NUM_SYMBOLS2 = 7352


@dataclass
class HParams:
    cleaners: str = "english_cleaners"
    # Audio (reference hparams.py:11-18)
    num_mels: int = 80
    num_freq: int = 1025
    sample_rate: int = 20000
    frame_length_ms: float = 50.0
    frame_shift_ms: float = 12.5
    preemphasis: float = 0.97
    min_level_db: float = -100.0
    ref_level_db: float = 20.0
    # Model (reference hparams.py:22)
    outputs_per_step: int = 1
    # Eval (reference hparams.py:34-36)
    max_iters: int = 2000
    griffin_lim_iters: int = 100
    power: float = 1.5
    # Network (reference hparams.py:39-40)
    embedding_text_channels: int = 256
    embedding_id_channels: int = 64
    # Input
    bucket_len: int = 1
    eos: bool = True
    chinese_symbol: bool = True
    # Not in the reference: size of the text embedding table (len(symbols2)).
    num_symbols: int = NUM_SYMBOLS2

    def parse(self, overrides: str) -> "HParams":
        """Apply ``"name=value,name=value"`` overrides in place (reference
        ``hparams.parse`` semantics: unknown names raise ``ValueError``)."""
        if not overrides:
            return self
        fields = {f.name: f for f in dataclasses.fields(self)}
        for item in overrides.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" not in item:
                raise ValueError("Could not parse hparam override %r" % item)
            name, value = (s.strip() for s in item.split("=", 1))
            if name not in fields:
                raise ValueError("Unknown hyperparameter: %s" % name)
            cur = getattr(self, name)
            if isinstance(cur, bool):
                if value.lower() not in ("true", "false", "0", "1"):
                    raise ValueError("Could not parse bool hparam %s=%s" % (name, value))
                new = value.lower() in ("true", "1")
            elif isinstance(cur, int):
                new = int(value)
            elif isinstance(cur, float):
                new = float(value)
            else:
                new = value
            setattr(self, name, new)
        return self

    def values(self) -> dict:
        return dataclasses.asdict(self)

    def copy(self) -> "HParams":
        return dataclasses.replace(self)


hparams = HParams()


def hparams_debug_string(hp: HParams | None = None) -> str:
    """Same format as reference ``hparams.py:56-59``."""
    values = (hp or hparams).values()
    lines = ["    %s: %s" % (name, values[name]) for name in sorted(values)]
    return "Hyperparameters:\n" + "\n".join(lines)
