"""B200-native Tacotron synthesis forward path (drop-in for the hot path of
Jim-Song/tacotron_multispeaker: ``Tacotron.initialize`` / ``Synthesizer``).

Importing the package does not import torch or load the CUDA library; the
first use of :class:`Tacotron` / :class:`Engine` does, and raises if the
library is not built or no sm_100 GPU is present (no CPU fallback).
"""
from .hparams import HParams, hparams, hparams_debug_string  # noqa: F401

__all__ = ["HParams", "hparams", "hparams_debug_string", "Tacotron", "create_model", "Synthesizer",
           "Engine"]


def __getattr__(name):
    if name in ("Tacotron", "create_model"):
        from . import tacotron
        return getattr(tacotron, name)
    if name == "Synthesizer":
        from .synthesizer import Synthesizer
        return Synthesizer
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(name)
