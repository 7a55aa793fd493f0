"""Host-side mirror of the reference model object.

``Tacotron(hparams).initialize(inputs, input_lengths, mel_targets,
linear_targets, identities, id_num)`` keeps the contract of reference
``models/tacotron.py:13-124``: same signature, same attributes set
(``mel_outputs``, ``linear_outputs``, ``alignments`` and the echoed inputs,
``models/tacotron.py:106-113``), same branching (``is_training`` keyed off
``linear_targets``, ``:36``; multi-speaker iff ``identities is not None and
id_num > 1``, ``:48``).  TF's deferred graph execution becomes eager: the call
runs the forward on the GPU through ``libtaco_b200.so`` and stores torch CUDA
tensors.

Weights: ``load_weights`` takes a dict keyed by TF checkpoint variable names;
without it ``initialize`` falls back to the reference's initializers with a
fixed seed (there is no network / checkpoint in this environment).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _abi
from .engine import Engine
from .hparams import HParams
from .weights import random_init


def log(msg: str) -> None:
    """Stand-in for ``util.infolog.log`` (reference util/infolog.py:17)."""
    print(msg)


class Tacotron:
    def __init__(self, hparams: HParams, device=None, verbose: bool = True):
        self._hparams = hparams
        self._device = device
        self._verbose = verbose
        self._weights: Optional[Dict[str, np.ndarray]] = None
        self._engine: Optional[Engine] = None
        self._engine_key = None
        self.random_seed = 1234

    # -- weights ----------------------------------------------------------
    def load_weights(self, weights: Dict[str, np.ndarray]) -> None:
        """Variables keyed by TF names (``model/inference/...`` or short)."""
        self._weights = dict(weights)
        self._engine = None

    def _get_engine(self, id_num_eff: int) -> Engine:
        hp = self._hparams
        key = (id_num_eff, hp.outputs_per_step, hp.max_iters, hp.num_mels, hp.num_freq)
        if self._engine is not None and self._engine_key == key:
            return self._engine
        if self._engine is not None:
            self._engine.close()
        eng = Engine(hp, id_num_eff, self._device)
        weights = self._weights
        if weights is None:
            if self._verbose:
                log("No checkpoint given: random-init weights (reference initializers, seed %d)" % self.random_seed)
            weights = random_init(hp, id_num_eff, seed=self.random_seed)
        eng.load_weights(weights)
        self._engine, self._engine_key = eng, key
        return eng

    @property
    def engine(self) -> Optional[Engine]:
        return self._engine

    # -- the reference entry point -----------------------------------------
    def initialize(self, inputs, input_lengths, mel_targets=None, linear_targets=None,
                   identities=None, id_num=0, teacher_force: Optional[bool] = None):
        """Runs the forward path and sets ``mel_outputs [N,T_out,M]``,
        ``linear_outputs [N,T_out,F]``, ``alignments [N,T_in,steps]``.

        ``teacher_force`` is an extension (reference cannot express it): with
        ``linear_targets is None`` and ``teacher_force=True`` the decoder is fed
        ``mel_targets`` but batch norm keeps its moving statistics.
        """
        hp = self._hparams
        is_training = linear_targets is not None                      # tacotron.py:36
        multi = identities is not None and id_num > 1                 # tacotron.py:48
        if is_training and mel_targets is None:
            raise ValueError("linear_targets given without mel_targets")
        tf_mode = is_training if teacher_force is None else bool(teacher_force)
        if tf_mode and mel_targets is None:
            raise ValueError("teacher forcing needs mel_targets")
        eng = self._get_engine(int(id_num) if multi else 0)
        bn_mode = _abi.BN_BATCH if is_training else _abi.BN_MOVING    # modules.py:101
        if self._verbose:
            log("multi-speaker" if multi else "single speaker")
            if is_training:
                print("training")
        mel, lin, al, steps = eng.forward(inputs, input_lengths, identities if multi else None,
                                          mel_targets if tf_mode else None, tf_mode, bn_mode)
        eng.check_ids()
        self.inputs = inputs
        self.input_lengths = input_lengths
        self.mel_outputs = mel
        self.linear_outputs = lin
        self.alignments = al
        self.identities = identities
        self.mel_targets = mel_targets
        self.linear_targets = linear_targets
        self.steps = steps
        if self._verbose:   # same block as tacotron.py:114-124
            emb = hp.embedding_text_channels + (hp.embedding_id_channels if multi else 0)
            log("Initialized Tacotron model. Dimensions: ")
            log("embedding:                 %d" % emb)
            log("prenet out:                %d" % 128)
            log("encoder out:               %d" % 256)
            log("attention out:             %d" % 256)
            log("concat attn & out:         %d" % 512)
            log("decoder cell out:          %d" % 256)
            log("decoder out (%d frames):   %d" % (hp.outputs_per_step, hp.num_mels * hp.outputs_per_step))
            log("decoder out (1 frame):     %d" % hp.num_mels)
            log("postnet out:               %d" % 256)
            log("linear out:                %d" % hp.num_freq)
        return self


def create_model(name: str, hparams: HParams, **kw) -> Tacotron:
    """reference models/__init__.py:4-8."""
    if name == "tacotron":
        return Tacotron(hparams, **kw)
    raise Exception("Unknown model: " + name)
