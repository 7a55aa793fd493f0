"""Serving facade with the reference's ``Synthesizer`` surface.

Mirrors reference ``synthesizer.py:13-58``: ``load(checkpoint_path,
model_name='tacotron')`` builds the model for batch-1 inference with
``max_iters = 400`` (``synthesizer.py:21``) and discovers ``id_num`` from the
shape of the checkpoint variable ``model/inference/embedding_id``
(``synthesizer.py:23-25``; a single-speaker checkpoint raises ``KeyError``
exactly like the reference); ``synthesize(text, identity, path, path_align)``
converts text with ``text_to_sequence2(...)[:-1]`` (``synthesizer.py:39``) and
runs the forward path.

``synthesize`` then runs the reference's Griffin-Lim graph (``synthesizer.py:27``,
``util/audio.py:39-46``) and ``inv_preemphasis`` (``synthesizer.py:50``) on the GPU
(``taco_griffin_lim``), writes the waveform with ``save_wav`` to ``path`` (``./1.wav``
when ``path`` is None, as the reference does) and the alignment image to
``path_align``.  It returns the bytes of the WAV file (the reference returns an
empty ``BytesIO`` it never writes to -- ``synthesizer.py:52-58``).
Checkpoints are TensorFlow V2 bundles (``model.ckpt-N.index`` + ``.data-*``, read by
``tf_checkpoint.py`` without TensorFlow) or ``.npz`` archives keyed by TF variable
names; without one, ``load(None, id_num=...)`` uses random-init weights.
"""
from __future__ import annotations

import io
from typing import Optional

import numpy as np

from . import audio, plot
from .hparams import HParams, hparams as default_hparams
from .tacotron import create_model
from .text import sequence_to_text2, text_to_sequence2
from .tf_checkpoint import load_weights
from .weights import PREFIX, random_init


class Synthesizer:
    def __init__(self, hparams: Optional[HParams] = None, device=None, verbose: bool = False):
        self.hparams = (hparams or default_hparams).copy()
        self._device = device
        self._verbose = verbose
        self.model = None
        self.id_num = 0

    def load(self, checkpoint_path, model_name="tacotron", id_num: Optional[int] = None, seed: int = 1234):
        print("Constructing model: %s" % model_name)
        hp = self.hparams
        hp.chinese_symbol = True          # synthesizer.py:20
        hp.max_iters = 400                # synthesizer.py:21
        self.model = create_model(model_name, hp, device=self._device, verbose=self._verbose)
        if checkpoint_path is None:
            if id_num is None:
                raise ValueError("load(None) needs id_num for the random-init fallback")
            weights = random_init(hp, id_num, seed=seed)
        else:
            print("Loading checkpoint: %s" % checkpoint_path)
            # a TF V2 checkpoint prefix ("model.ckpt-1000"), a log directory with a `checkpoint` state file, or .npz
            weights = load_weights(checkpoint_path)
        var_to_shape_map = {k: tuple(v.shape) for k, v in weights.items()}
        self.id_num = var_to_shape_map[PREFIX + "embedding_id"][0]   # KeyError if single-speaker: synthesizer.py:25
        self.model.load_weights(weights)
        return self

    def synthesize_sequence(self, seq, identity: int):
        """Batch-1 forward on an id sequence.  Returns (linear [T_out,F], alignment [T_in,steps])
        as numpy arrays (the two fetches of synthesizer.py:47 before Griffin-Lim)."""
        seq = np.asarray(seq, dtype=np.int32)[None, :]
        lengths = np.asarray([seq.shape[1]], dtype=np.int32)
        ident = np.asarray([identity], dtype=np.int32)
        self.model.initialize(seq, lengths, identities=ident, id_num=self.id_num)
        return (self.model.linear_outputs[0].cpu().numpy(), self.model.alignments[0].cpu().numpy())

    def synthesize_sequence_wav(self, seq, identity: int):
        """Batch-1 forward + vocoder on an id sequence.  Returns (wav float32 [L], alignment [T_in,steps])."""
        seq = np.asarray(seq, dtype=np.int32)[None, :]
        lengths = np.asarray([seq.shape[1]], dtype=np.int32)
        ident = np.asarray([identity], dtype=np.int32)
        self.model.initialize(seq, lengths, identities=ident, id_num=self.id_num)
        wav = audio.synthesize_wav(self.model.engine, self.model.linear_outputs[0])
        return wav.cpu().numpy(), self.model.alignments[0].cpu().numpy()

    def synthesize(self, text, identity, path=None, path_align=None):
        seq = text_to_sequence2(text, [x.strip() for x in self.hparams.cleaners.split(",")])[:-1]
        print(seq)
        print(sequence_to_text2(seq))
        wav, alignment = self.synthesize_sequence_wav(seq, identity)
        if path_align is not None:
            plot.plot_alignment(alignment, path_align)
        # wav = wav[:audio.find_endpoint(wav, self.hparams.sample_rate)]   (disabled in the reference too)
        out = io.BytesIO()
        audio.save_wav(wav, out, self.hparams.sample_rate)
        audio.save_wav(wav, path if path is not None else "./1.wav", self.hparams.sample_rate)
        return out.getvalue()
