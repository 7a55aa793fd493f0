"""Writes tests/golden/audio_gl.npz: a small normalised linear spectrogram and the waveform the float64 audio oracle
(oracle/audio_oracle.py, the restatement of the reference's util/audio.py TF Griffin-Lim + inv_preemphasis) gives for it
after 0 and 3 iterations.  TensorFlow cannot be installed here, so these are outputs of the restatement, not of TF.
Run from the repository root:  python tests/golden/make_golden_audio.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import audio_oracle as A  # noqa: E402
from tacotron_multispeaker_b200.hparams import HParams  # noqa: E402

hp = HParams()
rng = np.random.default_rng(20260101)
x = rng.uniform(-0.1, 1.05, (6, hp.num_freq)).astype(np.float32)
x[:, 1:] = 0.5 * (x[:, 1:] + x[:, :-1])
out = {"spectrogram": x}
for it in (0, 3):
    out["wav_iters%d" % it] = A.synthesize_wav(x, hp, iters=it)
    out["wav_noemph_iters%d" % it] = A.inv_spectrogram_tensorflow(x, hp, iters=it)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "audio_gl.npz"), **out)
print({k: v.shape for k, v in out.items()})
