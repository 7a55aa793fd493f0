"""CPU oracle of the vocoder step: restatement of the reference's
``util/audio.py`` TensorFlow Griffin-Lim in numpy.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tacotron_multispeaker_b200/`` may
import this module; only ``tests/`` and ``__graft_entry__.smoke()`` do.

PARITY UNPINNED for the same reason as ``taco_oracle.py``: the arithmetic is
``tf.contrib.signal.stft`` / ``inverse_stft`` of TensorFlow 1.3/1.4 (reference
``util/audio.py:105-112``), which cannot be installed here, and the reference
holds no audio fixtures.  The TF 1.4 conventions restated (``tensorflow/contrib/
signal/python/ops/spectral_ops.py``, ``window_ops.py``, ``reconstruction_ops.py``):

* ``stft(signals, frame_length, frame_step, fft_length, pad_end=False)``:
  ``frames = 1 + (len - frame_length) // frame_step`` frames, multiplied by
  ``hann_window(frame_length, periodic=True)`` = ``0.5 - 0.5 cos(2 pi n / N)``,
  then ``rfft`` of ``fft_length`` points (frame zero-padded at its END);
* ``inverse_stft(stfts, frame_length, frame_step, fft_length)``: ``irfft`` of
  ``fft_length`` points (1/n scaling, imaginary parts of DC and Nyquist
  ignored), cut to the first ``frame_length`` samples, multiplied by the SAME
  periodic Hann window (the default ``window_fn``; no window-sum
  normalisation), then ``overlap_and_add(frames, frame_step)``;
* Griffin-Lim (``util/audio.py:78-91``): zero-phase start, ``iters`` times
  ``est = stft(y); y = istft(S * est / max(1e-8, |est|))``;
* ``inv_spectrogram_tensorflow`` (``:39-46``) does not undo the pre-emphasis;
  ``Synthesizer.synthesize`` applies ``inv_preemphasis`` (``:23-24``,
  ``scipy.signal.lfilter([1], [1, -0.97])``) to the fetched waveform
  (``synthesizer.py:50``).

Everything is float64 (tie-breaker precision); the kernels compute in float32.
"""
from __future__ import annotations

import numpy as np


def stft_parameters(hp):
    """util/audio.py:114-118."""
    n_fft = (hp.num_freq - 1) * 2
    hop_length = int(hp.frame_shift_ms / 1000 * hp.sample_rate)
    win_length = int(hp.frame_length_ms / 1000 * hp.sample_rate)
    return n_fft, hop_length, win_length


def hann_periodic(n):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft_tf(y, win, hop, n_fft):
    """tf.contrib.signal.stft(y, win, hop, n_fft, pad_end=False) for one signal -> [frames, n_fft/2+1]."""
    y = np.asarray(y, np.float64)
    frames = 1 + (len(y) - win) // hop if len(y) >= win else 0
    idx = np.arange(win)[None, :] + hop * np.arange(frames)[:, None]
    return np.fft.rfft(y[idx] * hann_periodic(win)[None, :], n_fft, axis=-1)


def istft_tf(S, win, hop, n_fft):
    """tf.contrib.signal.inverse_stft(S, win, hop, n_fft) for one [frames, bins] matrix."""
    r = np.fft.irfft(S, n_fft, axis=-1)[:, :win] * hann_periodic(win)[None, :]
    T = r.shape[0]
    y = np.zeros((T - 1) * hop + win, np.float64)
    for t in range(T):
        y[t * hop:t * hop + win] += r[t]
    return y


def denormalize(S, hp):
    """util/audio.py:147-151."""
    return np.clip(S, 0, 1) * -hp.min_level_db + hp.min_level_db


def db_to_amp(x):
    """util/audio.py:138-142."""
    return np.power(10.0, x * 0.05)


def griffin_lim_tf(S, hp, iters=None):
    """util/audio.py:78-91 on magnitudes S [T, bins] (already raised to hparams.power)."""
    n_fft, hop, win = stft_parameters(hp)
    iters = hp.griffin_lim_iters if iters is None else iters
    S = np.asarray(S, np.float64)
    y = istft_tf(S.astype(np.complex128), win, hop, n_fft)
    for _ in range(iters):
        est = stft_tf(y, win, hop, n_fft)
        angles = est / np.maximum(1e-8, np.abs(est))
        y = istft_tf(S * angles, win, hop, n_fft)
    return y


def inv_spectrogram_tensorflow(spectrogram, hp, iters=None):
    """util/audio.py:39-46: normalised linear spectrogram [T, num_freq] -> waveform (pre-emphasis NOT undone)."""
    S = db_to_amp(denormalize(np.asarray(spectrogram, np.float64), hp) + hp.ref_level_db)
    return griffin_lim_tf(np.power(S, hp.power), hp, iters)


def inv_preemphasis(x, hp):
    """util/audio.py:23-24: lfilter([1], [1, -preemphasis], x)."""
    a = hp.preemphasis
    y = np.empty(len(x), np.float64)
    z = 0.0
    for i, v in enumerate(np.asarray(x, np.float64)):
        z = v + a * z
        y[i] = z
    return y


def synthesize_wav(spectrogram, hp, iters=None):
    """The waveform Synthesizer.synthesize hands to save_wav (synthesizer.py:47-50)."""
    return inv_preemphasis(inv_spectrogram_tensorflow(spectrogram, hp, iters), hp)


def save_wav_int16(wav):
    """util/audio.py:14-16: the int16 samples librosa.output.write_wav receives."""
    wav = np.asarray(wav, np.float64) * (32767 / max(0.01, np.max(np.abs(wav))))
    return wav.astype(np.int16)
