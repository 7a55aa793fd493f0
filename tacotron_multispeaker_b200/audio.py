"""Audio side of the synthesis surface: the parts of the reference's
``util/audio.py`` that ``Synthesizer.synthesize`` uses (``synthesizer.py:27,
47-56``), with the vocoder on the GPU.

* ``inv_spectrogram_tensorflow(engine, linear)`` -- reference
  ``util/audio.py:39-46,78-91`` (denormalise, dB -> amplitude, ``** power``,
  Griffin-Lim with ``tf.contrib.signal`` STFT conventions) as
  ``taco_griffin_lim`` kernels; batched over utterances.  Like the reference's
  it does NOT undo the pre-emphasis.
* ``synthesize_wav(engine, linear)`` -- the same followed by
  ``inv_preemphasis`` (``:23-24``), fused into the same call on the device.
* ``save_wav`` (``:14-16``), ``find_endpoint`` (``:55-63``),
  ``_stft_parameters`` (``:114-118``) are host code (numpy / ``wave``): they
  handle a finished waveform once per utterance.

There is no CPU Griffin-Lim here: without the CUDA library these raise.
"""
from __future__ import annotations

import wave

import numpy as np


def _stft_parameters(hp):
    n_fft = (hp.num_freq - 1) * 2
    hop_length = int(hp.frame_shift_ms / 1000 * hp.sample_rate)
    win_length = int(hp.frame_length_ms / 1000 * hp.sample_rate)
    return n_fft, hop_length, win_length


def inv_spectrogram_tensorflow(engine, spectrogram, griffin_lim_iters=None):
    """Normalised linear spectrogram ``[T,num_freq]`` or ``[N,T,num_freq]`` (torch CUDA tensor or array) ->
    waveform(s) on the device, pre-emphasis not undone (the caller applies ``inv_preemphasis``)."""
    return engine.griffin_lim(spectrogram, griffin_lim_iters, inv_preemphasis=False)


def synthesize_wav(engine, spectrogram, griffin_lim_iters=None):
    """``inv_preemphasis(inv_spectrogram_tensorflow(x))`` in one device call (synthesizer.py:47-50)."""
    return engine.griffin_lim(spectrogram, griffin_lim_iters, inv_preemphasis=True)


def wav_to_int16(wav) -> np.ndarray:
    """The samples ``save_wav`` writes: peak-normalised to 32767 with the reference's 0.01 floor, truncated."""
    wav = np.asarray(wav, dtype=np.float32)
    wav = wav * np.float32(32767 / max(0.01, float(np.max(np.abs(wav))) if wav.size else 0.01))
    return wav.astype(np.int16)


def save_wav(wav, path, sample_rate):
    """Mono 16-bit PCM RIFF file, as ``librosa.output.write_wav`` produces for an int16 array."""
    pcm = wav_to_int16(wav)
    with wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(int(sample_rate))
        f.writeframes(pcm.astype("<i2").tobytes())


def _db_to_amp(x):
    return np.power(10.0, x * 0.05)


def find_endpoint(wav, sample_rate, threshold_db=-10, min_silence_sec=2):
    wav = np.asarray(wav)
    window_length = int(sample_rate * min_silence_sec)
    hop_length = int(window_length / 4)
    threshold = _db_to_amp(threshold_db)
    for x in range(hop_length, len(wav) - window_length, hop_length):
        if np.max(wav[x:x + window_length]) < threshold:
            return x + hop_length
    return len(wav)
