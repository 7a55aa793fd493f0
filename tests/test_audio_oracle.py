"""CPU tests of the vocoder step: known-answer tests of the audio oracle's TF conventions (reference
util/audio.py:39-46,78-91,105-118), the host-side wav / alignment writers and the eval-loop naming."""
import ctypes as C
import io
import struct
import wave
import zlib

import numpy as np
import pytest

from oracle import audio_oracle as A
from tacotron_multispeaker_b200.hparams import HParams


@pytest.fixture(scope="module")
def hp():
    return HParams()


def test_stft_parameters(hp):
    assert A.stft_parameters(hp) == (2048, 250, 1000)          # util/audio.py:114-118 at hparams.py:12-15


def test_hann_is_periodic():
    w = A.hann_periodic(1000)
    assert w[0] == 0.0 and abs(w[500] - 1.0) < 1e-15 and abs(w[1] - w[999]) < 1e-15   # symmetric about n = N/2


def test_stft_frame_count_and_zero_padding(hp):
    n_fft, hop, win = A.stft_parameters(hp)
    y = np.random.default_rng(0).standard_normal(3 * hop + win + 17)      # pad_end=False: the tail is dropped
    S = A.stft_tf(y, win, hop, n_fft)
    assert S.shape == (4, 1025)
    # DC bin of frame 2 = sum of the windowed frame (zero padding adds nothing)
    assert abs(S[2, 0] - np.sum(y[2 * hop:2 * hop + win] * A.hann_periodic(win))) < 1e-9


def test_stft_matches_scipy(hp):
    # independent implementation: scipy.signal.stft without boundary extension or padding, periodic Hann, divides by sum(w)
    from scipy import signal
    n_fft, hop, win = A.stft_parameters(hp)
    y = np.random.default_rng(6).standard_normal(7 * hop + win)
    _, _, Z = signal.stft(y, window=signal.get_window("hann", win, fftbins=True), nperseg=win, noverlap=win - hop,
                          nfft=n_fft, boundary=None, padded=False)
    S = A.stft_tf(y, win, hop, n_fft)
    assert Z.T.shape == S.shape
    assert np.max(np.abs(Z.T * A.hann_periodic(win).sum() - S)) < 1e-9


def test_golden_audio_fixture_reproduces(hp):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "audio_gl.npz"))
    x = g["spectrogram"]
    for it in (0, 3):
        assert np.max(np.abs(A.synthesize_wav(x, hp, iters=it) - g["wav_iters%d" % it])) < 1e-12
        assert np.max(np.abs(A.inv_spectrogram_tensorflow(x, hp, iters=it) - g["wav_noemph_iters%d" % it])) < 1e-12


def test_istft_of_stft_scales_by_window_square_sum(hp):
    # no window-sum normalisation in inverse_stft: Hann^2 at 75 % overlap sums to 1.5 in the interior
    n_fft, hop, win = A.stft_parameters(hp)
    y = np.random.default_rng(1).standard_normal(9 * hop + win)
    z = A.istft_tf(A.stft_tf(y, win, hop, n_fft), win, hop, n_fft)
    assert z.shape == y.shape
    inner = slice(win, len(y) - win)
    assert np.max(np.abs(z[inner] - 1.5 * y[inner])) < 1e-9


def test_irfft_ignores_imaginary_dc_and_nyquist(hp):
    n_fft, hop, win = A.stft_parameters(hp)
    S = np.zeros((1, 1025), np.complex128)
    S[0, 0] = 3.0 + 5.0j
    S[0, 1024] = 2.0 - 7.0j
    y = A.istft_tf(S, win, hop, n_fft)
    n = np.arange(win)
    assert np.max(np.abs(y - (3.0 + 2.0 * (-1.0) ** n) / n_fft * A.hann_periodic(win))) < 1e-15


def test_denormalise_and_power(hp):
    # x = 1 -> 0 dB + ref 20 dB -> amplitude 10 -> ** 1.5; x <= 0 -> -100 + 20 dB
    x = np.array([[1.0, 0.0, -3.0, 0.5, 7.0]])
    S = A.db_to_amp(A.denormalize(x, hp) + hp.ref_level_db) ** hp.power
    assert np.allclose(S[0], [10 ** 1.5, 1e-4 ** 1.5, 1e-4 ** 1.5, 10 ** (-30 * 0.05 * 1.5), 10 ** 1.5])


def test_griffin_lim_zero_iterations_is_zero_phase_istft(hp):
    n_fft, hop, win = A.stft_parameters(hp)
    S = np.abs(np.random.default_rng(2).standard_normal((5, 1025)))
    assert np.array_equal(A.griffin_lim_tf(S, hp, iters=0), A.istft_tf(S.astype(np.complex128), win, hop, n_fft))


def test_griffin_lim_reduces_spectral_inconsistency(hp):
    n_fft, hop, win = A.stft_parameters(hp)
    rng = np.random.default_rng(3)
    sig = np.sin(2 * np.pi * 440 / hp.sample_rate * np.arange(11 * hop + win)) + 0.1 * rng.standard_normal(11 * hop + win)
    S = np.abs(A.stft_tf(sig, win, hop, n_fft))

    def err(y):
        return np.linalg.norm(np.abs(A.stft_tf(y, win, hop, n_fft)) / 1.5 - S) / np.linalg.norm(S)
    assert err(A.griffin_lim_tf(S, hp, iters=30)) < err(A.griffin_lim_tf(S, hp, iters=0))


def test_inv_preemphasis_matches_scipy_lfilter(hp):
    from scipy import signal
    x = np.random.default_rng(4).standard_normal(5000)
    assert np.max(np.abs(A.inv_preemphasis(x, hp) - signal.lfilter([1], [1, -hp.preemphasis], x))) < 1e-10


def test_save_wav_peak_normalises_and_truncates(tmp_path, hp):
    from tacotron_multispeaker_b200 import audio
    wav = np.array([0.0, 0.5, -1.0, 0.25, 0.99999], np.float32)
    pcm = audio.wav_to_int16(wav)
    assert pcm.dtype == np.int16 and list(pcm) == list(A.save_wav_int16(wav))
    assert pcm[2] == -32767 and pcm[1] == 16383                 # truncation toward zero, util/audio.py:15-16
    assert list(audio.wav_to_int16(np.array([0.001, -0.002], np.float32))) == [3276, -6553]   # the 0.01 floor
    p = tmp_path / "a.wav"
    audio.save_wav(wav, str(p), hp.sample_rate)
    with wave.open(str(p), "rb") as f:
        assert (f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()) == (1, 2, 20000, 5)
        assert np.array_equal(np.frombuffer(f.readframes(5), "<i2"), pcm)
    buf = io.BytesIO()
    audio.save_wav(wav, buf, hp.sample_rate)
    assert buf.getvalue() == p.read_bytes()


def test_find_endpoint(hp):
    from tacotron_multispeaker_b200 import audio
    sr = 100
    wav = np.concatenate([np.ones(300), np.zeros(600)])
    # windows of 200 samples hopping 50: the first all-quiet window starts at 300 -> returns 300 + 50
    assert audio.find_endpoint(wav, sr, min_silence_sec=2) == 350
    assert audio.find_endpoint(np.ones(900), sr, min_silence_sec=2) == 900


def test_plot_alignment_writes_a_valid_png(tmp_path):
    from tacotron_multispeaker_b200 import plot
    a = np.random.default_rng(5).random((13, 40))
    p = tmp_path / "a.png"
    plot.plot_alignment(a, str(p), info="step 1")
    data = p.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, {}
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xFFFFFFFF
        chunks[tag] = body
        pos += 12 + n
    w, h, depth, ctype = struct.unpack(">IIBB", chunks[b"IHDR"][:10])
    assert depth == 8 and ctype == 2 and w % 40 == 0 and h % 13 == 0
    raw = np.frombuffer(zlib.decompress(chunks[b"IDAT"]), np.uint8).reshape(h, 1 + 3 * w)
    assert not raw[:, 0].any()
    # origin='lower': the image's last row is alignment row 0; its brightest cell maps to the brightest colour
    j = int(np.argmax(a[0]))
    sx = w // 40
    row = raw[h - 1, 1:].reshape(w, 3).astype(int)
    assert row[j * sx].sum() == row.sum(axis=1).max()
    with pytest.raises(ValueError):
        plot.plot_alignment(np.zeros((0, 3)), str(p))


def test_eval_output_paths():
    from tacotron_multispeaker_b200.eval import get_output_base_path
    assert get_output_base_path("/x/logs-a/model.ckpt-1000") == "/x/logs-a/eval/eval-1000"   # eval.py:31-36
    assert get_output_base_path("/x/logs-a/weights.npz") == "/x/logs-a/eval/eval"


def test_audio_params_struct_and_wav_length():
    from tacotron_multispeaker_b200 import _abi
    from tacotron_multispeaker_b200.build import build_library
    build_library()
    lib = _abi.load()
    ap = _abi.TacoAudioParams(20000, 100, 50.0, 12.5, 0.97, -100.0, 20.0, 1.5)
    assert C.sizeof(ap) == 56
    assert lib.taco_wav_length(C.byref(ap), 1000) == 999 * 250 + 1000     # host arithmetic only, no GPU
    assert lib.taco_wav_length(C.byref(ap), 0) == -1
    ap.sample_rate = 0
    assert lib.taco_wav_length(C.byref(ap), 10) == -1
