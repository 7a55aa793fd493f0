"""Parity of the GPU Griffin-Lim vocoder (taco_griffin_lim, through the C ABI) against the float64 numpy
restatement of the reference's TF graph (oracle/audio_oracle.py; reference util/audio.py:39-46,78-91).

Tolerances, relative to the waveform's peak: 2e-5 for the zero-phase start and the first iterations (fp32 FFTs
against float64); Griffin-Lim renormalises every bin's phase each iteration, so bins with |est| near zero amplify
rounding differences -- after many iterations the comparison is on the spectral inconsistency the algorithm
minimises, which must match the oracle's within 5 %."""
import io
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import audio_oracle as A  # noqa: E402  (tests may use the oracle)


@pytest.fixture(scope="module")
def eng():
    from tacotron_multispeaker_b200.engine import Engine
    from tacotron_multispeaker_b200.hparams import HParams
    e = Engine(HParams(), id_num=0)       # the vocoder needs no weights
    yield e
    e.close()


def spectrogram(T, seed, lo=-0.1, hi=1.05):
    rng = np.random.default_rng(seed)
    # smooth-ish in time and frequency, leaving [0, 1] at both ends so that the clip is exercised
    x = rng.uniform(lo, hi, (T, 1025)).astype(np.float32)
    x[:, 1:] = 0.5 * (x[:, 1:] + x[:, :-1])
    return x


def rel_err(got, want):
    want = np.asarray(want, np.float64)
    got = got.detach().cpu().numpy().astype(np.float64) if isinstance(got, torch.Tensor) else np.asarray(got, np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    return float(np.max(np.abs(got - want)) / np.max(np.abs(want)))


@pytest.mark.parametrize("T", [1, 2, 7, 33])
def test_zero_phase_start_matches_oracle(eng, T):
    x = spectrogram(T, T)
    got = eng.griffin_lim(x, griffin_lim_iters=0, inv_preemphasis=False)
    assert got.shape == ((T - 1) * 250 + 1000,)
    assert rel_err(got, A.inv_spectrogram_tensorflow(x, eng.hp, iters=0)) < 2e-5


def test_golden_audio_fixture(eng):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "audio_gl.npz"))
    x = g["spectrogram"]
    for it, tol in ((0, 2e-5), (3, 2e-4)):
        assert rel_err(eng.griffin_lim(x, griffin_lim_iters=it), g["wav_iters%d" % it]) < tol
        assert rel_err(eng.griffin_lim(x, griffin_lim_iters=it, inv_preemphasis=False), g["wav_noemph_iters%d" % it]) < tol


@pytest.mark.parametrize("iters", [1, 2, 4])
def test_first_iterations_match_oracle(eng, iters):
    x = spectrogram(24, 100 + iters)
    got = eng.griffin_lim(x, griffin_lim_iters=iters, inv_preemphasis=False)
    assert rel_err(got, A.inv_spectrogram_tensorflow(x, eng.hp, iters=iters)) < 2e-4


@pytest.mark.parametrize("frame_length_ms,frame_shift_ms", [(50.0, 10.0), (49.95, 12.5), (30.0, 12.5), (102.4, 6.25)])
def test_other_frame_geometries(frame_length_ms, frame_shift_ms):
    # generic overlap counts (win 1000 / hop 200), odd window length (999: scalar loads), short and full-length windows
    from tacotron_multispeaker_b200.engine import Engine
    from tacotron_multispeaker_b200.hparams import HParams
    hp = HParams(frame_length_ms=frame_length_ms, frame_shift_ms=frame_shift_ms)
    e = Engine(hp, id_num=0)
    try:
        x = spectrogram(19, 77)
        n_fft, hop, win = A.stft_parameters(hp)
        got = e.griffin_lim(x, griffin_lim_iters=3)
        assert got.shape == (18 * hop + win,)
        assert rel_err(got, A.synthesize_wav(x, hp, iters=3)) < 2e-4
    finally:
        e.close()


def test_inverse_preemphasis_on_device(eng):
    x = spectrogram(50, 7)
    got = eng.griffin_lim(x, griffin_lim_iters=0, inv_preemphasis=True)
    assert rel_err(got, A.synthesize_wav(x, eng.hp, iters=0)) < 2e-5


@pytest.mark.parametrize("T,preemphasis", [(3, 0.97), (200, 0.97), (200, 0.999), (37, 0.5)])
def test_inverse_preemphasis_lengths_and_coefficients(T, preemphasis):
    # segment chaining across the 32 warps: short signals (segments of 32 samples, carries matter everywhere), long ones,
    # and a coefficient close to 1 whose carry never becomes negligible inside a segment
    from tacotron_multispeaker_b200.engine import Engine
    from tacotron_multispeaker_b200.hparams import HParams
    hp = HParams(preemphasis=preemphasis)
    e = Engine(hp, id_num=0)
    try:
        x = spectrogram(T, 300 + T)
        got = e.griffin_lim(x, griffin_lim_iters=1)
        assert rel_err(got, A.synthesize_wav(x, hp, iters=1)) < 1e-4
    finally:
        e.close()


def test_batch_with_stride_equals_single_calls(eng):
    xs = np.stack([spectrogram(12, 20 + i) for i in range(3)])
    big = torch.zeros(3, 20, 1025, device=eng.device)
    big[:, :12] = torch.as_tensor(xs, device=eng.device)
    got = eng.griffin_lim(big[:, :12].contiguous(), griffin_lim_iters=3)
    assert got.shape == (3, 11 * 250 + 1000)
    for i in range(3):
        one = eng.griffin_lim(xs[i], griffin_lim_iters=3)
        assert torch.equal(got[i], one)                       # bit-identical: no cross-utterance arithmetic
        assert rel_err(one, A.synthesize_wav(xs[i], eng.hp, iters=3)) < 2e-4
    # batch stride through the C ABI directly (the post-net writes [N, max_steps*r, F] with only steps*r rows valid)
    import ctypes as C
    ap = eng.audio_params(3)
    out = torch.empty(3, 11 * 250 + 1000, device=eng.device)
    eng._ck(eng.lib.taco_griffin_lim(eng._h, C.byref(ap), C.c_void_p(big.data_ptr()), 3, 12, 20 * 1025,
                                     C.c_void_p(out.data_ptr()), eng.stream))
    assert torch.equal(out, got)


def test_hundred_iterations_reach_the_oracles_consistency(eng):
    hp = eng.hp
    n_fft, hop, win = A.stft_parameters(hp)
    x = spectrogram(40, 11, lo=0.2, hi=0.9)
    S = np.power(A.db_to_amp(A.denormalize(x.astype(np.float64), hp) + hp.ref_level_db), hp.power)

    def inconsistency(y):
        return np.linalg.norm(np.abs(A.stft_tf(y, win, hop, n_fft)) - S) / np.linalg.norm(S)
    got = eng.griffin_lim(x, inv_preemphasis=False).cpu().numpy().astype(np.float64)     # hparams: 100 iterations
    want = A.inv_spectrogram_tensorflow(x, hp)
    start = A.inv_spectrogram_tensorflow(x, hp, iters=0)
    e_got, e_want, e_start = inconsistency(got), inconsistency(want), inconsistency(start)
    assert np.isfinite(got).all()
    assert e_want < e_start
    assert abs(e_got - e_want) < 0.05 * e_want, (e_got, e_want, e_start)


def test_full_size_utterances(eng):
    # config-3 sized utterances (1000 frames): two iterations against the oracle
    xs = np.stack([spectrogram(1000, 50 + i) for i in range(2)])
    got = eng.griffin_lim(xs, griffin_lim_iters=2)
    assert got.shape == (2, 999 * 250 + 1000)
    for i in range(2):
        assert rel_err(got[i], A.synthesize_wav(xs[i], eng.hp, iters=2)) < 2e-4


def test_cuda_graph_replay_is_bit_identical(eng):
    """taco_griffin_lim on a non-default stream with the same buffers: captured into a CUDA graph on the second call and replayed
    afterwards; every call must give the plain launches' waveform bit for bit and count the same number of kernels."""
    dev = torch.device("cuda", 0)
    lin = torch.from_numpy(np.stack([spectrogram(23, 5), spectrogram(23, 6)])).to(dev)
    eng.set_cuda_graphs(False)
    ref = eng.griffin_lim(lin, 4)
    eng.set_cuda_graphs(True)
    st = torch.cuda.Stream(device=dev)
    st.wait_stream(torch.cuda.current_stream())
    counts = []
    with torch.cuda.stream(st):
        out = torch.zeros_like(ref)
        for rep in range(4):
            out.zero_()
            c0 = eng.launch_count()
            eng.griffin_lim(lin, 4, out=out)
            st.synchronize()
            counts.append(eng.launch_count() - c0)
            assert torch.equal(out, ref), "call %d differs" % rep
    assert len(set(counts)) == 1 and counts[0] == 4 + 4 + 1, counts


def test_bad_arguments(eng):
    from tacotron_multispeaker_b200 import _abi
    with pytest.raises(ValueError):
        eng.griffin_lim(np.zeros((4, 513), np.float32))
    import ctypes as C
    ap = eng.audio_params(1)
    ap.frame_length_ms = 200.0                                   # win 4000 > n_fft
    x = torch.zeros(1, 2, 1025, device=eng.device)
    out = torch.zeros(1, 8000, device=eng.device)
    rc = eng.lib.taco_griffin_lim(eng._h, C.byref(ap), C.c_void_p(x.data_ptr()), 1, 2, 0, C.c_void_p(out.data_ptr()), eng.stream)
    assert rc == _abi.TACO_ERR_UNSUPPORTED


def test_synthesizer_writes_wav_and_alignment(tmp_path):
    from tacotron_multispeaker_b200 import text
    from tacotron_multispeaker_b200.hparams import HParams
    from tacotron_multispeaker_b200.synthesizer import Synthesizer
    text.load_symbols([chr(0x4E00 + i) for i in range(7350)])
    hp = HParams(griffin_lim_iters=5)
    syn = Synthesizer(hp).load(None, id_num=4)
    syn.hparams.max_iters = 12
    sentence = "".join(chr(0x4E00 + i) for i in (5, 17, 300, 4000, 22))
    wav_path, png_path = tmp_path / "o.wav", tmp_path / "o.png"
    data = syn.synthesize(sentence, 2, str(wav_path), str(png_path))
    assert data == wav_path.read_bytes() and png_path.read_bytes()[:4] == b"\x89PNG"
    with wave.open(io.BytesIO(data), "rb") as f:
        steps = syn.model.steps
        assert f.getframerate() == 20000 and f.getnframes() == (steps * hp.outputs_per_step - 1) * 250 + 1000
        pcm = np.frombuffer(f.readframes(f.getnframes()), "<i2")
    # (random-init weights give a near-silent spectrogram: save_wav's 0.01 peak floor applies, not 32767)
    # the same waveform from the oracle on the model's own linear output
    lin = syn.model.linear_outputs[0].cpu().numpy()
    want = A.save_wav_int16(A.synthesize_wav(lin, hp, iters=5))
    assert np.max(np.abs(pcm.astype(np.int32) - want.astype(np.int32))) <= 8


def test_eval_loop_writes_the_reference_file_names(tmp_path):
    # eval.py:55-65: <i>-identity-<id>-<text>.wav and <i>-identity-<id>.png under <base>/eval/eval[-step]/
    from tacotron_multispeaker_b200 import eval as ev, text
    from tacotron_multispeaker_b200.hparams import hparams
    text.load_symbols([chr(0x4E00 + i) for i in range(7350)])
    saved = hparams.copy()
    sentences = tmp_path / "s.txt"
    a, b = "".join(chr(0x4E00 + i) for i in (1, 2, 3)), "".join(chr(0x4E00 + i) for i in (40, 50, 60, 70))
    sentences.write_text(a + "，abc\n" + b + "\n", encoding="utf-8")
    try:
        written = ev.main(["--id_num", "3", "--identity", "1", "--base_dir", str(tmp_path), "--sentences_file", str(sentences),
                           "--hparams", "griffin_lim_iters=2"])
    finally:
        for f in vars(saved):
            setattr(hparams, f, getattr(saved, f))
    base = tmp_path / "eval" / "eval"
    assert [str(p) for p in (base / ("0-identity-1-%s.wav" % a), base / ("1-identity-1-%s.wav" % b))] == written
    for i in range(2):
        assert (base / ("%d-identity-1.png" % i)).read_bytes()[:4] == b"\x89PNG"
    with wave.open(written[0], "rb") as f:
        assert f.getframerate() == 20000 and f.getnframes() == 399 * 250 + 1000      # Synthesizer: max_iters 400, r = 1
