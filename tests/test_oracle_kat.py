"""Known-answer tests that pin each TensorFlow-1.4 convention the oracle restates
(SURVEY.md §8c KATs 1-13).  Analytic expectations only: the reference ships no
golden vectors and TF cannot run here, so these are what anchors the oracle.
"""
import numpy as np
import pytest
import torch

from oracle import taco_oracle as O
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init, weight_specs

F64 = torch.float64


def sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def gru_w(d, n, gk=None, gb=None, ck=None, cb=None):
    return O.W({
        "g/gates/kernel": np.zeros((d + n, 2 * n)) if gk is None else gk,
        "g/gates/bias": np.zeros(2 * n) if gb is None else gb,
        "g/candidate/kernel": np.zeros((d + n, n)) if ck is None else ck,
        "g/candidate/bias": np.zeros(n) if cb is None else cb,
    }, F64)


def test_kat01_gru_zero_kernels_bias_one():
    # all-zero kernels, gate bias +1 (TF GRUCell default) => c = tanh(0) = 0, h' = sigmoid(1) * h
    w = gru_w(3, 4, gb=np.ones(8))
    h = torch.tensor([[0.3, -1.0, 2.0, 0.5]], dtype=F64)
    out = O.gru_cell(torch.zeros(1, 3, dtype=F64), h, w, "g")
    assert torch.allclose(out, sig(1.0) * h, atol=1e-12)


def test_kat02_gru_gate_order_r_then_u():
    # first half of the gate columns is r, second half is u.  r -> 0 (bias -40), u = 0.5 (bias 0):
    # h' = 0.5 h + 0.5 tanh(x Wcx)   (a swapped order would give u -> 0, h' = c)
    n, d = 2, 2
    gb = np.array([-40.0, -40.0, 0.0, 0.0])
    ck = np.zeros((d + n, n)); ck[0, 0] = 1.0; ck[1, 1] = 1.0; ck[2, 0] = 5.0; ck[3, 1] = 5.0
    w = gru_w(d, n, gb=gb, ck=ck)
    x = torch.tensor([[0.2, -0.4]], dtype=F64)
    h = torch.tensor([[1.0, 2.0]], dtype=F64)
    out = O.gru_cell(x, h, w, "g")
    exp = 0.5 * h + 0.5 * torch.tanh(x)          # r = 0 removes the h rows of the candidate
    assert torch.allclose(out, exp, atol=1e-9)


def test_kat03_gru_reset_applied_before_candidate_matmul():
    # TF: c = tanh([x, r*h] Wc); cuDNN/PyTorch: c = tanh(x Wx + r * (h Uh)).  Off-diagonal Uc tells them apart.
    n, d = 2, 1
    gb = np.array([-40.0, 40.0, -40.0, -40.0])    # r = (0, 1), u = (0, 0) -> h' = c
    ck = np.zeros((d + n, n)); ck[1, 1] = 1.0     # h_0 -> candidate unit 1
    w = gru_w(d, n, gb=gb, ck=ck)
    h = torch.tensor([[1.0, 0.0]], dtype=F64)
    out = O.gru_cell(torch.zeros(1, 1, dtype=F64), h, w, "g")
    # (r*h) = (0*1, 1*0) = 0 -> c = 0.  The other convention would give c_1 = tanh(r_1 * h_0) = tanh(1).
    assert torch.allclose(out, torch.zeros(1, 2, dtype=F64), atol=1e-12)


def test_kat04_conv_same_even_kernel_pads_left_k_minus_1_over_2():
    k, T = 4, 9
    x = torch.zeros(1, T, 1, dtype=F64); x[0, 4, 0] = 1.0
    w = torch.arange(1, k + 1, dtype=F64).reshape(k, 1, 1)
    y = O.conv1d_same(x, w, torch.zeros(1, dtype=F64))[0, :, 0]
    exp = torch.zeros(T, dtype=F64)
    for j in range(k):
        exp[4 - j + 1] = j + 1                   # y[t] = sum_j x[t + j - 1] w[j]
    assert torch.equal(y, exp)
    # boundaries: output length T and the taps that fall off both ends are dropped
    x = torch.ones(1, 3, 1, dtype=F64)
    y = O.conv1d_same(x, w, torch.zeros(1, dtype=F64))[0, :, 0]
    assert torch.equal(y, torch.tensor([2 + 3 + 4, 1 + 2 + 3, 1 + 2], dtype=F64))


def test_kat05_maxpool_same_right_padded():
    x = torch.tensor([[[1.0], [5.0], [2.0], [3.0]]], dtype=F64)
    assert torch.equal(O.max_pool_same2(x)[0, :, 0], torch.tensor([5.0, 5.0, 3.0, 3.0], dtype=F64))


def test_kat06_batch_norm_modes_and_activation_order():
    C = 2
    w = O.W({"c/conv1d/kernel": np.array([[[1.0, -1.0]]]).reshape(1, 1, 2), "c/conv1d/bias": np.zeros(C),
             "c/batch_normalization/gamma": np.ones(C), "c/batch_normalization/beta": np.array([0.5, 0.5]),
             "c/batch_normalization/moving_mean": np.zeros(C), "c/batch_normalization/moving_variance": np.ones(C)}, F64)
    x = torch.tensor([[[1.0], [3.0]]], dtype=F64)
    y = O.conv1d_block(x, w, "c", "relu", "moving")
    # ReLU BEFORE BN: channel 1 is relu(-x) = 0 -> beta;  fresh BN scales by 1/sqrt(1 + 1e-3)
    s = 1.0 / np.sqrt(1.001)
    assert torch.allclose(y[0, :, 0], torch.tensor([1.0 * s + 0.5, 3.0 * s + 0.5], dtype=F64), atol=1e-12)
    assert torch.allclose(y[0, :, 1], torch.tensor([0.5, 0.5], dtype=F64), atol=1e-12)
    # batch mode: biased variance over (N,T): mean 2, var 1
    yb = O.conv1d_block(x, w, "c", "relu", "batch")
    assert torch.allclose(yb[0, :, 0], torch.tensor([-1.0, 1.0], dtype=F64) / np.sqrt(1.0 + 1e-3) + 0.5, atol=1e-12)


def test_kat07_bigru_masking_and_backward_start():
    hp = HParams()
    w = O.W(random_init(hp, 0, seed=1), F64)
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((2, 6, 128)))
    out = O.bigru(x, np.array([6, 3]), w, "encoder_cbhg")
    assert torch.equal(out[1, 3:], torch.zeros(3, 256, dtype=F64))          # zero output beyond the length
    # the backward pass of the short sample starts at t = len-1 from a zero state
    h0 = torch.zeros(1, 128, dtype=F64)
    first = O.gru_cell(x[1:2, 2], h0, w, "encoder_cbhg/bidirectional_rnn/bw/gru_cell")
    assert torch.allclose(out[1, 2, 128:], first[0], atol=1e-12)
    # and the full-length sample equals an unmasked run
    full = O.bigru(x[:1], None, w, "encoder_cbhg")
    assert torch.allclose(out[0], full[0], atol=1e-12)


def _decoder_weights(hp, seed=2):
    return random_init(hp, 0, seed=seed)


def test_kat08_attention_is_not_length_masked():
    hp = HParams(outputs_per_step=2, max_iters=1)
    wd = _decoder_weights(hp)
    w = O.W(wd, F64)
    rng = np.random.default_rng(3)
    memory = torch.from_numpy(rng.standard_normal((1, 5, 256)))
    memory[0, 3:] = 0.0                                                     # padded encoder rows are zeros
    dec, al, steps = O.decode(memory, w, hp.num_mels, 2, 1)
    a = al[0, :, 0]
    assert float(a[3]) > 0 and float(a[4]) > 0 and abs(float(a.sum()) - 1) < 1e-12
    assert abs(float(a[3]) - float(a[4])) < 1e-15                           # both see keys = 0 -> same score


def test_kat09_decoder_prenet_input_order_frame_then_context():
    hp = HParams(outputs_per_step=1, max_iters=1)
    wd = _decoder_weights(hp)
    dp = ("model/inference/decoder/output_projection_wrapper/multi_rnn_cell/cell_0/output_projection_wrapper/"
          "concat_output_and_attention_wrapper/attention_wrapper/decoder_prenet_wrapper/decoder_prenet/dense_1/kernel")
    assert wd[dp].shape == (80 + 256, 256)
    w = O.W(wd, F64)
    st = O.DecoderState(*(torch.zeros(1, 256, dtype=F64) for _ in range(4)))
    st.ctx = torch.ones(1, 256, dtype=F64)
    memory = torch.zeros(1, 3, 256, dtype=F64)
    keys = torch.zeros(1, 3, 256, dtype=F64)
    out_a, _, _ = O.decoder_step(torch.zeros(1, 80, dtype=F64), st, memory, keys, w)
    wd2 = dict(wd); k = wd[dp].copy(); k[:80] = 0.0; wd2[dp] = k           # frame rows are the FIRST 80 rows
    out_b, _, _ = O.decoder_step(torch.zeros(1, 80, dtype=F64), st, memory, keys, O.W(wd2, F64))
    assert torch.equal(out_a, out_b)                                        # zero frame: frame rows are irrelevant
    k = wd[dp].copy(); k[80:] = 0.0; wd2[dp] = k
    out_c, _, _ = O.decoder_step(torch.zeros(1, 80, dtype=F64), st, memory, keys, O.W(wd2, F64))
    assert not torch.equal(out_a, out_c)                                    # context rows do matter


def test_kat10_stop_on_exact_zero_outputs():
    hp = HParams(outputs_per_step=3, max_iters=7)
    wd = _decoder_weights(hp)
    w_free = O.W(wd, F64)
    memory = torch.from_numpy(np.random.default_rng(1).standard_normal((2, 4, 256)))
    _, _, steps = O.decode(memory, w_free, 80, 3, 7)
    assert steps == 7                                                       # never all-zero: runs max_iters
    wd["model/inference/decoder/output_projection_wrapper/kernel"][:] = 0
    wd["model/inference/decoder/output_projection_wrapper/bias"][:] = 0
    dec, al, steps = O.decode(memory, O.W(wd, F64), 80, 3, 7)
    assert steps == 1 and tuple(dec.shape) == (2, 1, 240) and float(dec.abs().max()) == 0.0


def test_kat11_teacher_forcing_feeds_every_rth_frame():
    hp = HParams(outputs_per_step=3, max_iters=50)
    w = O.W(_decoder_weights(hp), F64)
    rng = np.random.default_rng(5)
    memory = torch.from_numpy(rng.standard_normal((1, 4, 256)))
    tg = torch.from_numpy(rng.uniform(0, 1, (1, 9, 80)))
    dec, al, steps = O.decode(memory, w, 80, 3, 50, tg, True)
    assert steps == 3                                                       # T_out / r steps, not max_iters
    tg2 = tg.clone(); tg2[0, [0, 1, 3, 4, 6, 7, 8]] += 1.0                  # only frames r-1, 2r-1 are ever read
    dec2, _, _ = O.decode(memory, w, 80, 3, 50, tg2, True)
    assert torch.equal(dec, dec2)
    tg3 = tg.clone(); tg3[0, 2] += 1.0                                      # frame r-1 feeds step 1
    dec3, _, _ = O.decode(memory, w, 80, 3, 50, tg3, True)
    assert torch.equal(dec3[:, 0], dec[:, 0]) and not torch.equal(dec3[:, 1], dec[:, 1])


def test_kat12_alignment_layout_and_shapes():
    hp = HParams(outputs_per_step=5, max_iters=4)
    wd = random_init(hp, 3, seed=4)
    ids = np.array([[5, 6, 7, 0, 0], [9, 8, 7, 6, 5]], np.int32)
    out = O.tacotron_forward(wd, hp, ids, np.array([3, 5], np.int32), identities=np.array([0, 2], np.int32), id_num=3)
    assert tuple(out["alignments"].shape) == (2, 5, 4)                      # [N, T_in, steps]
    assert torch.allclose(out["alignments"].sum(dim=1), torch.ones(2, 4), atol=1e-5)
    assert tuple(out["mel_outputs"].shape) == (2, 20, 80) and tuple(out["linear_outputs"].shape) == (2, 20, 1025)


def test_kat13_highway_bias_init_and_blend():
    hp = HParams()
    specs = weight_specs(hp, 0)
    assert specs["encoder_cbhg/highway_1/T/bias"][1] == "const:-1.0"
    w = O.W({"h/H/kernel": np.zeros((128, 128)), "h/H/bias": np.full(128, 2.0),
             "h/T/kernel": np.zeros((128, 128)), "h/T/bias": np.full(128, -1.0)}, F64)
    x = torch.full((1, 1, 128), 0.25, dtype=F64)
    y = O.highwaynet(x, w, "h")
    t = sig(-1.0)
    assert torch.allclose(y, torch.full_like(y, 2.0 * t + 0.25 * (1 - t)), atol=1e-12)


def test_is_training_is_keyed_off_linear_targets():
    # mel_targets alone = free-running inference that ignores the targets (reference tacotron.py:36,86-90)
    hp = HParams(outputs_per_step=2, max_iters=3)
    wd = random_init(hp, 0, seed=6, randomize_bn=True)
    ids = np.array([[4, 9, 11]], np.int32); lens = np.array([3], np.int32)
    a = O.tacotron_forward(wd, hp, ids, lens)
    b = O.tacotron_forward(wd, hp, ids, lens, mel_targets=np.ones((1, 6, 80), np.float32))
    assert torch.equal(a["mel_outputs"], b["mel_outputs"])
    c = O.tacotron_forward(wd, hp, ids, lens, mel_targets=np.ones((1, 6, 80), np.float32),
                           linear_targets=np.zeros((1, 6, 1025), np.float32))
    assert not torch.equal(a["mel_outputs"], c["mel_outputs"])


def test_embedding_out_of_range_raises():
    hp = HParams()
    w = O.W({"embedding": np.zeros((7352, 256), np.float32)})
    with pytest.raises(IndexError):
        O.embed(np.array([[7352]]), None, w)
