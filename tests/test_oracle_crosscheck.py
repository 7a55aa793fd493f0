"""Independent cross-check of the oracle's layer functions against library implementations of the same maths
(``torch.nn.functional.conv1d / max_pool1d / batch_norm``, ``torch.nn.GRUCell`` with the TF->cuDNN weight mapping
where the two conventions coincide, numpy ``einsum``): a different code path from ``oracle/taco_oracle.py``, so a slip
in the restatement (tap order, padding side, gate order, BN epsilon placement) shows up here even though TensorFlow
itself cannot be run.  What this cannot pin are the TF conventions themselves (DESIGN.md section 2 lists them)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import taco_oracle as O


def _rng(seed):
    return np.random.default_rng(seed)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 8, 15, 16])
def test_conv1d_same_vs_functional_conv1d(k):
    """tf 'same' padding: pad_left = (k-1)//2, pad_right = k-1-pad_left; cross-correlation (no kernel flip)."""
    r = _rng(k)
    N, T, Cin, Cout = 3, 23, 7, 5
    x = torch.from_numpy(r.standard_normal((N, T, Cin))).double()
    w = torch.from_numpy(r.standard_normal((k, Cin, Cout))).double()
    b = torch.from_numpy(r.standard_normal(Cout)).double()
    got = O.conv1d_same(x, w, b)
    pl, pr = (k - 1) // 2, k - 1 - (k - 1) // 2
    xp = F.pad(x.transpose(1, 2), (pl, pr))                       # [N,Cin,T+k-1], explicit asymmetric padding
    want = F.conv1d(xp, w.permute(2, 1, 0), b).transpose(1, 2)    # weight [Cout,Cin,k]
    torch.testing.assert_close(got, want, rtol=1e-12, atol=1e-12)


def test_max_pool_same2_vs_functional():
    r = _rng(1)
    x = torch.from_numpy(r.standard_normal((2, 17, 6))).double()
    xp = F.pad(x.transpose(1, 2), (0, 1), value=float("-inf"))    # TF pads -inf on the right for pool 2, stride 1
    want = F.max_pool1d(xp, 2, 1).transpose(1, 2)
    torch.testing.assert_close(O.max_pool_same2(x), want, rtol=0, atol=0)


@pytest.mark.parametrize("mode", ["moving", "batch"])
def test_batch_norm_vs_functional(mode):
    r = _rng(2)
    x = torch.from_numpy(r.standard_normal((4, 11, 9))).double()
    g, b = (torch.from_numpy(r.standard_normal(9)).double() for _ in range(2))
    m = torch.from_numpy(r.standard_normal(9)).double()
    v = torch.from_numpy(r.uniform(0.5, 2.0, 9)).double()
    got = O.batch_norm(x, g, b, m, v, mode)
    flat = x.reshape(-1, 9)
    want = F.batch_norm(flat, m.clone(), v.clone(), g, b, training=(mode == "batch"), momentum=0.0, eps=1e-3)
    torch.testing.assert_close(got, want.reshape(x.shape), rtol=1e-12, atol=1e-12)


def test_dense_prenet_highway_vs_numpy():
    r = _rng(3)
    x = r.standard_normal((2, 5, 12))
    w = {"p/dense_1/kernel": r.standard_normal((12, 8)), "p/dense_1/bias": r.standard_normal(8),
         "p/dense_2/kernel": r.standard_normal((8, 6)), "p/dense_2/bias": r.standard_normal(6),
         "h/H/kernel": r.standard_normal((12, 12)), "h/H/bias": r.standard_normal(12),
         "h/T/kernel": r.standard_normal((12, 12)), "h/T/bias": r.standard_normal(12)}
    W = O.W(w, torch.float64)
    xt = torch.from_numpy(x)
    a1 = np.maximum(np.einsum("ntc,cd->ntd", x, w["p/dense_1/kernel"]) + w["p/dense_1/bias"], 0)
    a2 = np.maximum(np.einsum("ntc,cd->ntd", a1, w["p/dense_2/kernel"]) + w["p/dense_2/bias"], 0)
    np.testing.assert_allclose(O.prenet(xt, W, "p").numpy(), a2, rtol=1e-12, atol=1e-12)
    H = np.maximum(x @ w["h/H/kernel"] + w["h/H/bias"], 0)
    T = 1.0 / (1.0 + np.exp(-(x @ w["h/T/kernel"] + w["h/T/bias"])))
    np.testing.assert_allclose(O.highwaynet(xt, W, "h").numpy(), H * T + x * (1 - T), rtol=1e-12, atol=1e-12)


def test_gru_cell_vs_explicit_numpy_and_torch_where_they_agree():
    """TF GRUCell: [r|u] = sigmoid([x,h] Wg + bg); c = tanh([x, r*h] Wc + bc); h' = u h + (1-u) c.
    (a) against an explicit numpy evaluation written from Appendix B.1;
    (b) torch.nn.GRUCell applies the reset gate AFTER the recurrent matmul, so the two only coincide when the candidate's
        recurrent kernel is diagonal-free in effect (zero): checked in that regime with the gate mapping z = u, n = c."""
    r = _rng(4)
    d, n, N = 6, 5, 3
    x, h = r.standard_normal((N, d)), r.standard_normal((N, n))
    wg, bg = r.standard_normal((d + n, 2 * n)), r.standard_normal(2 * n)
    wc, bc = r.standard_normal((d + n, n)), r.standard_normal(n)
    Wt = O.W({"g/gates/kernel": wg, "g/gates/bias": bg, "g/candidate/kernel": wc, "g/candidate/bias": bc}, torch.float64)
    got = O.gru_cell(torch.from_numpy(x), torch.from_numpy(h), Wt, "g").numpy()
    sig = lambda a: 1.0 / (1.0 + np.exp(-a))
    gates = sig(np.concatenate([x, h], 1) @ wg + bg)
    rr, uu = gates[:, :n], gates[:, n:]
    c = np.tanh(np.concatenate([x, rr * h], 1) @ wc + bc)
    np.testing.assert_allclose(got, uu * h + (1 - uu) * c, rtol=1e-12, atol=1e-12)
    # (b) zero recurrent candidate kernel: TF and torch conventions coincide
    wc0 = wc.copy(); wc0[d:] = 0.0
    Wt0 = O.W({"g/gates/kernel": wg, "g/gates/bias": bg, "g/candidate/kernel": wc0, "g/candidate/bias": bc}, torch.float64)
    got0 = O.gru_cell(torch.from_numpy(x), torch.from_numpy(h), Wt0, "g")
    cell = torch.nn.GRUCell(d, n).double()
    with torch.no_grad():   # torch gate order (r, z, n); h' = (1-z) n + z h  ->  z = u
        cell.weight_ih.copy_(torch.from_numpy(np.concatenate([wg[:d, :n], wg[:d, n:], wc0[:d]], 1).T))
        cell.weight_hh.copy_(torch.from_numpy(np.concatenate([wg[d:, :n], wg[d:, n:], np.zeros((n, n))], 1).T))
        cell.bias_ih.copy_(torch.from_numpy(np.concatenate([bg[:n], bg[n:], bc])))
        cell.bias_hh.zero_()
        want0 = cell(torch.from_numpy(x), torch.from_numpy(h))
    torch.testing.assert_close(got0, want0, rtol=1e-12, atol=1e-12)


def test_bigru_vs_packed_reference_loop():
    """bidirectional_dynamic_rnn with sequence_length: per utterance, run the cell over its own length only (forward) and
    over its own reversed prefix (backward); zero outputs past the length."""
    r = _rng(5)
    N, T, n = 3, 9, 128
    lens = np.array([9, 4, 1])
    x = r.standard_normal((N, T, 128)) * 0.3
    w = {}
    for d in ("fw", "bw"):
        p = f"s/bidirectional_rnn/{d}/gru_cell/"
        w[p + "gates/kernel"] = r.standard_normal((256, 256)) * 0.1
        w[p + "gates/bias"] = np.ones(256)
        w[p + "candidate/kernel"] = r.standard_normal((256, 128)) * 0.1
        w[p + "candidate/bias"] = np.zeros(128)
    W = O.W(w, torch.float64)
    got = O.bigru(torch.from_numpy(x), lens, W, "s").numpy()
    want = np.zeros((N, T, 256))
    for i in range(N):
        L = lens[i]
        h = torch.zeros(1, n, dtype=torch.float64)
        for t in range(L):
            h = O.gru_cell(torch.from_numpy(x[i:i + 1, t]), h, W, "s/bidirectional_rnn/fw/gru_cell")
            want[i, t, :n] = h.numpy()[0]
        h = torch.zeros(1, n, dtype=torch.float64)
        for t in range(L - 1, -1, -1):
            h = O.gru_cell(torch.from_numpy(x[i:i + 1, t]), h, W, "s/bidirectional_rnn/bw/gru_cell")
            want[i, t, n:] = h.numpy()[0]
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)


def test_decoder_step_vs_numpy():
    """One full step of the output cell (Appendix B.5: prenet on [frame | previous context], attention GRU, Bahdanau
    score / unmasked softmax / context, 512->256 projection, two residual GRUs, 80r projection) against a numpy
    evaluation written independently of the oracle's helpers."""
    from tacotron_multispeaker_b200.hparams import HParams
    from tacotron_multispeaker_b200.weights import random_init
    hp = HParams(outputs_per_step=3)
    wd = random_init(hp, id_num=4, seed=3)
    W = O.W(wd, torch.float64)
    g = lambda name: np.asarray(wd["model/inference/" + name], np.float64)
    r = _rng(6)
    N, T = 2, 7
    mem = r.standard_normal((N, T, 256)) * 0.5
    x = r.uniform(0, 1, (N, 80))
    st = [r.standard_normal((N, 256)) * 0.3 for _ in range(4)]   # h_att, ctx, h1, h2
    keys = mem @ g("memory_layer/kernel")
    out, a, ns = O.decoder_step(torch.from_numpy(x), O.DecoderState(*(torch.from_numpy(s) for s in st)),
                                torch.from_numpy(mem), torch.from_numpy(keys), W)
    sig = lambda z: 1.0 / (1.0 + np.exp(-z))

    def gru(xx, h, scope):
        n = h.shape[1]
        gt = sig(np.concatenate([xx, h], 1) @ g(scope + "/gates/kernel") + g(scope + "/gates/bias"))
        rr, uu = gt[:, :n], gt[:, n:]
        c = np.tanh(np.concatenate([xx, rr * h], 1) @ g(scope + "/candidate/kernel") + g(scope + "/candidate/bias"))
        return uu * h + (1 - uu) * c

    att = O._ATT; dpw = O._DPW; mrc = O._MRC
    p = np.concatenate([x, st[1]], 1)
    for i in (1, 2):
        p = np.maximum(p @ g(dpw + "decoder_prenet/dense_%d/kernel" % i) + g(dpw + "decoder_prenet/dense_%d/bias" % i), 0)
    h_att = gru(p, st[0], dpw + "gru_cell")
    pq = h_att @ g(att + "bahdanau_attention/query_layer/kernel")
    score = np.einsum("k,ntk->nt", g(att + "bahdanau_attention/attention_v"), np.tanh(keys + pq[:, None, :]))
    al = np.exp(score - score.max(1, keepdims=True)); al /= al.sum(1, keepdims=True)
    ctx = np.einsum("nt,ntd->nd", al, mem)
    y0 = np.concatenate([h_att, ctx], 1) @ g(mrc + "cell_0/output_projection_wrapper/kernel") + g(mrc + "cell_0/output_projection_wrapper/bias")
    h1 = gru(y0, st[2], mrc + "cell_1/gru_cell"); y1 = y0 + h1
    h2 = gru(y1, st[3], mrc + "cell_2/gru_cell"); y2 = y1 + h2
    o = y2 @ g("decoder/output_projection_wrapper/kernel") + g("decoder/output_projection_wrapper/bias")
    np.testing.assert_allclose(a.numpy(), al, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(ns.ctx.numpy(), ctx, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(ns.h_att.numpy(), h_att, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(ns.h1.numpy(), h1, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(ns.h2.numpy(), h2, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(out.numpy(), o, rtol=1e-11, atol=1e-12)
