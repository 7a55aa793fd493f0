"""CPU oracle: restatement of the reference's Tacotron forward graph.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tacotron_multispeaker_b200/`` may
import this module; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs do.

PARITY UNPINNED: the reference delegates all arithmetic to TensorFlow 1.3/1.4
(``tf.layers``, ``tf.contrib.rnn``, ``tf.contrib.seq2seq``), which is neither
vendored in ``/root/reference`` nor installable here, and the reference ships
no tests, golden vectors or checkpoints.  This file restates the graph that
``models/tacotron.py:35-104`` wires, with the TF 1.4 op semantics of
SURVEY.md Appendix B; every convention is isolated by an analytic known-answer
test in ``tests/test_oracle_kat.py``.

Eager, unfused, op-for-op in the reference's order, on torch CPU tensors
(``dtype`` float32 like the reference, or float64 as a tie-breaker).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

BN_EPS = 1e-3  # tf.layers.batch_normalization default epsilon


def _t(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(np.asarray(x)).to(dtype)


class W:
    """Weight lookup by short TF name (``model/inference/`` stripped)."""

    def __init__(self, weights: Dict[str, np.ndarray], dtype=torch.float32):
        self.dtype = dtype
        self._w = {}
        for k, v in weights.items():
            k = k[len("model/inference/"):] if k.startswith("model/inference/") else k
            self._w[k] = _t(v, dtype)

    def __call__(self, name):
        return self._w[name]

    def has(self, name):
        return name in self._w


# --------------------------------------------------------------------------
# Layers
# --------------------------------------------------------------------------
def dense(x, kernel, bias=None):
    """tf.layers.dense on the last axis."""
    y = x @ kernel
    return y if bias is None else y + bias


def conv1d_same(x, kernel, bias):
    """tf.layers.conv1d(padding='same', stride 1): cross-correlation with
    pad_left=(k-1)//2, pad_right=k-1-pad_left, zero padding.
    x [N,T,Cin], kernel [k,Cin,Cout] (reference models/modules.py:95-100)."""
    k, cin, cout = kernel.shape
    N, T, _ = x.shape
    pl = (k - 1) // 2
    pr = k - 1 - pl
    xp = torch.nn.functional.pad(x, (0, 0, pl, pr))
    y = torch.zeros(N, T, cout, dtype=x.dtype)
    for j in range(k):
        y = y + xp[:, j:j + T, :] @ kernel[j]
    return y + bias


def batch_norm(x, gamma, beta, mean, var, mode):
    """tf.layers.batch_normalization over the last axis, eps=1e-3.
    mode 'moving': inference statistics; mode 'batch': biased moments over
    axes (N,T) including padded rows (training=True forward)."""
    if mode == "batch":
        mean = x.mean(dim=(0, 1))
        var = ((x - mean) ** 2).mean(dim=(0, 1))
    return (x - mean) * (gamma / torch.sqrt(var + BN_EPS)) + beta


def conv1d_block(x, w: W, scope, activation, bn_mode):
    """reference conv1d(): conv -> activation -> batch norm (modules.py:93-101)."""
    y = conv1d_same(x, w(scope + "/conv1d/kernel"), w(scope + "/conv1d/bias"))
    if activation == "relu":
        y = torch.relu(y)
    bn = scope + "/batch_normalization/"
    return batch_norm(y, w(bn + "gamma"), w(bn + "beta"), w(bn + "moving_mean"),
                      w(bn + "moving_variance"), bn_mode)


def max_pool_same2(x):
    """tf.layers.max_pooling1d(pool_size=2, strides=1, padding='same'):
    out[t]=max(x[t],x[t+1]), out[T-1]=x[T-1] (pad -inf on the right)."""
    nxt = torch.cat([x[:, 1:, :], x[:, -1:, :]], dim=1)
    return torch.maximum(x, nxt)


def prenet(x, w: W, scope):
    """reference prenet(): two dense+ReLU; dropout is the identity because
    tf.layers.dropout is called without training= (modules.py:5-12)."""
    for i in (1, 2):
        x = torch.relu(dense(x, w(f"{scope}/dense_{i}/kernel"), w(f"{scope}/dense_{i}/bias")))
    return x


def highwaynet(x, w: W, scope):
    """reference highwaynet(): H*T + x*(1-T) (modules.py:77-90)."""
    H = torch.relu(dense(x, w(scope + "/H/kernel"), w(scope + "/H/bias")))
    T = torch.sigmoid(dense(x, w(scope + "/T/kernel"), w(scope + "/T/bias")))
    return H * T + x * (1.0 - T)


def gru_cell(x, h, w: W, scope):
    """tf.contrib.rnn.GRUCell (Appendix B.1): rows of the kernels are
    [input ; state]; r = first half of the gate columns, u = second half;
    the reset gate multiplies the state BEFORE the candidate matmul."""
    n = h.shape[-1]
    g = torch.sigmoid(dense(torch.cat([x, h], -1), w(scope + "/gates/kernel"), w(scope + "/gates/bias")))
    r, u = g[..., :n], g[..., n:]
    c = torch.tanh(dense(torch.cat([x, r * h], -1), w(scope + "/candidate/kernel"),
                         w(scope + "/candidate/bias")))
    return u * h + (1.0 - u) * c


def bigru(x, lengths, w: W, scope):
    """tf.nn.bidirectional_dynamic_rnn(GRUCell(128), GRUCell(128), x,
    sequence_length=lengths) -> concat(fw, bw) (modules.py:68-74).
    With lengths: outputs are zero and the state is copied through for
    t >= len; the backward pass runs on reverse_sequence(x, len)."""
    N, T, _ = x.shape
    n = 128
    if lengths is None:
        lens = torch.full((N,), T, dtype=torch.long)
    else:
        lens = torch.as_tensor(np.asarray(lengths)).long()
    out = torch.zeros(N, T, 2 * n, dtype=x.dtype)
    # forward
    h = torch.zeros(N, n, dtype=x.dtype)
    for t in range(T):
        hn = gru_cell(x[:, t], h, w, scope + "/bidirectional_rnn/fw/gru_cell")
        live = (t < lens).unsqueeze(1)
        h = torch.where(live, hn, h)
        out[:, t, :n] = torch.where(live, hn, torch.zeros_like(hn))
    # backward: step s consumes position len-1-s
    h = torch.zeros(N, n, dtype=x.dtype)
    ar = torch.arange(N)
    for s in range(T):
        pos = lens - 1 - s
        live = (pos >= 0).unsqueeze(1)
        posc = pos.clamp(min=0)
        hn = gru_cell(x[ar, posc], h, w, scope + "/bidirectional_rnn/bw/gru_cell")
        h = torch.where(live, hn, h)
        cur = out[ar, posc, n:]
        out[ar, posc, n:] = torch.where(live, hn, cur)
    return out


def cbhg(x, lengths, w: W, scope, K, bn_mode):
    """reference cbhg() (modules.py:35-74)."""
    bank = torch.cat([conv1d_block(x, w, f"{scope}/conv_bank/conv1d_{k}", "relu", bn_mode)
                      for k in range(1, K + 1)], dim=-1)
    pooled = max_pool_same2(bank)
    p1 = conv1d_block(pooled, w, scope + "/proj_1", "relu", bn_mode)
    p2 = conv1d_block(p1, w, scope + "/proj_2", None, bn_mode)
    hw = p2 + x
    if hw.shape[2] != 128:
        hw = dense(hw, w(scope + "/dense/kernel"), w(scope + "/dense/bias"))
    for i in range(1, 5):
        hw = highwaynet(hw, w, f"{scope}/highway_{i}")
    return bigru(hw, lengths, w, scope)


# --------------------------------------------------------------------------
# Front end (reference models/tacotron.py:40-62)
# --------------------------------------------------------------------------
def embed(ids, spk, w: W):
    """embedding_lookup of the text table, and -- multi-speaker -- the speaker
    row tiled over T_in and concatenated (tacotron.py:46-55).  ``spk`` None
    means the single-speaker branch."""
    ids = torch.as_tensor(np.asarray(ids)).long()
    table = w("embedding")
    if ids.min() < 0 or ids.max() >= table.shape[0]:
        raise IndexError("symbol id out of range")  # TF CPU gather raises
    e = table[ids]
    if spk is not None:
        spk = torch.as_tensor(np.asarray(spk)).long()
        tid = w("embedding_id")
        if spk.min() < 0 or spk.max() >= tid.shape[0]:
            raise IndexError("speaker id out of range")
        s = tid[spk].unsqueeze(1).expand(-1, e.shape[1], -1)
        e = torch.cat([e, s], dim=2)
    return e


def encoder(ids, lengths, spk, w: W, bn_mode):
    e = embed(ids, spk, w)
    p = prenet(e, w, "prenet")
    return cbhg(p, lengths, w, "encoder_cbhg", 16, bn_mode)


# --------------------------------------------------------------------------
# Decoder (Appendix B.2, B.3, B.5)
# --------------------------------------------------------------------------
_ATT = ("decoder/output_projection_wrapper/multi_rnn_cell/cell_0/output_projection_wrapper/"
        "concat_output_and_attention_wrapper/attention_wrapper/")
_DPW = _ATT + "decoder_prenet_wrapper/"
_MRC = "decoder/output_projection_wrapper/multi_rnn_cell/"


@dataclass
class DecoderState:
    h_att: torch.Tensor
    ctx: torch.Tensor
    h1: torch.Tensor
    h2: torch.Tensor


def decoder_step(x, st: DecoderState, memory, keys, w: W):
    """One step of output_cell (tacotron.py:66-83).  Returns (out[N,80r],
    alignments[N,T_in], new state)."""
    # AttentionWrapper: cell_input_fn concatenates the PREVIOUS context.
    cell_in = torch.cat([x, st.ctx], -1)                      # [N, 80+256]
    p = prenet(cell_in, w, _DPW + "decoder_prenet")           # rnn_wrappers.py:22-24
    h_att = gru_cell(p, st.h_att, w, _DPW + "gru_cell")
    # BahdanauAttention (normalize=False, no memory mask)
    pq = h_att @ w(_ATT + "bahdanau_attention/query_layer/kernel")
    v = w(_ATT + "bahdanau_attention/attention_v")
    score = (v * torch.tanh(keys + pq.unsqueeze(1))).sum(-1)  # [N, T_in]
    a = torch.softmax(score, dim=-1)
    ctx = (a.unsqueeze(1) @ memory).squeeze(1)                # [N, 256]
    # ConcatOutputAndAttentionWrapper -> OutputProjectionWrapper(256)
    y0 = dense(torch.cat([h_att, ctx], -1), w(_MRC + "cell_0/output_projection_wrapper/kernel"),
               w(_MRC + "cell_0/output_projection_wrapper/bias"))
    # ResidualWrapper(GRUCell(256)) x2
    h1 = gru_cell(y0, st.h1, w, _MRC + "cell_1/gru_cell")
    y1 = y0 + h1
    h2 = gru_cell(y1, st.h2, w, _MRC + "cell_2/gru_cell")
    y2 = y1 + h2
    out = dense(y2, w("decoder/output_projection_wrapper/kernel"),
                w("decoder/output_projection_wrapper/bias"))
    return out, a, DecoderState(h_att, ctx, h1, h2)


def decode(memory, w: W, num_mels, r, max_iters, mel_targets=None, teacher_force=False,
           trace: Optional[dict] = None):
    """dynamic_decode(BasicDecoder(output_cell, helper, zero_state),
    maximum_iterations=max_iters), impute_finished=False (tacotron.py:84-94,
    helpers.py).  Returns (decoder_outputs [N,steps,80r], alignments
    [N,T_in,steps], steps)."""
    N = memory.shape[0]
    dt = memory.dtype
    keys = memory @ w("memory_layer/kernel")
    st = DecoderState(*(torch.zeros(N, 256, dtype=dt) for _ in range(4)))
    x = torch.zeros(N, num_mels, dtype=dt)                     # _go_frames
    finished = torch.zeros(N, dtype=torch.bool)
    if teacher_force:
        fed = mel_targets[:, r - 1::r, :]                      # helpers.py:48
        n_steps = fed.shape[1]
    outs, aligns = [], []
    t = 0
    while not bool(finished.all()):
        out, a, st = decoder_step(x, st, memory, keys, w)
        if teacher_force:
            fin = torch.full((N,), t + 1 >= n_steps)           # helpers.py:73
            x = fed[:, min(t, n_steps - 1), :]
        else:
            fin = (out == 0.0).all(dim=1)                      # helpers.py:35
            x = out[:, -num_mels:]
        finished = fin | finished | (t + 1 >= max_iters)
        outs.append(out)
        aligns.append(a)
        if trace is not None:
            trace.setdefault("h_att", []).append(st.h_att.clone())
            trace.setdefault("ctx", []).append(st.ctx.clone())
        t += 1
    dec = torch.stack(outs, dim=1)
    al = torch.stack(aligns, dim=0).permute(1, 2, 0).contiguous()   # tacotron.py:104
    return dec, al, t


# --------------------------------------------------------------------------
# Whole graph (reference Tacotron.initialize, models/tacotron.py:18-113)
# --------------------------------------------------------------------------
def tacotron_forward(weights, hp, inputs, input_lengths, mel_targets=None, linear_targets=None,
                     identities=None, id_num=0, dtype=torch.float32, teacher_force=None,
                     bn_mode=None, stages: Optional[dict] = None):
    """Returns dict(mel_outputs [N,T_out,80], linear_outputs [N,T_out,F],
    alignments [N,T_in,steps], steps).

    Reference-faithful defaults: is_training == (linear_targets is not None)
    (tacotron.py:36) selects teacher forcing AND batch-statistics BN.
    ``teacher_force`` / ``bn_mode`` override that (our extension: teacher
    forcing with moving-statistics BN when only mel_targets is given)."""
    w = weights if isinstance(weights, W) else W(weights, dtype)
    is_training = linear_targets is not None
    if teacher_force is None:
        teacher_force = is_training
    if bn_mode is None:
        bn_mode = "batch" if is_training else "moving"
    multi = identities is not None and id_num > 1             # tacotron.py:48
    M, r = hp.num_mels, hp.outputs_per_step
    memory = encoder(inputs, input_lengths, identities if multi else None, w, bn_mode)
    tg = _t(mel_targets, w.dtype) if teacher_force else None
    dec, al, steps = decode(memory, w, M, r, hp.max_iters, tg, teacher_force)
    N = dec.shape[0]
    mel = dec.reshape(N, -1, M)                               # tacotron.py:97
    post = cbhg(mel, None, w, "post_cbhg", 8, bn_mode)        # tacotron.py:100
    lin = dense(post, w("dense/kernel"), w("dense/bias"))     # tacotron.py:101
    if stages is not None:
        stages.update(memory=memory, decoder_outputs=dec, post=post)
    return dict(mel_outputs=mel, linear_outputs=lin, alignments=al, steps=steps)
