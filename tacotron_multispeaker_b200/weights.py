"""Variable inventory of the synthesis path, keyed by TF checkpoint names.

Names, shapes and initializers follow the variable scopes opened by the
reference graph (``models/tacotron.py:35-101``, ``models/modules.py:5-101``,
``models/rnn_wrappers.py:22-24``) under the ``model/inference/`` prefix that
``synthesizer.py:19`` / ``train.py:101`` + ``models/tacotron.py:35`` create.
Only ``model/inference/embedding_id`` is confirmed by the reference itself
(``synthesizer.py:25``); the rest follow TF 1.4 Layer/RNNCell scope rules, so
the loader also matches by suffix (see :func:`canonicalize`).

There is no network in this environment, so :func:`random_init` reproduces the
reference's initializers (truncated normal sigma=0.5 for the embeddings,
glorot-uniform kernels, constant biases) with a seeded numpy generator.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

from .hparams import HParams

PREFIX = "model/inference/"

_ATT = ("decoder/output_projection_wrapper/multi_rnn_cell/cell_0/output_projection_wrapper/"
        "concat_output_and_attention_wrapper/attention_wrapper/")
_DPW = _ATT + "decoder_prenet_wrapper/"
_MRC = "decoder/output_projection_wrapper/multi_rnn_cell/"


def _cbhg_specs(scope: str, K: int, in_dim: int, proj: Tuple[int, int], specs):
    """Variables of ``cbhg`` (reference ``models/modules.py:35-74``)."""
    def conv(name, k, cin, cout):
        specs[f"{name}/conv1d/kernel"] = ((k, cin, cout), "glorot")
        specs[f"{name}/conv1d/bias"] = ((cout,), "zeros")
        specs[f"{name}/batch_normalization/gamma"] = ((cout,), "ones")
        specs[f"{name}/batch_normalization/beta"] = ((cout,), "zeros")
        specs[f"{name}/batch_normalization/moving_mean"] = ((cout,), "zeros")
        specs[f"{name}/batch_normalization/moving_variance"] = ((cout,), "ones")

    for k in range(1, K + 1):
        conv(f"{scope}/conv_bank/conv1d_{k}", k, in_dim, 128)
    conv(f"{scope}/proj_1", 3, K * 128, proj[0])
    conv(f"{scope}/proj_2", 3, proj[0], proj[1])
    if proj[1] != 128:  # reference modules.py:59-60
        specs[f"{scope}/dense/kernel"] = ((proj[1], 128), "glorot")
        specs[f"{scope}/dense/bias"] = ((128,), "zeros")
    for i in range(1, 5):
        specs[f"{scope}/highway_{i}/H/kernel"] = ((128, 128), "glorot")
        specs[f"{scope}/highway_{i}/H/bias"] = ((128,), "zeros")
        specs[f"{scope}/highway_{i}/T/kernel"] = ((128, 128), "glorot")
        specs[f"{scope}/highway_{i}/T/bias"] = ((128,), "const:-1.0")  # modules.py:89
    for d in ("fw", "bw"):
        base = f"{scope}/bidirectional_rnn/{d}/gru_cell"
        specs[f"{base}/gates/kernel"] = ((256, 256), "glorot")
        specs[f"{base}/gates/bias"] = ((256,), "const:1.0")
        specs[f"{base}/candidate/kernel"] = ((256, 128), "glorot")
        specs[f"{base}/candidate/bias"] = ((128,), "zeros")


def weight_specs(hp: HParams, id_num: int = 0) -> "OrderedDict[str, Tuple[tuple, str]]":
    """``name (without prefix) -> (shape, initializer)`` for every variable the
    forward path reads.  ``id_num > 1`` adds the speaker table (reference
    ``models/tacotron.py:48-51``) and widens the encoder prenet input."""
    M, r, F = hp.num_mels, hp.outputs_per_step, hp.num_freq
    E = hp.embedding_text_channels
    multi = id_num > 1
    emb_dim = E + (hp.embedding_id_channels if multi else 0)
    s: "OrderedDict[str, Tuple[tuple, str]]" = OrderedDict()
    s["embedding"] = ((hp.num_symbols, E), "truncnorm:0.5")
    if multi:
        s["embedding_id"] = ((id_num, hp.embedding_id_channels), "truncnorm:0.5")
    s["prenet/dense_1/kernel"] = ((emb_dim, 256), "glorot")
    s["prenet/dense_1/bias"] = ((256,), "zeros")
    s["prenet/dense_2/kernel"] = ((256, 128), "glorot")
    s["prenet/dense_2/bias"] = ((128,), "zeros")
    _cbhg_specs("encoder_cbhg", 16, 128, (128, 128), s)
    s["memory_layer/kernel"] = ((256, 256), "glorot")
    s["decoder/output_projection_wrapper/kernel"] = ((256, M * r), "glorot")
    s["decoder/output_projection_wrapper/bias"] = ((M * r,), "zeros")
    s[_MRC + "cell_0/output_projection_wrapper/kernel"] = ((512, 256), "glorot")
    s[_MRC + "cell_0/output_projection_wrapper/bias"] = ((256,), "zeros")
    s[_DPW + "decoder_prenet/dense_1/kernel"] = ((M + 256, 256), "glorot")
    s[_DPW + "decoder_prenet/dense_1/bias"] = ((256,), "zeros")
    s[_DPW + "decoder_prenet/dense_2/kernel"] = ((256, 128), "glorot")
    s[_DPW + "decoder_prenet/dense_2/bias"] = ((128,), "zeros")
    s[_DPW + "gru_cell/gates/kernel"] = ((384, 512), "glorot")
    s[_DPW + "gru_cell/gates/bias"] = ((512,), "const:1.0")
    s[_DPW + "gru_cell/candidate/kernel"] = ((384, 256), "glorot")
    s[_DPW + "gru_cell/candidate/bias"] = ((256,), "zeros")
    s[_ATT + "bahdanau_attention/query_layer/kernel"] = ((256, 256), "glorot")
    s[_ATT + "bahdanau_attention/attention_v"] = ((256,), "glorot_v")
    for c in (1, 2):
        s[_MRC + f"cell_{c}/gru_cell/gates/kernel"] = ((512, 512), "glorot")
        s[_MRC + f"cell_{c}/gru_cell/gates/bias"] = ((512,), "const:1.0")
        s[_MRC + f"cell_{c}/gru_cell/candidate/kernel"] = ((512, 256), "glorot")
        s[_MRC + f"cell_{c}/gru_cell/candidate/bias"] = ((256,), "zeros")
    _cbhg_specs("post_cbhg", 8, M, (256, M), s)
    s["dense/kernel"] = ((256, F), "glorot")
    s["dense/bias"] = ((F,), "zeros")
    return s


def _glorot_limit(shape) -> float:
    # TF variance_scaling fan computation: receptive field * in / out channels.
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    else:
        rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    return float(np.sqrt(6.0 / (fan_in + fan_out)))


def _init(rng: np.random.Generator, shape, kind: str) -> np.ndarray:
    if kind == "zeros":
        return np.zeros(shape, np.float32)
    if kind == "ones":
        return np.ones(shape, np.float32)
    if kind.startswith("const:"):
        return np.full(shape, float(kind[6:]), np.float32)
    if kind.startswith("truncnorm:"):
        sigma = float(kind[10:])
        x = rng.standard_normal(shape)
        bad = np.abs(x) > 2.0
        while bad.any():  # tf.truncated_normal: resample beyond 2 sigma
            x[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(x) > 2.0
        return (x * sigma).astype(np.float32)
    if kind == "glorot":
        lim = _glorot_limit(shape)
        return rng.uniform(-lim, lim, shape).astype(np.float32)
    if kind == "glorot_v":  # attention_v [256]: get_variable default on a 1-D shape
        lim = float(np.sqrt(6.0 / (2 * shape[0])))
        return rng.uniform(-lim, lim, shape).astype(np.float32)
    raise ValueError(kind)


def random_init(hp: HParams, id_num: int = 0, seed: int = 1234,
                randomize_bn: bool = False) -> Dict[str, np.ndarray]:
    """Random weights with the reference's initializers, keyed by full TF
    names (``model/inference/...``).  ``randomize_bn`` perturbs the BN
    gamma/beta/moving statistics and the zero biases so that parity tests
    exercise every term (a freshly initialised BN is almost the identity)."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = OrderedDict()
    for name, (shape, kind) in weight_specs(hp, id_num).items():
        w = _init(rng, shape, kind)
        if randomize_bn:
            if name.endswith("/gamma"):
                w = rng.uniform(0.5, 1.5, shape).astype(np.float32)
            elif name.endswith("/beta") or name.endswith("/moving_mean"):
                w = rng.uniform(-0.3, 0.3, shape).astype(np.float32)
            elif name.endswith("/moving_variance"):
                w = rng.uniform(0.5, 2.0, shape).astype(np.float32)
            elif name.endswith("/bias") and kind == "zeros":
                w = rng.uniform(-0.1, 0.1, shape).astype(np.float32)
        out[PREFIX + name] = w
    return out


def canonicalize(weights: Dict[str, np.ndarray], hp: HParams, id_num: int) -> Dict[str, np.ndarray]:
    """Map a checkpoint-style dict onto the canonical short names.

    Accepts full names (``model/inference/x``), short names (``x``) or any
    name whose suffix is unique among the expected variables; optimizer slots
    (``.../Adam``, ``.../Adam_1``) and ``global_step`` are ignored.  Raises
    ``KeyError`` listing what is missing, ``ValueError`` on a shape mismatch.
    """
    specs = weight_specs(hp, id_num)
    out: Dict[str, np.ndarray] = {}
    leftovers = {}
    for name, arr in weights.items():
        if name.endswith("/Adam") or name.endswith("/Adam_1") or name == "global_step":
            continue
        short = name[len(PREFIX):] if name.startswith(PREFIX) else name
        if short in specs:
            out[short] = arr
        else:
            leftovers[name] = arr
    ambiguous = {}
    if leftovers:
        for short in specs:
            if short in out:
                continue
            hits = [n for n in leftovers if n.endswith("/" + short) or n.endswith(short)]
            if len(hits) == 1:
                out[short] = leftovers.pop(hits[0])
            elif len(hits) > 1:
                ambiguous[short] = sorted(hits)      # reported below instead of being skipped silently
    missing = [n for n in specs if n not in out]
    if missing:
        msg = "missing variables: " + ", ".join(PREFIX + m for m in missing[:8]) + (" ..." if len(missing) > 8 else "")
        amb = [m for m in missing if m in ambiguous]
        if amb:
            msg += "; ambiguous suffix matches: " + "; ".join("%s <- %s" % (m, ambiguous[m][:3]) for m in amb[:4])
        raise KeyError(msg)
    for short, (shape, _) in specs.items():
        a = np.ascontiguousarray(np.asarray(out[short], dtype=np.float32))
        if tuple(a.shape) != tuple(shape):
            raise ValueError("%s: shape %s, expected %s" % (PREFIX + short, a.shape, shape))
        out[short] = a
    return out


def count_params(hp: HParams, id_num: int = 0) -> int:
    return int(sum(int(np.prod(s)) for s, _ in weight_specs(hp, id_num).values()))
