import os
import sys

os.environ.setdefault("TACO_DEV", "1")   # the C ABI reads its per-call developer switches (TACO_DEC_S, TACO_BIGRU, ...) only then

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """``-m gpu`` tests need a CUDA device: skip (not fail) them on a box without one."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def make_inputs(N, T_in, id_num, seed, min_len=None, vocab=(2, 7352)):
    """Synthetic batch in the reference feeder's layout (datasets/datafeeder_npy.py:163-171):
    ids padded with 0 beyond each length."""
    rng = np.random.default_rng(seed)
    lo = min_len if min_len is not None else max(1, T_in // 2)
    lengths = rng.integers(lo, T_in + 1, (N,)).astype(np.int32)
    lengths[rng.integers(0, N)] = T_in
    ids = rng.integers(vocab[0], vocab[1], (N, T_in)).astype(np.int32)
    for i in range(N):
        ids[i, lengths[i]:] = 0
    spk = rng.integers(0, max(id_num, 1), (N,)).astype(np.int32)
    return ids, lengths, spk


@pytest.fixture(scope="session")
def small_hp():
    from tacotron_multispeaker_b200.hparams import HParams
    return HParams(outputs_per_step=5, max_iters=8)


@pytest.fixture(scope="session")
def small_weights(small_hp):
    from tacotron_multispeaker_b200.weights import random_init
    return random_init(small_hp, id_num=6, seed=7, randomize_bn=True)
