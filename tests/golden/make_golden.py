"""Generates tests/golden/*.npz: outputs of the fp64 oracle on small seeded cases.

The reference cannot run here (TensorFlow 1.x) and ships no
golden vectors, so these fixtures pin the ORACLE (regression anchor) rather
than the reference -- parity stays "unpinned" in the sense of DESIGN.md.
Weights are regenerated from the seed by weights.random_init, inputs are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import taco_oracle as O  # noqa: E402
from tacotron_multispeaker_b200.hparams import HParams  # noqa: E402
from tacotron_multispeaker_b200.weights import random_init  # noqa: E402

CASES = {
    # name: (r, max_iters, id_num, N, T_in, mode, weight seed, input seed)
    "free_multispeaker": (5, 6, 5, 3, 17, "free", 21, 1),
    "teacher_batchnorm": (5, 50, 5, 3, 13, "teacher_batch", 22, 2),
    "teacher_moving_r2": (2, 50, 4, 2, 11, "teacher_moving", 23, 3),
    "single_speaker_r1": (1, 9, 0, 2, 8, "free", 24, 4),
}


def make_case(r, max_iters, id_num, N, T_in, mode, wseed, iseed):
    hp = HParams(outputs_per_step=r, max_iters=max_iters)
    w = random_init(hp, id_num, seed=wseed, randomize_bn=True)
    rng = np.random.default_rng(iseed)
    lengths = rng.integers(max(1, T_in // 2), T_in + 1, (N,)).astype(np.int32)
    lengths[0] = T_in
    ids = rng.integers(2, hp.num_symbols, (N, T_in)).astype(np.int32)
    for i in range(N):
        ids[i, lengths[i]:] = 0
    spk = rng.integers(0, max(id_num, 1), (N,)).astype(np.int32)
    T_tgt = 4 * r
    mel_t = rng.uniform(0, 1, (N, T_tgt, hp.num_mels)).astype(np.float32)
    kw = dict(identities=spk if id_num > 1 else None, id_num=id_num, dtype=torch.float64)
    if mode == "free":
        out = O.tacotron_forward(w, hp, ids, lengths, **kw)
    elif mode == "teacher_batch":
        out = O.tacotron_forward(w, hp, ids, lengths, mel_targets=mel_t,
                                 linear_targets=np.zeros((N, T_tgt, hp.num_freq), np.float32), **kw)
    else:
        out = O.tacotron_forward(w, hp, ids, lengths, mel_targets=mel_t, teacher_force=True, bn_mode="moving", **kw)
    return dict(r=r, max_iters=max_iters, id_num=id_num, mode=mode, wseed=wseed, ids=ids, lengths=lengths, spk=spk,
                mel_targets=mel_t, steps=out["steps"],
                mel=out["mel_outputs"].numpy().astype(np.float32),
                linear=out["linear_outputs"].numpy().astype(np.float32),
                alignments=out["alignments"].numpy().astype(np.float32))


if __name__ == "__main__":
    for name, spec in CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **make_case(*spec))
        print("wrote", name)
