"""Golden fixtures (fp64 oracle outputs, tests/golden/make_golden.py):
CPU: the fp32 oracle reproduces them; GPU: the CUDA path reproduces them."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import taco_oracle as O
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
              if not os.path.basename(p).startswith(("audio_", "ref_")))   # ref_*: tests/test_ref_pinned.py; audio_*.npz: tests/test_audio_oracle.py, test_gpu_audio.py


def load(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    hp = HParams(outputs_per_step=int(g["r"]), max_iters=int(g["max_iters"]))
    id_num = int(g["id_num"])
    w = random_init(hp, id_num, seed=int(g["wseed"]), randomize_bn=True)
    return g, hp, id_num, w, str(g["mode"])


def test_fixtures_exist():
    assert len(GOLD) >= 4


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_fp32_matches_golden(path):
    g, hp, id_num, w, mode = load(path)
    spk = g["spk"] if id_num > 1 else None
    kw = dict(identities=spk, id_num=id_num)
    if mode == "free":
        out = O.tacotron_forward(w, hp, g["ids"], g["lengths"], **kw)
    elif mode == "teacher_batch":
        out = O.tacotron_forward(w, hp, g["ids"], g["lengths"], mel_targets=g["mel_targets"],
                                 linear_targets=np.zeros(g["linear"].shape, np.float32), **kw)
    else:
        out = O.tacotron_forward(w, hp, g["ids"], g["lengths"], mel_targets=g["mel_targets"], teacher_force=True,
                                 bn_mode="moving", **kw)
    assert out["steps"] == int(g["steps"])
    assert np.abs(out["mel_outputs"].numpy() - g["mel"]).max() < 1e-4
    assert np.abs(out["linear_outputs"].numpy() - g["linear"]).max() < 1e-4
    assert np.abs(out["alignments"].numpy() - g["alignments"]).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_cuda_matches_golden(path):
    from tacotron_multispeaker_b200.tacotron import Tacotron
    g, hp, id_num, w, mode = load(path)
    m = Tacotron(hp, verbose=False)
    m.load_weights(w)
    spk = g["spk"] if id_num > 1 else None
    if mode == "free":
        m.initialize(g["ids"], g["lengths"], identities=spk, id_num=id_num)
    elif mode == "teacher_batch":      # reference-faithful training forward: linear_targets given
        m.initialize(g["ids"], g["lengths"], mel_targets=g["mel_targets"],
                     linear_targets=np.zeros(g["linear"].shape, np.float32), identities=spk, id_num=id_num)
    else:
        m.initialize(g["ids"], g["lengths"], mel_targets=g["mel_targets"], identities=spk, id_num=id_num,
                     teacher_force=True)
    assert m.steps == int(g["steps"])
    tol = 1e-3                          # north_star: 1e-3 max-abs, fp32 mode
    assert np.abs(m.mel_outputs.cpu().numpy() - g["mel"]).max() < tol
    assert np.abs(m.linear_outputs.cpu().numpy() - g["linear"]).max() < tol
    assert np.abs(m.alignments.cpu().numpy() - g["alignments"]).max() < 1e-4
    if mode != "free":                  # identical per-step attention argmax
        assert np.array_equal(m.alignments.cpu().numpy().argmax(axis=1), g["alignments"].argmax(axis=1))
