"""One forward of the benchmark shape (for ncu): python tools/profile_target.py [N] [T_in] [iters]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_inputs
from tacotron_multispeaker_b200.engine import Engine
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T_in = int(sys.argv[2]) if len(sys.argv) > 2 else 100
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
hp = HParams(outputs_per_step=5, max_iters=iters)
eng = Engine(hp, 60); eng.load_weights(random_init(hp, 60, seed=1234))
ids, lengths, spk = make_inputs(N, T_in, 60, 1, min_len=max(1, int(T_in * 0.6)), vocab=(7108, 7325))
for _ in range(2):
    mel, lin, al, s = eng.forward(ids, lengths, spk)
torch.cuda.synchronize()
print("ok", s, float(mel.abs().max()), eng.launch_count())
