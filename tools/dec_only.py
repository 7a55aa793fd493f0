"""Developer aid: run only the decoder loop kernel (for ncu): python tools/dec_only.py N T_in steps [calls]."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_inputs
from tacotron_multispeaker_b200.engine import Engine
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init
N, T_in, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 2
hp = HParams(outputs_per_step=5, max_iters=steps)
w = random_init(hp, 60, seed=1234)
ids, lengths, spk = make_inputs(N, T_in, 60, 1, min_len=max(1, int(T_in * 0.6)), vocab=(7108, 7325))
eng = Engine(hp, 60); eng.load_weights(w)
mem = eng.encoder(ids, lengths, spk, 0)
for _ in range(calls):
    eng.decode(mem, None, False, True)
torch.cuda.synchronize()
print("ok", eng.decoder_geometry(N))
