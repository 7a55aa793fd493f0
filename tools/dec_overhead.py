"""Developer aid: fixed cost of one decoder launch (prologue, cluster launch, cold first steps) vs. cost per step."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_inputs
from tacotron_multispeaker_b200.engine import Engine
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init

def timed(N, T_in, steps, reps=5):
    hp = HParams(outputs_per_step=5, max_iters=steps)
    w = random_init(hp, 60, seed=1234)
    ids, lengths, spk = make_inputs(N, T_in, 60, 1, min_len=max(1, int(T_in * 0.6)), vocab=(7108, 7325))
    eng = Engine(hp, 60); eng.load_weights(w)
    mem = eng.encoder(ids, lengths, spk, 0)
    for _ in range(2):
        eng.decode(mem, None, False, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.decode(mem, None, False, True)
    e1.record(); torch.cuda.synchronize()
    eng.close()
    return e0.elapsed_time(e1) / reps * 1e3   # us

for (N, T_in) in [(32, 100), (1, 50)]:
    t = {s: timed(N, T_in, s) for s in (1, 2, 11, 51, 201)}
    per = (t[201] - t[1]) / 200
    print("N=%d T_in=%d: decode call (kernel + find_steps) us:" % (N, T_in), {k: round(v, 1) for k, v in t.items()},
          "-> %.2f us/step, fixed %.1f us" % (per, t[1] - per), flush=True)
