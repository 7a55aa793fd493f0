"""Developer aid: stage times of the whole forward (config 3) + batch-1 latency sweep + config-4-sized decode."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_inputs
from tacotron_multispeaker_b200.engine import Engine
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init

def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

which = sys.argv[1] if len(sys.argv) > 1 else "all"
hp = HParams(outputs_per_step=5, max_iters=200)
w = random_init(hp, 60, seed=1234)
eng = Engine(hp, 60); eng.load_weights(w); eng.set_profiling(True)
if which in ("all", "stages"):
    ids, lengths, spk = make_inputs(32, 100, 60, 1, min_len=60, vocab=(7108, 7325))
    ms = timed(lambda: eng.forward(ids, lengths, spk))
    print("config 3 forward %.3f ms; stages %s" % (ms, eng.last_stage_ms()), flush=True)
if which in ("all", "lat"):
    for T_in in (20, 60, 100, 140, 200):
        ids, lengths, spk = make_inputs(1, T_in, 60, 3, min_len=T_in, vocab=(7108, 7325))
        ms = timed(lambda: eng.forward(ids, lengths, spk), 10)
        print("batch 1, T_in=%3d: forward %.3f ms  stages %s geometry %s" % (T_in, ms, {k: round(v, 3) for k, v in eng.last_stage_ms().items()}, eng.decoder_geometry(1)), flush=True)
if which in ("all", "big"):
    for N, T_in in ((64, 100), (128, 60), (256, 60)):
        ids, lengths, spk = make_inputs(N, T_in, 60, 2, min_len=20, vocab=(2, 7054))
        ms = timed(lambda: eng.forward(ids, lengths, spk), 3)
        print("batch %d, T_in=%d: forward %.3f ms -> %.2f M frames/s; stages %s geometry %s" % (N, T_in, ms, N * 1000 / ms / 1e3, {k: round(v, 3) for k, v in eng.last_stage_ms().items()}, eng.decoder_geometry(N)), flush=True)
eng.close()
