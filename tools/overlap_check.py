"""Developer aid: does a big pinned D2H copy on one stream overlap a forward on another stream?"""
import os, sys, time, threading
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_inputs
from tacotron_multispeaker_b200.engine import Engine
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init
hp = HParams(outputs_per_step=5, max_iters=200)
w = random_init(hp, 60, seed=1234)
eng = Engine(hp, 60); eng.load_weights(w)
ids, lengths, spk = make_inputs(32, 100, 60, 1, min_len=60, vocab=(7108, 7325))
dev_buf = torch.empty(36_000_000, dtype=torch.float32, device="cuda")
host_buf = torch.empty(36_000_000, dtype=torch.float32).pin_memory()
s_copy, s_comp = torch.cuda.Stream(), torch.cuda.Stream()
def copies(k):
    with torch.cuda.stream(s_copy):
        for _ in range(k): host_buf.copy_(dev_buf, non_blocking=True)
def forwards(k):
    with torch.cuda.stream(s_comp):
        for _ in range(k): eng.forward(ids, lengths, spk)
for _ in range(2): forwards(1); copies(1)
torch.cuda.synchronize()
def wall(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
K = 6
print("copies alone   : %.2f ms each" % (wall(lambda: copies(K)) / K))
print("forwards alone : %.2f ms each" % (wall(lambda: forwards(K)) / K))
def both():
    t = threading.Thread(target=copies, args=(K,)); t.start(); forwards(K); t.join()
print("both together  : %.2f ms per (copy + forward) pair" % (wall(both) / K))
eng.close()
