"""Developer aid: wall-clock timeline of taco_forward_host calls on L lanes (threads)."""
import os, sys, time, threading
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_inputs
from tacotron_multispeaker_b200 import _abi
from tacotron_multispeaker_b200.engine import Engine
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init
L = int(sys.argv[1]) if len(sys.argv) > 1 else 2
hp = HParams(outputs_per_step=5, max_iters=200)
w = random_init(hp, 60, seed=1234)
ids, lengths, spk = make_inputs(32, 100, 60, 1, min_len=60, vocab=(7108, 7325))
def pinned(shape, dt): return torch.empty(shape, dtype=dt).pin_memory().numpy()
lanes = []
for i in range(L):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        e = Engine(hp, 60); e.load_weights(w)
    b = dict(ids=pinned((32, 100), torch.int32), lens=pinned((32,), torch.int32), spk=pinned((32,), torch.int32),
             mel=pinned((32, 1000, 80), torch.float32), lin=pinned((32, 1000, 1025), torch.float32), al=pinned((32, 100, 200), torch.float32))
    b["ids"][:], b["lens"][:], b["spk"][:] = ids, lengths, spk
    lanes.append((e, st, b))
log = []
STAGGER = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
def work(i, k):
    e, st, b = lanes[i]
    if STAGGER and k > 2: time.sleep(i * STAGGER * 1e-3)
    with torch.cuda.stream(st):
        for j in range(k):
            t0 = time.perf_counter()
            e.forward_host(b["ids"], b["lens"], b["spk"], None, False, _abi.BN_MOVING, b["mel"], b["lin"], b["al"])
            log.append((i, j, t0, time.perf_counter()))
for i in range(L): work(i, 2)
torch.cuda.synchronize(); log.clear()
T0 = time.perf_counter()
ths = [threading.Thread(target=work, args=(i, 6)) for i in range(L)]
[t.start() for t in ths]; [t.join() for t in ths]
torch.cuda.synchronize()
T1 = time.perf_counter()
for i, j, a, b in sorted(log, key=lambda r: r[2])[-2 * L:]: print("lane %d call %d: start %7.2f end %7.2f  (%.2f ms)" % (i, j, (a - T0) * 1e3, (b - T0) * 1e3, (b - a) * 1e3))
print("total %.2f ms for %d batches -> %.2f ms per batch" % ((T1 - T0) * 1e3, 6 * L, (T1 - T0) * 1e3 / (6 * L)))
