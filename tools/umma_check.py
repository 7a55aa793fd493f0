"""Developer aid: tcgen05 conv kernel vs the oracle on a set of shapes, all GEMM modes."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import taco_oracle as O
from tacotron_multispeaker_b200.engine import Engine
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init

hp = HParams(outputs_per_step=5, max_iters=4)
eng = Engine(hp, 0)
eng.load_weights(random_init(hp, 0, seed=1))
shapes = [(1, 128, 128, 2, 100), (1, 64, 128, 1, 37), (3, 128, 128, 3, 129), (2, 80, 128, 2, 260), (4, 128, 128, 1, 50),
          (16, 128, 128, 1, 100), (3, 256, 80, 2, 300), (1, 256, 1025, 2, 200), (3, 2048, 128, 1, 100), (1, 320, 256, 2, 77),
          (7, 80, 128, 1, 1000)]
for mode in (1, 2, 0):
    eng.set_gemm_mode(mode)
    for (k, cin, cout, N, T) in shapes:
        rng = np.random.default_rng(k * 31 + cin + cout)
        x = rng.standard_normal((N, T, cin)).astype(np.float32)
        w = (rng.standard_normal((k, cin, cout)) / np.sqrt(k * cin)).astype(np.float32)
        b = rng.standard_normal((cout,)).astype(np.float32)
        ref = torch.relu(O.conv1d_same(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b)))
        try:
            t0 = time.time()
            got = eng.conv1d(x, w, b, 1).cpu()
            torch.cuda.synchronize()
            d = (got - ref).abs()
            print("mode %d k=%2d cin=%4d cout=%4d N=%d T=%4d: max|d|=%.3e mean|d|=%.3e max|ref|=%.2f nan=%d (%.3fs)"
                  % (mode, k, cin, cout, N, T, float(d.max()), float(d.mean()), float(ref.abs().max()), int(torch.isnan(got).sum()), time.time() - t0), flush=True)
        except Exception as e:
            print("mode %d k=%d cin=%d cout=%d: EXC %s" % (mode, k, cin, cout, e), flush=True)
print("done")
