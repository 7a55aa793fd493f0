"""Device time of the Griffin-Lim vocoder (taco_griffin_lim) for a batch of config-3 sized utterances.
usage: python tools/gl_bench.py [N] [T] [iters] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tacotron_multispeaker_b200.engine import Engine  # noqa: E402
from tacotron_multispeaker_b200.hparams import HParams  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 100
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
eng = Engine(HParams(), id_num=0)
x = torch.rand(N, T, 1025, device=eng.device)
if len(sys.argv) > 5 and sys.argv[5] == "silent":      # what a random-init model produces: everything clipped to 0
    x = x * 0.0 - 0.1
for _ in range(2):
    eng.griffin_lim(x, iters)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    w = eng.griffin_lim(x, iters)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
frames = N * T
# per iteration and frame: read mags 4.1 KB + up to 7 neighbour rows (16 KB, L2) , write 4 KB; 2 x 2048-pt complex FFT
flops = frames * (iters * 2 + 1) * 5 * 2048 * 11
print("griffin_lim N=%d T=%d iters=%d: %.2f ms/batch, %.1f us/iteration, %.2f M frames/s vocoded, %.1f GFLOP/s (5 N log2 N), audio %.0fx real time"
      % (N, T, iters, ms, ms * 1e3 / max(iters, 1), frames / ms / 1e3, flops / ms / 1e6,
         (N * w.shape[-1] / 20000.0) / (ms / 1e3)))
