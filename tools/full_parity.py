"""Developer aid: full-size free-running parity (config 3: N=32, T_in=100, 200 steps, r=5) of the CUDA path against the CPU
oracle, plus the same with the fp32 FFMA decoder (TACO_DEC_IMPL=v2) to separate rounding-model effects from recurrence growth."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_inputs
from oracle import taco_oracle as O
from tacotron_multispeaker_b200.engine import Engine
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init

N, T_in, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 100, int(sys.argv[2]) if len(sys.argv) > 2 else 200
hp = HParams(outputs_per_step=5, max_iters=steps)
w = random_init(hp, 60, seed=1234)
ids, lengths, spk = make_inputs(N, T_in, 60, 1, min_len=60, vocab=(7108, 7325))
t0 = time.time()
ref = O.tacotron_forward(w, hp, ids, lengths, identities=spk, id_num=60)
print("oracle %.2fs steps=%d" % (time.time() - t0, ref["steps"]), flush=True)
for impl in ("cw", "v2"):
    os.environ["TACO_DEC_IMPL"] = impl
    eng = Engine(hp, 60); eng.load_weights(w)
    mel, lin, al, s = eng.forward(ids, lengths, spk)
    torch.cuda.synchronize()
    mel, lin, al = mel.cpu(), lin.cpu(), al.cpu()
    T = ref["mel_outputs"].shape[1]
    for upto in (50, 250, 1000):
        k = min(upto, T)
        print("%s: first %4d frames: max|d| mel %.2e linear %.2e | align(first %d steps) %.2e | argmax equal %s" % (
            impl, k, float((mel[:, :k] - ref["mel_outputs"][:, :k]).abs().max()), float((lin[:, :k] - ref["linear_outputs"][:, :k]).abs().max()),
            k // 5, float((al[:, :, :k // 5] - ref["alignments"][:, :, :k // 5]).abs().max()),
            bool(torch.equal(al[:, :, :k // 5].argmax(1), ref["alignments"][:, :, :k // 5].argmax(1)))), flush=True)
    print("%s: steps %d geometry %s" % (impl, s, eng.decoder_geometry(N)), flush=True)
    eng.close()
