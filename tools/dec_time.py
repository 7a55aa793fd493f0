"""Developer aid: decoder kernel time per step for a few geometries and implementations (TACO_DEC_IMPL)."""
import os, sys
os.environ.setdefault("TACO_DEV", "1")   # per-call developer switches of the C ABI
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_inputs
from tacotron_multispeaker_b200.engine import Engine
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init


def run(impl, N, T_in, S, iters=200, teacher=False):
    os.environ["TACO_DEC_IMPL"] = impl
    if S:
        os.environ["TACO_DEC_S"] = str(S)
    else:
        os.environ.pop("TACO_DEC_S", None)
    hp = HParams(outputs_per_step=5, max_iters=iters)
    w = random_init(hp, 60, seed=1234)
    ids, lengths, spk = make_inputs(N, T_in, 60, 1, min_len=max(1, int(T_in * 0.6)), vocab=(7108, 7325))
    eng = Engine(hp, 60); eng.load_weights(w); eng.set_profiling(True)
    mem = eng.encoder(ids, lengths, spk, 0)
    tg = torch.rand(N, iters * 5, 80, device="cuda") if teacher else None
    for _ in range(2):
        eng.decode(mem, tg, teacher, True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.decode(mem, tg, teacher, True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print("%-4s N=%-3d T_in=%-3d S=%s %s: %.3f ms  %.2f us/step" % (impl, N, T_in, S, "teacher" if teacher else "free", ms, ms * 1e3 / iters), flush=True)
    eng.close()


if __name__ == "__main__":
    impls = os.environ.get("IMPLS", "cw").split(",")   # "cw,v2": also the fp32 FFMA decoder (TACO_DEC_IMPL is read at taco_finalize_weights)
    for impl in impls:
        run(impl, 32, 100, 5)
        run(impl, 32, 100, 8)
        run(impl, 1, 100, 1)
        run(impl, 32, 100, 5, teacher=True)
