"""Developer aid: decoder time per step over (cluster size, samples per cluster)."""
import os, sys, time
os.environ.setdefault("TACO_DEV", "1")   # per-call developer switches of the C ABI
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_inputs
from tacotron_multispeaker_b200.engine import Engine
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init

def run(N, T_in, iters, cs_list=(16, 8), s_list=(1, 2, 4, 8)):
    hp = HParams(outputs_per_step=5, max_iters=iters)
    w = random_init(hp, 60, seed=1234)
    ids, lengths, spk = make_inputs(N, T_in, 60, 1, min_len=max(1, int(T_in * 0.6)), vocab=(7108, 7325))
    for cs in cs_list:
        os.environ["TACO_DEC_CS"] = str(cs)
        eng = Engine(hp, 60); eng.load_weights(w); eng.set_profiling(True)
        mem = eng.encoder(ids, lengths, spk, 0)
        for S in s_list:
            os.environ["TACO_DEC_S"] = str(S)
            try:
                for _ in range(2):
                    eng.decode(mem, None, False, True)
                torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                for _ in range(3):
                    eng.decode(mem, None, False, True)
                ev1.record(); torch.cuda.synchronize()
                ms = ev0.elapsed_time(ev1) / 3
                print("N=%d T_in=%d CS=%d S=%d geom=%s: decode %.3f ms, %.2f us/step" % (N, T_in, cs, S, eng.decoder_geometry(N), ms, ms * 1e3 / iters), flush=True)
            except Exception as e:
                print("N=%d CS=%d S=%d failed: %s" % (N, cs, S, e), flush=True)
        os.environ.pop("TACO_DEC_S", None)
        eng.close()
    os.environ.pop("TACO_DEC_CS", None)

if __name__ == "__main__":
    for v3 in ("1", "0"):
        os.environ["TACO_DEC_V3"] = v3
        print("=== TACO_DEC_V3=%s ===" % v3, flush=True)
        run(32, 100, 200, cs_list=(16,))
        run(1, 50, 200, cs_list=(16,), s_list=(1,))
    os.environ["TACO_DEC_V3"] = "0"
    run(32, 100, 200, cs_list=(8,), s_list=(4,))
