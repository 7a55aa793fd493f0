"""Developer aid: per-phase clock stamps of the decoder kernel (TACO_DEC_TRACE) for a few geometries."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_inputs
from tacotron_multispeaker_b200.engine import Engine
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init

def impl_mma():
    return os.environ.get("TACO_DEC_IMPL", "mma") == "mma"

def run(N, T_in, cs, S, tag, iters=40):
    hp = HParams(outputs_per_step=5, max_iters=iters)
    w = random_init(hp, 60, seed=1234)
    ids, lengths, spk = make_inputs(N, T_in, 60, 1, min_len=max(1, int(T_in * 0.6)), vocab=(7108, 7325))
    os.environ["TACO_DEC_CS"] = str(cs); os.environ["TACO_DEC_S"] = str(S)
    eng = Engine(hp, 60); eng.load_weights(w); eng.set_profiling(True)
    mem = eng.encoder(ids, lengths, spk, 0)
    for _ in range(2):
        eng.decode(mem, None, False, True)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(3):
        eng.decode(mem, None, False, True)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 3
    path = os.path.join(ROOT, "gpurun_out", "trace_%s.txt" % tag)
    os.environ["TACO_DEC_TRACE"] = path
    eng.decode(mem, None, False, True)
    torch.cuda.synchronize()
    os.environ.pop("TACO_DEC_TRACE")
    st = [int(l.split()[1]) for l in open(path) if not l.startswith("#")]
    idx = [i for i in range(40) if st[i]]
    d = ["%d:%d" % (idx[k], st[idx[k]] - st[idx[k - 1]]) for k in range(1, len(idx))]
    print("%s N=%d CS=%d S=%d: %.2f us/step; traced step %d clk; stamp:delta %s" % (tag, N, cs, S, ms * 1e3 / iters, st[idx[-1]] - st[idx[0]], " ".join(d)), flush=True)
    if impl_mma():
        names = {160: "P8 sent, before TLOADP(P9)", 144: "TLOADP(P9) issued", 64: "TWAIT(P9) done", 80: "PRE(P9) done",
                 192: "P3 start (after wait P2)", 208: "P3 mma done", 224: "P3 reduce done",
                 96: "P10 start (after wait P9)", 112: "P10 mma done", 128: "P10 reduce/loads done",
                 256: "P6 start", 272: "P6 compute done", 288: "P6 sent + RTAKE(P8)", 304: "P7 start", 320: "P7 compute done"}
        stamps = [int(l.split()[1]) for l in open(path) if not l.startswith("#")]
        for b0 in (160, 144, 64, 80, 192, 208, 224, 96, 112, 128, 256, 272, 288, 304, 320):
            ref = {160: 22, 144: 22, 64: 22, 80: 22, 192: 6, 208: 6, 224: 6, 96: 27, 112: 27, 128: 27, 256: 15, 272: 15, 288: 15, 304: 18, 320: 18}[b0]
            print("   per-warp %-28s rel. stamp %d: %s" % (names[b0], ref, [stamps[b0 + w] - stamps[ref] if stamps[b0 + w] else None for w in range(16)]), flush=True)
    eng.close()

if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    impl = os.environ.get("TACO_DEC_IMPL", "mma")
    if impl == "mma":
        pass
        ctas = [int(c) for c in os.environ.get("TRACE_CTAS", "0").split(",")]
        for cta in ctas:
            os.environ["TACO_DEC_TRACE_CTA"] = str(cta)
            for (N, T_in, S) in ([(1, 50, 1), (32, 100, 5), (32, 100, 8)] if cta == 0 else [(32, 100, 5)]):
                run(N, T_in, 16, S, "mma_n%d_s%d_cta%d" % (N, S, cta), iters=200)
    else:
        run(1, 50, 16, 1, "n1_cs16_s1")
        run(4, 100, 16, 4, "n4_cs16_s4")
        run(4, 100, 8, 4, "n4_cs8_s4")
        run(32, 100, 8, 4, "n32_cs8_s4")
        run(28, 100, 16, 4, "n28_cs16_s4")
