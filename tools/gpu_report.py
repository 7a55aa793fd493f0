"""Developer aid: per-stage max-abs error of the CUDA path against the oracle at a
small shape, then CUDA-event timings of the stages at the benchmark shape.
Writes everything to stdout; run under gpurun and redirect into gpurun_out/.
"""
import os
import sys
import time
import traceback
os.environ.setdefault("TACO_DEV", "1")   # per-call developer switches of the C ABI

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import make_inputs  # noqa: E402
from oracle import taco_oracle as O  # noqa: E402
from tacotron_multispeaker_b200.engine import Engine  # noqa: E402
from tacotron_multispeaker_b200.hparams import HParams  # noqa: E402
from tacotron_multispeaker_b200.weights import random_init  # noqa: E402


def err(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double() if isinstance(b, torch.Tensor) else torch.as_tensor(b).double()
    if a.shape != b.shape:
        return "SHAPE %s vs %s" % (tuple(a.shape), tuple(b.shape))
    d = (a - b).abs()
    return "max|d|=%.3e  max|ref|=%.3e  nan=%d" % (float(d.max()), float(b.abs().max()), int(torch.isnan(a).sum()))


def stage(name, fn):
    try:
        t0 = time.time()
        msg = fn()
        torch.cuda.synchronize()
        print("[%-28s] %s  (%.2fs)" % (name, msg, time.time() - t0), flush=True)
    except Exception:
        print("[%-28s] EXCEPTION\n%s" % (name, traceback.format_exc()), flush=True)


def small_checks():
    hp = HParams(outputs_per_step=5, max_iters=8)
    w = random_init(hp, 6, seed=7, randomize_bn=True)
    ow = O.W(w)
    eng = Engine(hp, 6)
    eng.load_weights(w)
    print("geometry N=4:", eng.decoder_geometry(4), "N=32:", eng.decoder_geometry(32), flush=True)
    ids, lengths, spk = make_inputs(4, 21, 6, 3)
    rng = np.random.default_rng(0)
    stage("embed", lambda: err(eng.embed(ids, spk), O.embed(ids, spk, ow)))
    x = rng.standard_normal((3, 37, 128)).astype(np.float32)
    wk = (rng.standard_normal((4, 128, 128)) / 20).astype(np.float32)
    b = rng.standard_normal((128,)).astype(np.float32)
    stage("conv1d k=4", lambda: err(eng.conv1d(x, wk, b, 1),
                                    torch.relu(O.conv1d_same(torch.from_numpy(x), torch.from_numpy(wk), torch.from_numpy(b)))))
    xl = np.array([37, 20, 3], np.int32)
    stage("bigru enc masked", lambda: err(eng.bigru(0, x, xl), O.bigru(torch.from_numpy(x), xl, ow, "encoder_cbhg")))
    stage("bigru post", lambda: err(eng.bigru(1, x, None), O.bigru(torch.from_numpy(x), None, ow, "post_cbhg")))
    stage("cbhg enc moving", lambda: err(eng.cbhg(0, x, xl, 0), O.cbhg(torch.from_numpy(x), xl, ow, "encoder_cbhg", 16, "moving")))
    stage("cbhg enc batch", lambda: err(eng.cbhg(0, x, xl, 1), O.cbhg(torch.from_numpy(x), xl, ow, "encoder_cbhg", 16, "batch")))
    xm = rng.uniform(-0.5, 1, (3, 37, 80)).astype(np.float32)
    stage("cbhg post moving", lambda: err(eng.cbhg(1, xm, None, 0), O.cbhg(torch.from_numpy(xm), None, ow, "post_cbhg", 8, "moving")))
    stage("encoder", lambda: err(eng.encoder(ids, lengths, spk, 0), O.encoder(ids, lengths, spk, ow, "moving")))
    mem = (rng.standard_normal((4, 21, 256)) * 0.5).astype(np.float32)
    tg = rng.uniform(0, 1, (4, 40, 80)).astype(np.float32)

    def dec_case(teacher, S=None):
        def f():
            if S:
                os.environ["TACO_DEC_S"] = str(S)
            try:
                d, a, s = eng.decode(mem, tg if teacher else None, teacher)
            finally:
                os.environ.pop("TACO_DEC_S", None)
            rd, ra, rs = O.decode(torch.from_numpy(mem), ow, 80, 5, 8, torch.from_numpy(tg) if teacher else None, teacher)
            return "steps %d/%d dec: %s | align: %s" % (s, rs, err(d, rd), err(a, ra))
        return f
    stage("decode teacher", dec_case(True))
    stage("decode teacher S=2", dec_case(True, 2))
    stage("decode teacher S=4", dec_case(True, 4))
    stage("decode teacher S=8", dec_case(True, 8))
    stage("decode free", dec_case(False))

    def whole():
        ref = O.tacotron_forward(w, hp, ids, lengths, identities=spk, id_num=6)
        mel, lin, al, s = eng.forward(ids, lengths, spk)
        return "steps %d/%d mel: %s | lin: %s | al: %s" % (s, ref["steps"], err(mel, ref["mel_outputs"]),
                                                          err(lin, ref["linear_outputs"]), err(al, ref["alignments"]))
    stage("forward free", whole)
    eng.close()


def timings(N=32, T_in=100, iters=200):
    hp = HParams(outputs_per_step=5, max_iters=iters)
    w = random_init(hp, 60, seed=1234)
    eng = Engine(hp, 60)
    eng.load_weights(w)
    eng.set_profiling(True)
    ids, lengths, spk = make_inputs(N, T_in, 60, 1, min_len=max(1, int(T_in * 0.6)), vocab=(7108, 7325))
    print("decoder geometry:", eng.decoder_geometry(N), flush=True)
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.time()
        mel, lin, al, s = eng.forward(ids, lengths, spk)
        torch.cuda.synchronize()
        dt = time.time() - t0
        ms = eng.last_stage_ms()
        print("iter %d: wall %.2f ms  stages(ms) enc %.3f dec %.3f post %.3f  -> %.0f frames/s; dec us/step %.2f"
              % (i, dt * 1e3, ms["encoder"], ms["decoder"], ms["postnet"], N * s * 5 / dt, ms["decoder"] * 1e3 / s),
              flush=True)
    print("finite:", bool(torch.isfinite(mel).all()), bool(torch.isfinite(lin).all()), "launches", eng.launch_count())
    eng.close()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    small_checks()
    if "--no-timing" not in sys.argv:
        stage("timings N=32", lambda: timings() or "ok")
        stage("timings N=1", lambda: timings(1, 50) or "ok")
