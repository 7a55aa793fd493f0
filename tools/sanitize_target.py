"""Small shapes for compute-sanitizer (memcheck / racecheck / synccheck): the persistent cluster decoder, the BiGRU recurrence,
the whole forward and a short Griffin-Lim.  usage: compute-sanitizer --tool racecheck python tools/sanitize_target.py [what]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_inputs
from tacotron_multispeaker_b200.engine import Engine
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init

what = sys.argv[1] if len(sys.argv) > 1 else "all"
hp = HParams(outputs_per_step=5, max_iters=3)
eng = Engine(hp, 6)
eng.load_weights(random_init(hp, 6, seed=2))
rng = np.random.default_rng(0)
if what in ("all", "decoder"):
    mem = (rng.standard_normal((3, 19, 256)) * 0.5).astype(np.float32)
    dec, al, steps = eng.decode(mem, None, False)                      # free running, 3 utterances -> one cluster of 16 CTAs
    tg = rng.uniform(0, 1, (3, 15, 80)).astype(np.float32)
    dec2, al2, steps2 = eng.decode(mem, tg, True)                      # teacher forced
    torch.cuda.synchronize()
    print("decoder ok", steps, steps2, float(dec.abs().max()), float(dec2.abs().max()))
if what in ("all", "bigru"):
    x = (rng.standard_normal((3, 13, 128)) * 0.3).astype(np.float32)
    out = eng.bigru(0, x, np.array([13, 7, 1], np.int32))
    torch.cuda.synchronize()
    print("bigru ok", float(out.abs().max()))
if what in ("all", "forward"):
    ids, lengths, spk = make_inputs(2, 11, 6, 3)
    mel, lin, al, s = eng.forward(ids, lengths, spk)
    torch.cuda.synchronize()
    print("forward ok", s, float(lin.abs().max()))
if what in ("all", "vocoder"):
    lin = torch.rand(1, 6, hp.num_freq, device="cuda")
    wav = eng.griffin_lim(lin, 2)
    torch.cuda.synchronize()
    print("vocoder ok", float(wav.abs().max()))
eng.close()
