"""Extra legs of bench.py: the other configurations BASELINE.json names, reported as extra keys of the one
JSON line (the headline `value` stays config 3).

 * config1  single-speaker, batch 1, 50 symbols, r=5, max_iters=200, free running (reference hparams defaults shape)
 * config2  multispeaker teacher-forced forward, batch 32 x 1000 frames, BN in `batch` (reference-faithful, models/tacotron.py:36)
            and `moving` mode
 * config4  AISHELL-shaped synthesis, GLOBAL batch 256 sharded by utterance over the ranks (256 / 128 / 64 / 32 per GPU),
            T_in <= 60, id_num = 400 (preprocess_data.py:59-64), hanzi ids; strong scaling; plus one timed NCCL gather of the outputs
 * config5  batch-1 latency sweep over 20..200 symbols, >= 50 runs per length, p50 and p99, for (r=5, max_iters=200) and for the
            fork's own Synthesizer setting (r=1, max_iters=400; synthesizer.py:21, hparams.py:22)
 * e2e_synthesize  what the reference's Synthesizer.synthesize fetches (synthesizer.py:47): ids -> forward -> Griffin-Lim ->
            D2H of the WAVEFORM and the ALIGNMENT only (the linear spectrogram never leaves the device)
"""
from __future__ import annotations

import statistics
import threading
import time

import numpy as np


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _pct(xs, p):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(round(p / 100.0 * (len(xs) - 1))))]


def run(torch, Engine, HParams, random_init, _abi, sharding, dev, local, rank, world, dist, barrier, headline_engine, hp, id_num,
        steps: int, quick: bool = False):
    out = {}
    R, MAX_ITERS = hp.outputs_per_step, hp.max_iters
    T_out = R * MAX_ITERS

    side = torch.cuda.Stream(device=dev)     # a non-default stream: the C ABI replays its CUDA graph of the forward there

    def time_forward(eng, n_runs, fn):
        ts = []
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(n_runs + 3):
                a0, a1 = _events(torch)
                a0.record()
                fn()
                a1.record()
                a1.synchronize()
                if i >= 3:
                    ts.append(a0.elapsed_time(a1))
        torch.cuda.current_stream().wait_stream(side)
        return ts

    # ---------------------------------------------------------------- config 4: global batch 256, sharded, strong scaling
    G = 256
    hp4 = HParams(outputs_per_step=R, max_iters=MAX_ITERS)
    rng = np.random.default_rng(4)
    len4 = rng.integers(20, 61, (G,)).astype(np.int32)
    ids4 = rng.integers(2, 7054, (G, 60)).astype(np.int32)            # hanzi block of symbols2
    for i in range(G):
        ids4[i, len4[i]:] = 0
    spk4 = rng.integers(0, 400, (G,)).astype(np.int32)
    lo, hi = sharding.shard_bounds(G, world, rank)
    e4 = Engine(hp4, 400, local)
    e4.load_weights(random_init(hp4, 400, seed=1234))
    ids_l, len_l, spk_l = (torch.from_numpy(x[lo:hi]).to(dev) for x in (ids4, len4, spk4))
    n_l = hi - lo
    outs4 = (torch.zeros(n_l, T_out, hp4.num_mels, device=dev), torch.zeros(n_l, T_out, hp4.num_freq, device=dev),
             torch.zeros(n_l, 60, MAX_ITERS, device=dev))
    for _ in range(3):
        e4.forward(ids_l, len_l, spk_l, out=outs4)
    barrier()
    k4 = 3 if quick else max(3, min(steps, 8))
    a0, a1 = _events(torch)
    a0.record()
    for _ in range(k4):
        st4 = e4.forward(ids_l, len_l, spk_l, out=outs4)[3]
    a1.record()
    barrier()
    ms4 = a0.elapsed_time(a1) / k4
    gather_ms = None
    if dist is not None:
        tt = torch.tensor([ms4], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms4 = float(tt.item())
        # NCCL gather of the outputs to rank 0 (the only collective of the path; off the hot loop)
        for t in (outs4[0],):
            sharding.gather_outputs(t, G, dst=0)
        barrier()
        g0, g1 = _events(torch)
        g0.record()
        mel_all = sharding.gather_outputs(outs4[0], G, dst=0)
        lin_all = sharding.gather_outputs(outs4[1], G, dst=0)
        al_all = sharding.gather_outputs(outs4[2], G, dst=0)
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1)
        tt = torch.tensor([gather_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        gather_ms = float(tt.item())
        if rank == 0:
            assert mel_all.shape[0] == G and lin_all.shape[0] == G and al_all.shape[0] == G
        del mel_all, lin_all, al_all
    out["config4"] = {
        "workload": "AISHELL-shaped multispeaker free-running synthesis, GLOBAL batch 256 sharded by utterance (%d per GPU), T_in<=60 "
                    "(lengths U{20..60}), id_num=400, hanzi ids, max_iters=%d, r=%d" % (-(-G // world), MAX_ITERS, R),
        "scaling": "strong", "global_batch": G, "per_gpu_batch": n_l, "ms_per_batch": ms4,
        "value": G * st4 * R / (ms4 / 1e3), "unit": "frames/s",
        "gather_outputs_ms": gather_ms,
        "gather_bytes": int(G * T_out * (hp4.num_mels + hp4.num_freq) * 4 + G * 60 * MAX_ITERS * 4) if dist is not None else 0,
        "decoder_geometry": e4.decoder_geometry(n_l),
    }
    e4.close()
    del outs4
    torch.cuda.empty_cache()

    # ---------------------------------------------------------------- e2e shaped like Synthesizer.synthesize: wav + alignment out
    B, T_IN = 32, 100
    rng = np.random.default_rng(77 + rank)
    len_s = rng.integers(60, T_IN + 1, (B,)).astype(np.int32)
    ids_s = rng.integers(7108, 7325, (B, T_IN)).astype(np.int32)
    for i in range(B):
        ids_s[i, len_s[i]:] = 0
    spk_s = rng.integers(0, id_num, (B,)).astype(np.int32)

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype).pin_memory()
    ap = headline_engine.audio_params()
    wav_len = int(headline_engine.lib.taco_wav_length(__import__("ctypes").byref(ap), T_out))
    n_lanes = 2
    lanes = []
    for li in range(n_lanes):
        e = headline_engine if li == 0 else Engine(hp, id_num, local)
        if li > 0:
            e.load_weights(random_init(hp, id_num, seed=1234))
        lanes.append(dict(
            eng=e, stream=torch.cuda.Stream(device=dev),
            h_ids=pinned((B, T_IN), torch.int32), h_len=pinned((B,), torch.int32), h_spk=pinned((B,), torch.int32),
            d_ids=torch.empty(B, T_IN, dtype=torch.int32, device=dev), d_len=torch.empty(B, dtype=torch.int32, device=dev),
            d_spk=torch.empty(B, dtype=torch.int32, device=dev),
            outs=(torch.zeros(B, T_out, hp.num_mels, device=dev), torch.zeros(B, T_out, hp.num_freq, device=dev),
                  torch.zeros(B, T_IN, MAX_ITERS, device=dev)),
            h_wav=pinned((B, wav_len), torch.float32), h_al=pinned((B, T_IN, MAX_ITERS), torch.float32)))
        lanes[-1]["h_ids"].copy_(torch.from_numpy(ids_s)); lanes[-1]["h_len"].copy_(torch.from_numpy(len_s))
        lanes[-1]["h_spk"].copy_(torch.from_numpy(spk_s))

    def synth_step(l):
        # H2D of the ids, forward (linear stays on the device), Griffin-Lim + inverse pre-emphasis, D2H of wav + alignment
        l["d_ids"].copy_(l["h_ids"], non_blocking=True)
        l["d_len"].copy_(l["h_len"], non_blocking=True)
        l["d_spk"].copy_(l["h_spk"], non_blocking=True)
        _, lin, al, _ = l["eng"].forward(l["d_ids"], l["d_len"], l["d_spk"], out=l["outs"])
        if "d_wav" not in l:
            l["d_wav"] = torch.empty(lin.shape[0], l["h_wav"].shape[-1], device=lin.device, dtype=torch.float32)
        wav = l["eng"].griffin_lim(lin, out=l["d_wav"])
        l["h_wav"].copy_(wav, non_blocking=True)
        l["h_al"].copy_(al, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def synth_timed(n_steps):
        barrier()
        share = [n_steps // n_lanes + (1 if i < n_steps % n_lanes else 0) for i in range(n_lanes)]

        def work(l, k):
            with torch.cuda.stream(l["stream"]):
                for _ in range(k):
                    synth_step(l)
        t0 = time.perf_counter()
        ths = [threading.Thread(target=work, args=(l, k)) for l, k in zip(lanes, share) if k > 0]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([dt], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        return dt / n_steps
    for l in lanes:
        with torch.cuda.stream(l["stream"]):
            synth_step(l); synth_step(l)
    ks = 4 if quick else 8
    s_per = synth_timed(ks)
    out["e2e_synthesize"] = {
        "what": "reference Synthesizer.synthesize shape (synthesizer.py:47): pinned-host ids -> forward -> Griffin-Lim (%d iterations) + "
                "inverse pre-emphasis -> D2H of waveform and alignment only; %d batches in flight per GPU" % (hp.griffin_lim_iters, n_lanes),
        "value": B * world * T_out / s_per, "unit": "frames/s", "ms_per_step": 1e3 * s_per,
        "h2d_bytes_per_step": int(ids_s.nbytes + len_s.nbytes + spk_s.nbytes),
        "d2h_bytes_per_step": int(B * wav_len * 4 + B * T_IN * MAX_ITERS * 4),
        "audio_seconds_per_step": B * wav_len / hp.sample_rate,
    }
    for l in lanes[1:]:
        l["eng"].close()
    del lanes
    torch.cuda.empty_cache()

    if world > 1 or rank != 0:
        return out

    # ---------------------------------------------------------------- config 2: teacher forced, batch 32 x 1000 frames
    eng = headline_engine
    rng = np.random.default_rng(1)
    len2 = rng.integers(60, T_IN + 1, (B,)).astype(np.int32)
    ids2 = rng.integers(7108, 7325, (B, T_IN)).astype(np.int32)
    for i in range(B):
        ids2[i, len2[i]:] = 0
    spk2 = rng.integers(0, id_num, (B,)).astype(np.int32)
    tg2 = torch.from_numpy(rng.uniform(0, 1, (B, T_out, hp.num_mels)).astype(np.float32)).to(dev)
    d_ids, d_len, d_spk = (torch.from_numpy(x).to(dev) for x in (ids2, len2, spk2))
    outs = (torch.zeros(B, T_out, hp.num_mels, device=dev), torch.zeros(B, T_out, hp.num_freq, device=dev),
            torch.zeros(B, T_IN, MAX_ITERS, device=dev))
    c2 = {"workload": "multispeaker teacher-forced forward, batch 32, T_in<=100, 1000 mel frames (200 steps, r=5), mel targets U(0,1)",
          "unit": "frames/s"}
    for name, mode in (("batch", _abi.BN_BATCH), ("moving", _abi.BN_MOVING)):
        ts = time_forward(eng, 5 if quick else 10,
                          lambda: eng.forward(d_ids, d_len, d_spk, mel_targets=tg2, teacher_force=True, bn_mode=mode, out=outs))
        ms = statistics.median(ts)
        c2["bn_" + name] = {"ms_per_batch": ms, "value": B * T_out / (ms / 1e3)}
    out["config2"] = c2

    # ---------------------------------------------------------------- config 5: batch-1 latency sweep, (r=5, 200) and (r=1, 400)
    n_runs = 20 if quick else 50
    c5 = {"what": "one utterance, device resident, one stream, CUDA events around taco_forward (gather -> linear output); "
                  "%d runs per length after 3 warm-ups" % n_runs, "unit": "ms"}
    for tag, r5, iters5 in (("r5_iters200", R, MAX_ITERS), ("r1_iters400", 1, 400)):
        if (r5, iters5) == (R, MAX_ITERS):
            e5 = eng
        else:
            hp5 = HParams(outputs_per_step=r5, max_iters=iters5)
            e5 = Engine(hp5, id_num, local)
            e5.load_weights(random_init(hp5, id_num, seed=1234))
        T5 = r5 * iters5
        o5 = (torch.zeros(1, T5, hp.num_mels, device=dev), torch.zeros(1, T5, hp.num_freq, device=dev))
        p50, p99 = {}, {}
        for t_in in range(20, 201, 20):
            rng1 = np.random.default_rng(500 + t_in)
            ids1 = torch.from_numpy(rng1.integers(7108, 7325, (1, t_in)).astype(np.int32)).to(dev)
            len1 = torch.tensor([t_in], dtype=torch.int32, device=dev)
            spk1 = torch.tensor([3], dtype=torch.int32, device=dev)
            al1 = torch.zeros(1, t_in, iters5, device=dev)
            ts = time_forward(e5, n_runs, lambda: e5.forward(ids1, len1, spk1, out=(o5[0], o5[1], al1)))
            p50["T_in=%d" % t_in] = statistics.median(ts)
            p99["T_in=%d" % t_in] = _pct(ts, 99)
        c5[tag] = {"p50": p50, "p99": p99, "mel_frames": T5}
        if e5 is not eng:
            e5.close()
    out["config5"] = c5

    # ---------------------------------------------------------------- config 1: single speaker, batch 1, 50 symbols
    hp1 = HParams(outputs_per_step=R, max_iters=MAX_ITERS)
    e1 = Engine(hp1, 0, local)
    e1.load_weights(random_init(hp1, 0, seed=1234))
    rng = np.random.default_rng(0)
    ids1 = torch.from_numpy(rng.integers(2, 7352, (1, 50)).astype(np.int32)).to(dev)
    len1 = torch.tensor([50], dtype=torch.int32, device=dev)
    o1 = (torch.zeros(1, T_out, hp.num_mels, device=dev), torch.zeros(1, T_out, hp.num_freq, device=dev),
          torch.zeros(1, 50, MAX_ITERS, device=dev))
    ts = time_forward(e1, n_runs, lambda: e1.forward(ids1, len1, None, out=o1))
    out["config1"] = {"workload": "single-speaker, batch 1, 50 symbols, r=5, max_iters=200, free running",
                      "p50_ms": statistics.median(ts), "p99_ms": _pct(ts, 99),
                      "value": T_out / (statistics.median(ts) / 1e3), "unit": "frames/s"}
    e1.close()
    return out
