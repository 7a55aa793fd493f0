"""Host-side logic that needs no GPU: hparams, variable inventory, checkpoint-name
canonicalisation, the text front-end, utterance sharding (world_size-2 gloo)."""
import os
import socket

import numpy as np
import pytest
import torch

from tacotron_multispeaker_b200.hparams import HParams, hparams_debug_string
from tacotron_multispeaker_b200.sharding import gather_outputs, shard_batch, shard_bounds
from tacotron_multispeaker_b200.weights import PREFIX, canonicalize, count_params, random_init, weight_specs


def test_hparams_defaults_and_parse():
    hp = HParams()
    # fork defaults (reference hparams.py:11-12,22,34,39-40)
    assert (hp.num_mels, hp.num_freq, hp.outputs_per_step, hp.max_iters) == (80, 1025, 1, 2000)
    assert (hp.embedding_text_channels, hp.embedding_id_channels, hp.sample_rate) == (256, 64, 20000)
    hp.parse("outputs_per_step=5, max_iters=200,preemphasis=0.9,eos=false")
    assert hp.outputs_per_step == 5 and hp.max_iters == 200 and hp.preemphasis == 0.9 and hp.eos is False
    with pytest.raises(ValueError):
        hp.parse("no_such_param=1")
    assert "max_iters: 200" in hparams_debug_string(hp)


def test_weight_inventory_shapes_and_count():
    hp = HParams(outputs_per_step=5)
    s = weight_specs(hp, 60)
    assert s["embedding"][0] == (7352, 256) and s["embedding_id"][0] == (60, 64)
    assert s["prenet/dense_1/kernel"][0] == (320, 256)
    assert s["encoder_cbhg/conv_bank/conv1d_16/conv1d/kernel"][0] == (16, 128, 128)
    assert s["encoder_cbhg/proj_1/conv1d/kernel"][0] == (3, 2048, 128)
    assert s["post_cbhg/proj_1/conv1d/kernel"][0] == (3, 1024, 256)
    assert s["post_cbhg/proj_2/conv1d/kernel"][0] == (3, 256, 80)
    assert s["post_cbhg/dense/kernel"][0] == (80, 128) and "encoder_cbhg/dense/kernel" not in s
    assert s["decoder/output_projection_wrapper/kernel"][0] == (256, 400)
    assert s["dense/kernel"][0] == (256, 1025)
    gru = [k for k in s if k.endswith("decoder_prenet_wrapper/gru_cell/gates/kernel")]
    assert len(gru) == 1 and s[gru[0]][0] == (384, 512)
    # 8.80 M + id_num*64 parameters (BN moving statistics included here)
    assert abs(count_params(hp, 60) - 8.80e6) < 0.02e6
    single = weight_specs(hp, 0)
    assert "embedding_id" not in single and single["prenet/dense_1/kernel"][0] == (256, 256)


def test_random_init_follows_reference_initializers():
    hp = HParams()
    w = random_init(hp, 3, seed=0)
    e = w[PREFIX + "embedding"]
    assert np.abs(e).max() <= 1.0 + 1e-6 and 0.40 < e.std() < 0.47       # truncated normal, sigma 0.5 cut at 2 sigma
    assert np.all(w[PREFIX + "encoder_cbhg/highway_2/T/bias"] == -1.0)
    assert np.all(w[PREFIX + "encoder_cbhg/bidirectional_rnn/fw/gru_cell/gates/bias"] == 1.0)
    k = w[PREFIX + "prenet/dense_1/kernel"]
    assert np.abs(k).max() <= np.sqrt(6.0 / (320 + 256)) + 1e-7           # glorot uniform
    assert np.all(w[PREFIX + "post_cbhg/proj_1/batch_normalization/moving_variance"] == 1.0)
    w2 = random_init(hp, 3, seed=0)
    assert all(np.array_equal(w[n], w2[n]) for n in w)


def test_canonicalize_accepts_checkpoint_style_names():
    hp = HParams()
    w = random_init(hp, 2, seed=1)
    ck = dict(w)
    ck["global_step"] = np.array(7)
    ck[PREFIX + "embedding/Adam"] = np.zeros((7352, 256), np.float32)     # optimizer slots are ignored
    # a checkpoint written under a different outer scope still matches by suffix
    ck["tower_0/inference/memory_layer/kernel"] = ck.pop(PREFIX + "memory_layer/kernel")
    c = canonicalize(ck, hp, 2)
    assert set(c) == set(weight_specs(hp, 2))
    del ck[PREFIX + "dense/bias"]
    with pytest.raises(KeyError):
        canonicalize(ck, hp, 2)
    bad = dict(w); bad[PREFIX + "dense/bias"] = np.zeros(3, np.float32)
    with pytest.raises(ValueError):
        canonicalize(bad, hp, 2)


def test_text_front_end_reference_quirks():
    from tacotron_multispeaker_b200 import text
    chars = ["一", "二", "a", "b", " ", "sh", "ang4", "~", " "]            # duplicates: '~' and ' '
    text.load_symbols(chars)
    assert len(text.symbols2) == len(chars) + 2
    seq = text.text_to_sequence2("一x二{sh ang4}a", ["basic_cleaners"])
    # unknown 'x' dropped; braces split on spaces into multi-char symbols; EOS id 1 appended
    assert seq == [2, 3, 7, 8, 4, 1]
    assert text.text_to_sequence2(" ", [])[0] == 10                       # later duplicate wins
    assert text.sequence_to_text2(seq[:-1]) == "一二shang4a"
    assert text.text_to_sequence2("", []) == [1]


def test_synthesizer_load_requires_embedding_id(tmp_path):
    """single-speaker checkpoints make load() raise KeyError like reference synthesizer.py:25"""
    from tacotron_multispeaker_b200.synthesizer import Synthesizer
    hp = HParams()
    w = random_init(hp, 0, seed=0)
    p = tmp_path / "ckpt.npz"
    np.savez(p, **{k: v for k, v in w.items() if "conv_bank" not in k})
    with pytest.raises(KeyError):
        Synthesizer(hp).load(str(p))


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    ids = np.arange(10).reshape(5, 2)
    a, b = shard_batch([ids, None], 2, 1)
    assert b is None and np.array_equal(a, ids[3:])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_total * 3, dtype=torch.float32).reshape(n_total, 3)
    lo, hi = shard_bounds(n_total, world, rank)
    local = full[lo:hi] * 2.0                      # the "forward" of this rank's utterances
    out = gather_outputs(local, n_total)
    only0 = gather_outputs(local, n_total, dst=0)
    ok = torch.equal(out, full * 2.0) and ((only0 is None) if rank else torch.equal(only0, full * 2.0))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 8])
def test_output_gather_world_size_2_gloo(n_total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]




def test_canonicalize_reports_ambiguous_suffix_matches():
    """Two checkpoint variables ending in the same expected name are not skipped silently: the KeyError names them."""
    from tacotron_multispeaker_b200.hparams import HParams
    from tacotron_multispeaker_b200 import weights as W
    hp = HParams()
    full = W.random_init(hp, 4, seed=0)
    name = W.PREFIX + "embedding"
    arr = full.pop(name)
    full["tower_0/" + name[len(W.PREFIX):]] = arr
    full["tower_1/" + name[len(W.PREFIX):]] = arr
    with pytest.raises(KeyError) as ei:
        W.canonicalize(full, hp, 4)
    assert "ambiguous suffix matches" in str(ei.value) and "tower_0/" in str(ei.value)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's reference arm: the CPU restatement of the reference graph on the host cores)
    prints ONE JSON line with the contract's keys, the same metric / unit / config as the GPU arm, and needs no GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mel_frames_per_s" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "config 3" in d["config"]["workload"] and d["config"]["global_batch"] == 32
