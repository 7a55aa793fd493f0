"""`gather_outputs` over NCCL on real GPUs (the only collective of the path: SURVEY.md §8e).  Needs two GPUs on one box;
skipped otherwise (the world-size-2 gloo test in tests/test_host_logic.py covers the host logic on CPU)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
from conftest import make_inputs
from tacotron_multispeaker_b200 import sharding
from tacotron_multispeaker_b200.engine import Engine
from tacotron_multispeaker_b200.hparams import HParams
from tacotron_multispeaker_b200.weights import random_init
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hp = HParams(outputs_per_step=5, max_iters=7)
w = random_init(hp, 6, seed=3)
N = 5                                   # uneven split: 3 + 2
ids, lengths, spk = make_inputs(N, 21, 6, 9)
lo, hi = sharding.shard_bounds(N, world, rank)
eng = Engine(hp, 6, local); eng.load_weights(w)
mel, lin, al, steps = eng.forward(ids[lo:hi], lengths[lo:hi], spk[lo:hi])
full = [sharding.gather_outputs(t.contiguous(), N, dst=0) for t in (mel, lin, al)]
allg = sharding.gather_outputs(mel.contiguous(), N)          # all_gather flavour: every rank gets everything
assert allg.shape[0] == N
if rank == 0:
    ref = Engine(hp, 6, local); ref.load_weights(w)
    rmel, rlin, ral, rsteps = ref.forward(ids, lengths, spk)
    assert steps == rsteps
    for got, want in zip(full, (rmel, rlin, ral)):
        assert got.shape == want.shape, (got.shape, want.shape)
        assert float((got - want).abs().max()) < 2e-5, float((got - want).abs().max())
    assert float((allg - rmel).abs().max()) < 2e-5
    print("GATHER_OK", tuple(full[1].shape))
dist.barrier()
dist.destroy_process_group()
'''


def test_gather_outputs_nccl_two_ranks(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box (run under `gpurun --gpus 2`)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "GATHER_OK" in r.stdout
