"""Data-parallel training on real GPUs (needs >= 2 devices; skipped on a single-GPU box):
2 ranks x 1 clip over NCCL must produce the same loss, gradient norm and updated parameters as
1 rank x 2 clips - with and without overlapping the bucketed all-reduce with backward."""
import importlib
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model as om
from tests.helpers import SMALL_ARCH, make_inputs, small_cfg

pytestmark = pytest.mark.gpu


def _setup(seed=9, B=2):
    spa = importlib.import_module("3dspa_code_b200")
    c = small_cfg()
    model = spa.TrackAutoEncoder3D(**{k: getattr(c, k) for k in om.Config3D.__dataclass_fields__})
    inp, noise = make_inputs(c, B=B, N=9, Q=5, targets=True, seed=seed)
    tree = model.init(seed, inp, arch=SMALL_ARCH)["params"]
    om._randomize(tree, np.random.RandomState(seed))
    return spa, model, tree, inp, noise


def _worker(rank, world, port, overlap, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    te = importlib.import_module("3dspa_code_b200.train_engine")
    dp = importlib.import_module("3dspa_code_b200.dp")
    spa, model, tree, inp, noise = _setup()
    lo, hi = dp.shard_range(2, world, rank)
    mine = {k: v[lo:hi] for k, v in inp.items()}
    tr = te.Trainer(model, tree, precision="fp32", device=f"cuda:{rank}", base_lr=1e-3, warmup_steps=1, micro_batch=1, bucket_mb=1)
    tr.bucket_elems = 4096   # many buckets, so the overlap path really interleaves with backward
    tr.overlap = overlap
    logs = [tr.train_step(mine, noise[lo:hi]) for _ in range(2)]
    flat = tr.store.flat.detach().cpu().numpy().copy()
    q.put((rank, logs, flat))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_two_rank_dp_equals_single_rank(overlap):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    te = importlib.import_module("3dspa_code_b200.train_engine")
    spa, model, tree, inp, noise = _setup()
    ref = te.Trainer(model, tree, precision="fp32", device="cuda:0", base_lr=1e-3, warmup_steps=1, micro_batch=1)
    ref_logs = [ref.train_step(inp, noise) for _ in range(2)]
    ref_flat = ref.store.flat.detach().cpu().numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29731 + int(overlap)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, logs, flat in outs:
        for a, b in zip(logs, ref_logs):
            assert abs(a["total_loss"] - b["total_loss"]) < 1e-5 * abs(b["total_loss"])
            assert abs(a["grad_norm"] - b["grad_norm"]) < 1e-4 * b["grad_norm"]
        assert np.abs(flat - ref_flat).max() < 1e-6 + 1e-4 * np.abs(ref_flat).max()
    assert np.array_equal(outs[0][2], outs[1][2])   # replicas stay bit-identical
