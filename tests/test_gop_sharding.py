"""CPU, world_size 2 over gloo: GOPs sharded across ranks reconstruct to the same per-frame digests
as a single rank (the multi-GPU path of bench.py with the oracle standing in for the GPU engine)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from common import *  # noqa: F401,F403
from videodecoder_b200 import sharding, workload

N_GOPS, GOP_LEN, W, H = 5, 3, 4, 3


def _gop_digests(g):
    rng = np.random.default_rng(100 + g)
    gop = workload.make_gop(rng, W, H, GOP_LEN, workload.WorkloadConfig(compat=False, p_t8x8=0.3, p_i8x8=0.3))
    outs = run_sequence(oracle.recon_picture, W, H, 0, gop)
    return np.stack([np.frombuffer(oracle.digest(o["y"], o["u"], o["v"]), dtype=np.uint8) for o in outs])


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.assign_gops(N_GOPS, world)[rank]
    local = np.concatenate([_gop_digests(g) for g in mine]) if mine else np.zeros((0, 16), np.uint8)
    out = sharding.gather_digests(local, mine, N_GOPS, GOP_LEN, rank, world)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_assign_gops():
    assert sharding.assign_gops(5, 2) == [[0, 2, 4], [1, 3]]
    assert sharding.assign_gops(3, 8)[3:] == [[]] * 5
    got = sorted(g for r in sharding.assign_gops(20, 8) for g in r)
    assert got == list(range(20))


def test_two_ranks_match_single_rank(built):
    single = np.concatenate([_gop_digests(g) for g in range(N_GOPS)])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(out, single)
