"""CPU, world_size 2 over gloo: the N>1 host logic — group-range sharding with a replicated node and
group table, one all-gather of the per-rank admit bitmaps — reproduces the unsharded round exactly.
(The per-rank round is computed by the CPU oracle here; on the GPU box the same sharding feeds the
CUDA engine, see tests/test_gpu_multi.py.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, seed, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    from oracle import oracle
    from randsnap import random_snapshot
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = importlib.import_module("batch-scheduler_b200.snapshot")
    full = (S.config(2, 0.15) if seed < 0 else random_snapshot(seed, P=400, N=60, G=40, L=5)).resolve_groups()
    local = full.shard_groups(rank, world)
    g0, g1 = local.meta["group_range"]
    r = oracle.round(local, want_bitmap=False)
    # this rank's slice of the verdicts; groups outside the range carry no pods here
    G = full.groups.n
    mine = np.zeros(G, np.uint8)
    mine[g0:g1] = r.admit[g0:g1] + 1           # 0 = not mine
    t = torch.from_numpy(mine)
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    merged = np.zeros(G, np.uint8)
    for p in parts:
        a = p.numpy()
        assert not ((merged > 0) & (a > 0)).any()   # ranges are disjoint
        merged = np.maximum(merged, a)
    assert (merged > 0).all()
    # per-pod outputs, scattered back to global pod indices
    idx = local.meta["pod_index"]
    pf = torch.zeros(full.pods.n, dtype=torch.int32)
    pf[torch.from_numpy(idx)] = torch.from_numpy(r.prefilter.astype(np.int32)) + 1
    dist.all_reduce(pf, op=dist.ReduceOp.SUM)
    if rank == 0:
        ref = oracle.round(full, want_bitmap=False)
        np.testing.assert_array_equal(merged - 1, ref.admit)
        np.testing.assert_array_equal(pf.numpy() - 1, ref.prefilter.astype(np.int32))
        assert r.max_group == ref.max_group
        open(os.path.join(out_dir, f"ok{seed}"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("seed", [-1, 3, 8])
def test_group_sharding_world2(tmp_path, seed):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, seed, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), f"ok{seed}"))


def test_resolve_groups_matches_round(oracle, snapshot_mod):
    # resolving first pods on the host must not change the round
    from randsnap import random_snapshot
    for seed in range(6):
        snap = random_snapshot(seed, P=200, N=30, G=15, L=6)
        a = oracle.round(snap, want_bitmap=False)
        b = oracle.round(snap.resolve_groups(), want_bitmap=False)
        for f in ("prefilter", "admit", "new_denied", "order"):
            np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f)
        assert a.max_group == b.max_group
