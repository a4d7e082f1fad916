"""GPU, 2 ranks: group-sharded evaluation on the CUDA engine + the all-gather of the admit bitmap over
peer memory (CUDA IPC) reproduces the unsharded oracle round.

On a box with >= 2 GPUs every rank owns a GPU and the same all-gather through NCCL is the cross-check.
On a ONE-GPU box (the driver's GPU-test box) both ranks share cuda:0: CUDA IPC maps a buffer of another
process on the same device just as well, the two contexts time-slice, and the exchange protocol
(push / flags / two slot sets / side-stream wait) is exercised end to end — only NCCL is left out
(it refuses two ranks on one device); the handles travel over gloo."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(rank, world, port, shared_gpu):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = 0 if shared_gpu else rank
    torch.cuda.set_device(dev)
    if shared_gpu:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    pkg = importlib.import_module("batch-scheduler_b200")
    return pkg, torch, dist, dev


def _ag_factory(dist, world):
    def _ag(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    return _ag


def _worker(rank, world, port, out_dir, shared_gpu):
    pkg, torch, dist, dev = _setup(rank, world, port, shared_gpu)
    S = pkg.snapshot
    full = S.config(3, 0.1).resolve_groups()
    local = full.shard_groups(rank, world)
    g0, g1 = local.meta["group_range"]
    G = full.groups.n
    eng = pkg.Engine(local.lanes, dev, fit_bitmap=False, score=False)
    eng.upload(local)
    words_nccl = None
    if not shared_gpu:
        eng.evaluate_async()
        # the admit bitmap straight from the engine's device buffer, gathered on the engine's stream
        ptr, nbytes = eng.device_buffer(pkg.capi.BUF_ADMIT_BITMAP)

        class H:
            pass
        h = H()
        h.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        mine = torch.as_tensor(h, device=f"cuda:{dev}")
        gathered = torch.empty(world * mine.numel(), dtype=torch.int32, device=f"cuda:{dev}")
        ext = torch.cuda.ExternalStream(eng.stream(), device=dev)
        with torch.cuda.stream(ext):
            dist.all_gather_into_tensor(gathered, mine)
        eng.sync()
        torch.cuda.synchronize()
        words_nccl = gathered.cpu().numpy().view(np.uint32).reshape(world, -1)
    # the all-gather by the engine's own peer-memory kernels (CUDA IPC), several rounds back to back: both
    # slot sets are used and round k+1 is enqueued while round k's wait may still be pending
    eng.peer_setup(rank, world, (G + 31) // 32, _ag_factory(dist, world))
    dist.barrier()
    for _ in range(5):
        eng.evaluate_async()
    eng.sync()
    p2p = eng.gathered_admit()
    res = eng.fetch()
    if words_nccl is not None:
        np.testing.assert_array_equal(p2p, words_nccl[:, :p2p.shape[1]])
    # a different round right behind: every rank flips one carried-in matched count of its first group, so the
    # gathered words must follow the new round (no stale slot set)
    groups_b = local.groups.copy()             # (the shard shares the group table object with `full`)
    groups_b.matched[g0] = groups_b.min_member[g0]
    eng.upload_groups(groups_b)
    eng.evaluate_async()
    eng.sync()
    p2p_b = eng.gathered_admit()
    res_b = eng.fetch()
    own = np.unpackbits(p2p_b[rank].view(np.uint8), bitorder="little")[:G].astype(bool)
    np.testing.assert_array_equal(own[g0:g1], res_b.admit[g0:g1] == S.ADMIT)
    ranges = [None] * world
    dist.all_gather_object(ranges, (int(g0), int(g1)))
    merged = np.zeros(G, bool)
    for r, (a0, a1) in enumerate(ranges):
        bits = np.unpackbits(p2p[r].view(np.uint8), bitorder="little")[:G].astype(bool)
        merged[a0:a1] = bits[a0:a1]
    dist.barrier()
    eng.peer_detach()
    if rank == 0:
        from oracle import oracle
        ref = oracle.round(full, want_bitmap=False)
        np.testing.assert_array_equal(merged, ref.admit == S.ADMIT)
        np.testing.assert_array_equal(res.admit[g0:g1], ref.admit[g0:g1])
        idx = local.meta["pod_index"]
        np.testing.assert_array_equal(res.prefilter, ref.prefilter[idx])
        np.testing.assert_array_equal(res.feasible_count, ref.feasible_count[idx])
        assert res.max_group == ref.max_group
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


def _worker_timeout(rank, world, port, out_dir, shared_gpu):
    """A rank that never arrives: the other rank's round ends with BS_E_PEER after the bounded wait, later
    rounds fail fast, and a fresh attach (new epoch) works again."""
    os.environ["BS_PEER_TIMEOUT_MS"] = "300"
    pkg, torch, dist, dev = _setup(rank, world, port, shared_gpu)
    import time
    S = pkg.snapshot
    full = S.config(2, 0.2).resolve_groups()
    local = full.shard_groups(rank, world)
    G = full.groups.n
    eng = pkg.Engine(local.lanes, dev, fit_bitmap=False, score=False)
    eng.upload(local)
    eng.peer_setup(rank, world, (G + 31) // 32, _ag_factory(dist, world))
    dist.barrier()
    if rank == 0:
        eng.evaluate_async()                      # rank 1 stays away
        t0 = time.perf_counter()
        with pytest.raises(pkg.capi.BsError) as ei:
            eng.sync()
        assert ei.value.code == pkg.capi.BS_E_PEER
        assert time.perf_counter() - t0 < 5.0
        t0 = time.perf_counter()
        with pytest.raises(pkg.capi.BsError) as ei:
            eng.evaluate_async()                  # fast-fail: no further spinning
        assert ei.value.code == pkg.capi.BS_E_PEER and time.perf_counter() - t0 < 0.1
    dist.barrier()
    eng.peer_detach()
    dist.barrier()
    eng.peer_setup(rank, world, (G + 31) // 32, _ag_factory(dist, world))   # new epoch
    dist.barrier()
    for _ in range(2):
        eng.evaluate_async()
    eng.sync()
    words = eng.gathered_admit()
    res = eng.fetch()
    g0, g1 = local.meta["group_range"]
    own = np.unpackbits(words[rank].view(np.uint8), bitorder="little")[:G].astype(bool)
    np.testing.assert_array_equal(own[g0:g1], res.admit[g0:g1] == S.ADMIT)
    dist.barrier()
    eng.peer_detach()
    eng.close()
    if rank == 0:
        open(os.path.join(out_dir, "ok_timeout"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_parity(tmp_path):
    import torch
    import torch.multiprocessing as mp
    shared = torch.cuda.device_count() < 2
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), shared), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok"))


def test_peer_timeout_fast_fail_and_new_epoch(tmp_path):
    import torch
    import torch.multiprocessing as mp
    shared = torch.cuda.device_count() < 2
    mp.spawn(_worker_timeout, args=(2, _free_port(), str(tmp_path), shared), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok_timeout"))
