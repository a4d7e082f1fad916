"""GPU, 2 ranks over NCCL: group-sharded evaluation on the CUDA engine + one all-gather of the admit
bitmap reproduces the unsharded oracle round.  Skipped on a single-GPU box."""
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


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pkg = importlib.import_module("batch-scheduler_b200")
    S = pkg.snapshot
    full = S.config(3, 0.1).resolve_groups()
    local = full.shard_groups(rank, world)
    g0, g1 = local.meta["group_range"]
    eng = pkg.Engine(local.lanes, rank, fit_bitmap=False, score=False)
    eng.upload(local)
    eng.evaluate_async()
    # the admit bitmap straight from the engine's device buffer, gathered on the engine's stream
    ptr, nbytes = eng.device_buffer(pkg.capi.BUF_ADMIT_BITMAP)

    class H:
        pass
    h = H()
    h.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4", "data": (ptr, False), "version": 2}
    mine = torch.as_tensor(h, device=f"cuda:{rank}")
    gathered = torch.empty(world * mine.numel(), dtype=torch.int32, device=f"cuda:{rank}")
    ext = torch.cuda.ExternalStream(eng.stream(), device=rank)
    with torch.cuda.stream(ext):
        dist.all_gather_into_tensor(gathered, mine)
    eng.sync()
    torch.cuda.synchronize()
    res = eng.fetch()
    G = full.groups.n
    words = gathered.cpu().numpy().view(np.uint32).reshape(world, -1)
    bounds = [None] * world
    b = torch.tensor([g0, g1], dtype=torch.int64, device=f"cuda:{rank}")
    allb = [torch.zeros_like(b) for _ in range(world)]
    dist.all_gather(allb, b)
    merged = np.zeros(G, bool)
    for r in range(world):
        a0, a1 = (int(x) for x in allb[r].cpu())
        bits = np.unpackbits(words[r].view(np.uint8), bitorder="little")[:G].astype(bool)
        merged[a0:a1] = bits[a0:a1]
    # the same all-gather by the engine's own peer-memory kernel (CUDA IPC over NVLink), three rounds
    def _ag(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    eng.peer_setup(rank, world, (G + 31) // 32, _ag)
    dist.barrier()
    for _ in range(3):
        eng.evaluate_async()
        eng.sync()
    p2p = eng.gathered_admit()
    np.testing.assert_array_equal(p2p, words[:, :p2p.shape[1]])
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


def test_two_rank_sharded_parity(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok"))
