"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard a frame sequence (vision-enhanced-lidar-odometry_b200/shard.py),
each computes its frames' normal equations (oracle on thinned scans stands in for the device here — this test is about
the sharding + host gather, not the kernels), rank 0 gathers and the result equals the single-process run."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_range_partitions(velo):
    sh = velo.shard if hasattr(velo, "shard") else __import__("importlib").import_module("vision-enhanced-lidar-odometry_b200.shard")
    for total in (1, 7, 8, 8000, 8001):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                s, c = sh.frame_range(total, world, r)
                seen += list(range(s, s + c))
                hs, hc = sh.halo_range(total, world, r)
                assert (hs, hc) == (s - 1, c + 1)
            assert seen == list(range(total))


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import importlib
    import torch.distributed as dist
    import pyoracle
    from conftest import small_scan
    velo = importlib.import_module("vision-enhanced-lidar-odometry_b200")
    sh = importlib.import_module("vision-enhanced-lidar-odometry_b200.shard")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = pyoracle.Oracle()
    P, Tr, w, h = velo.synth.calib_raw(0)
    cal = orc.calib_from_kitti(P, Tr, w, h)
    prm = velo.api.default_params()
    start, count = sh.frame_range(total, world, rank)
    rows = np.zeros((count, velo.abi.NEQ_STRIDE))
    base = 500
    prev = orc.segment(small_scan(velo, base + start - 1, range(28, 34), 6), cal)      # halo scan
    for i in range(count):
        cur = orc.segment(small_scan(velo, base + start + i, range(28, 34), 6), cal)
        _, neq, _ = orc.icp_pass(cur[0], cur[1], prev[0], prev[1], velo.synth.pose_guess(base + start + i, 0), 1, 2, prm, 1)
        rows[i] = neq
        prev = cur
    out = sh.gather_rows(dist, rows, dst=0)
    g = sh.RowGather(dist, count, rows.shape[1:], rows.dtype, dst=0)      # the asynchronous form bench.py overlaps with the next step
    for _ in range(2):
        g.start(rows)
        out2 = g.wait()
    if rank == 0:
        assert out2.tobytes() == out.tobytes()
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    total = 5
    res = {}
    for world, port in ((1, 29731), (2, 29732)):
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
        for p in procs:
            p.start()
        res[world] = q.get(timeout=300)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    assert res[1].shape == (total, 64) and res[1].tobytes() == res[2].tobytes()
    assert (res[1][:, 56] > 10).all()
