"""CPU, world_size 2 over gloo: the multi-rank plumbing of the spp-sharded render (disjoint sample
ranges per rank + a SUM reduce of the float accumulation buffers onto rank 0), with the oracle
standing in for the per-rank render since there is no GPU here."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, scene_path


def test_shard_samples_partition(rtc):
    for spp in (1, 2, 7, 128, 1024):
        for world in (1, 2, 3, 4, 8):
            ranges = [rtc.shard_samples(spp, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == spp
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        rtc.shard_samples(8, 2, 2)


def _worker(rank, world, port, scene, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orclib
    import raytracing_course_b200 as rtc
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    s = orclib.Scene(orclib.oracle(), scene)
    s.override(24, 16, 6)
    lo, hi = rtc.shard_samples(s.samples, rank, world)
    part, paths, rays = s.render_sum(77, lo, hi - lo, nthreads=1)
    acc = torch.from_numpy(part.copy())
    dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
    cnt = torch.tensor([paths, rays], dtype=torch.int64)
    dist.all_reduce(cnt)
    if rank == 0:
        full, fpaths, frays = s.render_sum(77, 0, s.samples, nthreads=1)
        np.savez(out_path, reduced=acc.numpy(), full=full, cnt=cnt.numpy(), fcnt=np.array([fpaths, frays]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_spp_shards_sum_to_the_full_render(tmp_path):
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    out = str(tmp_path / "r.npz")
    mp.spawn(_worker, args=(2, port, scene_path("lights_mix"), out), nprocs=2, join=True)
    z = np.load(out)
    assert np.array_equal(z["cnt"], z["fcnt"])  # every path and ray accounted for exactly once
    ok = np.isfinite(z["full"]) & np.isfinite(z["reduced"])
    assert ok.mean() > 0.999
    assert np.allclose(z["reduced"][ok], z["full"][ok], rtol=1e-5, atol=1e-6)
