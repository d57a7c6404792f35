"""N > 1 path on CPU (gloo, world_size 2): the host-side rule that shards the hot path over GPUs -- disjoint sample-index
ranges per rank, one SUM reduce of the fp32 accumulation buffers to rank 0.  The renderer stand-in here is
the oracle in counter-stream mode, which obeys the same (seed, pixel, sample) addressing as the CUDA product."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yune_b200.dist import shard_samples, reduce_sum_to_root


def test_shard_samples_partition():
    for total in (0, 1, 7, 64, 1024, 16384):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_samples(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (b0, c0), (b1, c1) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_samples(8, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tests.refbind import Oracle, default_cam_array, load_golden_scene
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tris, mats, nodes = load_golden_scene("cornellbox")
    o = Oracle(); cam = default_cam_array()
    cfg = Oracle.config("udpt", rng_mode=1, seed=5, threads=2)
    W, total = 24, 6
    begin, count = shard_samples(total, rank, world)
    acc = np.zeros((W, W, 4), np.float32)
    for s in range(begin, begin + count):
        acc += o.samples(cfg, cam, tris, mats, nodes, W, W, s)      # rgb = radiance, a = 1  ->  sum buffer semantics
    t = torch.from_numpy(acc)
    reduce_sum_to_root(t)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.destroy_process_group()


def test_two_rank_sample_sharding_equals_single_rank(tmp_path, oracle):
    from tests.refbind import Oracle, default_cam_array, load_golden_scene
    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    tris, mats, nodes = load_golden_scene("cornellbox")
    cfg = Oracle.config("udpt", rng_mode=1, seed=5, threads=2)
    want = np.zeros_like(got)
    for s in range(6):
        want += oracle.samples(cfg, default_cam_array(), tris, mats, nodes, 24, 24, s)
    assert (got[..., 3] == 6).all()
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
