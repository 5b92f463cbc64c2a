"""N > 1 path on CPU: world_size-2 gloo job.  Each rank generates and decodes ITS shard of a
texture (with the oracle standing in for the GPU kernels, which need a device), writes it into
the shared result at the offsets the sharding module computes, and the validity flag is
AND-reduced -- the result must equal the single-rank decode byte for byte."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from detex_b200 import sharding as S  # noqa: E402


def test_tiled_shards_partition_the_range():
    for n in (0, 1, 7, 1000, 100_000_000, 33_333_334):
        for world in (1, 2, 3, 4, 8):
            for align in (1, 2):
                edges = [S.tiled_shard(n, r, world, align) for r in range(world)]
                assert edges[0][0] == 0 and edges[-1][1] == n
                for a, b in zip(edges, edges[1:]):
                    assert a[1] == b[0] and a[0] <= a[1]
                assert all(lo % align == 0 for lo, _ in edges)


def test_linear_shard_views_tile_the_image():
    w, h, wib, hib, bb, ps = 1021, 37, 256, 10, 16, 4
    for world in (1, 2, 3, 8, 16):
        views = [S.linear_shard_view(w, h, wib, hib, bb, ps, r, world) for r in range(world)]
        assert sum(v["in_bytes"] for v in views) == wib * hib * bb
        assert sum(v["out_bytes"] for v in views) == w * h * ps
        off = 0
        for v in views:
            assert v["out_offset"] == off or v["out_bytes"] == 0
            off += v["out_bytes"]


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist
    from oracle import harness as H
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    orc = H.load_oracle()
    fmt = "BPTC"
    tf, bs, ps = H.FORMATS[fmt]
    w, h = 253, 147
    wib, hib = 64, 37
    v = S.linear_shard_view(w, h, wib, hib, bs, ps, rank, world)
    # each rank synthesises only its own index range (counter-based generator)
    blocks = H.synth_blocks(fmt, v["rows"] * wib, seed=9, first_block=v["row_lo"] * wib)
    out, ok, _ = H.cpu_decode(orc, fmt, blocks, w, v["height"], linear=True)
    assert out.size == v["out_bytes"]
    np.save(os.path.join(tmpdir, "out%d.npy" % rank), out)
    ok_all = S.all_valid(ok)
    np.save(os.path.join(tmpdir, "ok%d.npy" % rank), np.array([ok, ok_all]))
    # optional final-image gather: ragged slabs (18 and 19 block rows), then equal ones
    import torch
    full = S.gather_image(torch.from_numpy(out))
    np.save(os.path.join(tmpdir, "gathered%d.npy" % rank), full.numpy())
    eq = S.gather_image(torch.full((1000,), rank + 1, dtype=torch.uint8), out=torch.zeros(2048, dtype=torch.uint8))
    assert eq[:1000].eq(1).all() and eq[1000:2000].eq(2).all() and eq[2000:].eq(0).all()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_linear_decode_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    from oracle import harness as H
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    orc = H.load_oracle()
    blocks = H.synth_blocks("BPTC", 64 * 37, seed=9)
    want, ok, _ = H.cpu_decode(orc, "BPTC", blocks, 253, 147, linear=True)
    got = np.concatenate([np.load(tmp_path / ("out%d.npy" % r)) for r in range(world)])
    assert np.array_equal(got, want)
    oks = [np.load(tmp_path / ("ok%d.npy" % r)) for r in range(world)]
    assert all(bool(o[1]) == ok for o in oks)          # reduced flag == single-rank result
    assert any(not bool(o[0]) for o in oks) == (not ok)
    for r in range(world):                             # every rank holds the whole image after the gather
        assert np.array_equal(np.load(tmp_path / ("gathered%d.npy" % r)), want)
