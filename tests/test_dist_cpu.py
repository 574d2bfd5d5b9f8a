"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo for the collectives that carry
U and the survivors, plus the pure layout functions. (The CUDA work of each rank is covered by
the -m gpu tests; `test_kmer_range_shards_partition_the_union` checks range sharding itself.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from phenotypeseeker_b200 import dist as psdist
from phenotypeseeker_b200.pipeline import PhenoResult


def test_sample_blocks_cover_every_sample_once():
    for n, w in ((250, 8), (5, 8), (7, 2), (1, 1), (5000, 4)):
        got = [s for r in range(w) for s in psdist.sample_block(r, w, n)]
        assert got == list(range(n))


def test_range_of_partitions_the_kmer_space():
    spl = [100, 2000, 2 ** 31]
    rs = [psdist.range_of(r, spl) for r in range(4)]
    assert rs[0] == (0, 100) and rs[1] == (100, 2000) and rs[2] == (2000, 2 ** 31) and rs[3] == (2 ** 31, 0)
    assert psdist.range_of(0, []) is None


def test_stream_layout():
    lens = np.array([4096, 8192, 4096, 12288, 4096])
    per_rank, pos, mx = psdist.stream_layout(lens, 2)
    assert per_rank == [[0, 1], [2, 3, 4]] and pos == [12288, 20480] and mx == 20480


def _fake_result(rank, j):
    """rank doubles as the index of a k-mer range: ranges ascend with it."""
    rng = np.random.default_rng(10 * rank + j)
    n = 3 + rank + j
    return PhenoResult(name=f"ph{j}", kmer=np.sort(rng.integers(0, 1000, n).astype(np.uint64)) + np.uint64(1000 * rank),
                       row=np.arange(n, dtype=np.uint64), stat=rng.random(n), p=rng.random(n), mean_x=rng.random(n),
                       mean_y=rng.random(n), n_with=rng.integers(2, 9, n).astype(np.uint32),
                       presence=rng.integers(0, 2, (n, 6)).astype(np.uint8))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # two passes over the k-mer space: ranges ascend pass-major, then by rank -> range index = pass * world + rank
        parts = [(100 + 11 * (ps * world + rank), [_fake_result(ps * world + rank, j) for j in range(2)]) for ps in range(2)]
        u = torch.tensor([sum(p[0] for p in parts)], dtype=torch.int64)
        dist.all_reduce(u)                                   # Bonferroni denominator
        merged = psdist.gather_results(parts, rank, world, torch.device("cpu"), dist)
        if rank == 0:
            np.savez(out_path, U=int(u.item()), **{f"row{j}": merged[j].row for j in range(2)},
                     **{f"kmer{j}": merged[j].kmer for j in range(2)}, **{f"pres{j}": merged[j].presence for j in range(2)})
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def test_pack_unpack_roundtrip():
    res = [_fake_result(0, 0), _fake_result(1, 1)]
    counts, buf = psdist.pack_results(res)
    back = psdist.unpack_results(["a", "b"], counts, buf, 6)
    for r, t in zip(res, back):
        assert np.array_equal(t[1], r.kmer) and np.array_equal(t[3], r.stat) and np.array_equal(t[8], r.rowbits)
        assert np.array_equal(t[7], r.n_with)


def test_gloo_world2_union_size_and_survivor_gather(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    us = [100 + 11 * i for i in range(4)]
    assert int(z["U"]) == sum(us)
    for j in range(2):
        rs = [_fake_result(i, j) for i in range(4)]
        assert np.array_equal(z[f"kmer{j}"], np.concatenate([r.kmer for r in rs]))
        assert np.array_equal(z[f"row{j}"], np.concatenate([r.row + np.uint64(sum(us[:i])) for i, r in enumerate(rs)]))   # base = U of lower ranges
        assert np.array_equal(z[f"pres{j}"], np.concatenate([r.presence for r in rs]))
        assert (np.diff(z[f"kmer{j}"].astype(np.int64)) >= 0).all()                             # still ascending
