"""Multi-GPU sharding of the hot path: one process per GPU (torch.distributed, NCCL).

Counting is independent per sample and testing is independent per k-mer; the one coupling
is regrouping sample-major data into k-mer-major rows. The exchange is done
on the 2-bit packed streams, not on k-mers: every rank decodes its own samples
(rank r owns the contiguous block of samples [N*r/G, N*(r+1)/G)), the packed streams (3 bits / base) are all-gathered over
NVLink, and each rank then extracts only the canonical k-mers of its own contiguous k-mer
range — 16x less traffic than routing 4-byte k-mers to hash owners, no routing kernel, and
contiguous ranges keep the global sorted order (a survivor's global rank = range base + row),
which the reference's output order depends on. Ranges are balanced with splitters taken from
the quantiles of sample 0's sorted k-mer list (all samples are the same species).

Collectives: all_gather (stream lengths, then streams), all_reduce (U, the Bonferroni
denominator, modeling.py:641-644), gather of survivors to rank 0.
"""
import ctypes

import numpy as np

from .pipeline import KmerAssociation, PhenoResult, unpack_rows


def splitters_from_sample(ctx, n_ranges, sample_idx=0):
    """G-1 k-mer range boundaries = quantiles of one sample's sorted distinct k-mers."""
    if n_ranges <= 1:
        return []
    km, _ = ctx.sample_kmers(sample_idx)
    if len(km) < n_ranges:
        return [int(x) for x in np.linspace(0, 1 << (2 * ctx.k), n_ranges + 1)[1:-1]]
    return [int(km[len(km) * i // n_ranges]) for i in range(1, n_ranges)]


def sample_block(rank, world, n_samples):
    """Samples owned (ingested) by `rank`: a contiguous block."""
    return range(n_samples * rank // world, n_samples * (rank + 1) // world)


def range_of(rank, splitters):
    lo = 0 if rank == 0 else splitters[rank - 1]
    hi = 0 if rank == len(splitters) else splitters[rank]     # 0 = unbounded
    if lo == 0 and hi == 0:
        return None
    return lo, hi


class _DevView:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def exchange_streams(ka: KmerAssociation, my_samples, n_samples, rank, world, device):
    """All-gather every rank's packed streams; import the ones this rank does not own."""
    import torch
    import torch.distributed as dist

    ctx = ka.ctx
    exported = {s: ctx.export_stream(s) for s in my_samples}
    # 1. lengths of all samples
    lens = torch.zeros(n_samples, dtype=torch.int64, device=device)
    for s, (_, _, n) in exported.items():
        lens[s] = n
    dist.all_reduce(lens)
    lens_h = lens.cpu().numpy()
    per_rank = [list(sample_block(r, world, n_samples)) for r in range(world)]
    rank_pos = [int(sum(lens_h[s] for s in per_rank[r])) for r in range(world)]
    max_pos = max(rank_pos)
    # 2. one padded all_gather per array (seq: 2 bits/pos, bad: 1 bit/pos)
    seq_mine = torch.zeros(max_pos // 4, dtype=torch.uint8, device=device)
    bad_mine = torch.zeros(max_pos // 8, dtype=torch.uint8, device=device)
    off = 0
    for s in per_rank[rank]:
        sp, bp, n = exported[s]
        seq_mine[off // 4:(off + n) // 4] = torch.as_tensor(_DevView(sp, n // 4), device=device)
        bad_mine[off // 8:(off + n) // 8] = torch.as_tensor(_DevView(bp, n // 8), device=device)
        off += n
    seq_all = torch.empty(world * (max_pos // 4), dtype=torch.uint8, device=device)
    bad_all = torch.empty(world * (max_pos // 8), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(seq_all, seq_mine)
    dist.all_gather_into_tensor(bad_all, bad_mine)
    torch.cuda.synchronize(device)
    # 3. import the other ranks' samples
    for r in range(world):
        if r == rank:
            continue
        off = 0
        for s in per_rank[r]:
            n = int(lens_h[s])
            ctx.import_stream(s, seq_all.data_ptr() + r * (max_pos // 4) + off // 4,
                              bad_all.data_ptr() + r * (max_pos // 8) + off // 8, n)
            off += n
    return int(seq_mine.numel() + bad_mine.numel()) * (world - 1)   # bytes received per rank


def run_sharded(ka: KmerAssociation, buffers_by_sample, n_samples, k, pheno, binary, weights,
                rank, world, device, cutoff=1, **test_kw):
    """Whole hot path on `world` GPUs. buffers_by_sample: {sample_idx: bytes or (dev_ptr, n)} for the
    samples this rank owns (sample_block(rank, world, n_samples)). Returns (U_total, results-on-rank-0 or None, info)."""
    import torch
    import torch.distributed as dist

    ctx = ka.ctx
    ka.k, ka.n_samples = int(k), int(n_samples)
    ctx.begin(int(k), int(n_samples), int(cutoff))
    mine = list(sample_block(rank, world, n_samples))
    assert sorted(buffers_by_sample) == mine, "this rank must hold exactly its own block of samples"
    if mine:
        ctx.add_samples(mine[0], [buffers_by_sample[s] for s in mine])
    nvl_bytes = 0
    if world > 1:
        nvl_bytes = exchange_streams(ka, mine, n_samples, rank, world, device)
    spl = splitters_from_sample(ctx, world, 0)
    rng = range_of(rank, spl) if world > 1 else None
    U_local = ka.build(rng)
    u = torch.tensor([U_local], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(u)
    U_total = int(u.item())
    res = ka.test(pheno, binary, weights, n_union_total=U_total, **test_kw)
    info = {"U_local": U_local, "nvlink_bytes": nvl_bytes, "splitters": spl}
    if world == 1:
        return U_total, res, info
    # global rank of a survivor = (sum of U of lower ranges) + local row
    all_u = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_u, torch.tensor([U_local], dtype=torch.int64, device=device))
    base = int(sum(int(x.item()) for x in all_u[:rank]))
    payload = [(r.name, r.kmer, r.row + np.uint64(base), r.stat, r.p, r.mean_x, r.mean_y, r.n_with,
                r.presence) for r in res]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if rank != 0:
        return U_total, None, info
    out = []
    for j in range(len(res)):
        parts = [g[j] for g in gathered]
        out.append(PhenoResult(
            name=parts[0][0], kmer=np.concatenate([p[1] for p in parts]), row=np.concatenate([p[2] for p in parts]),
            stat=np.concatenate([p[3] for p in parts]), p=np.concatenate([p[4] for p in parts]),
            mean_x=np.concatenate([p[5] for p in parts]), mean_y=np.concatenate([p[6] for p in parts]),
            n_with=np.concatenate([p[7] for p in parts]), presence=np.concatenate([p[8] for p in parts])))
    return U_total, out, info
