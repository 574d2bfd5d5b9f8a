"""Multi-GPU sharding of the hot path: one process per GPU (torch.distributed, NCCL).

Counting is independent per sample and testing is independent per k-mer; the one coupling
is regrouping sample-major data into k-mer-major rows. The exchange is done
on the 2-bit packed streams, not on k-mers: every rank decodes its own samples (rank r owns
the contiguous block of samples [N*r/G, N*(r+1)/G)), the packed streams (3 bits / base) are
all-gathered over NVLink, and each rank then extracts only the canonical k-mers of its own
contiguous k-mer range — 16x less traffic than routing 4-byte k-mers to hash owners, no
routing kernel, and contiguous ranges keep the global sorted order (a survivor's global rank
= range base + row), which the reference's output order depends on. Ranges are balanced with
splitters taken from the quantiles of sample 0's sorted k-mer list (same species).

Collectives: all_reduce (stream lengths), all_gather_into_tensor (streams), all_reduce (U, the
Bonferroni denominator, modeling.py:641-644), all_gather (per-range U), gather (survivors).
"""
import numpy as np

from .pipeline import KmerAssociation, PhenoResult


def sample_block(rank, world, n_samples):
    """Samples owned (ingested) by `rank`: a contiguous block."""
    return range(n_samples * rank // world, n_samples * (rank + 1) // world)


def range_of(rank, splitters):
    """k-mer range [lo, hi) of `rank` given world-1 ascending splitters; None = whole space."""
    if not splitters:
        return None
    lo = 0 if rank == 0 else splitters[rank - 1]
    hi = 0 if rank == len(splitters) else splitters[rank]     # 0 = unbounded
    return lo, hi


def stream_layout(lens, world):
    """Per-rank byte layout of the padded all-gather. lens[s] = padded positions of sample s.
    -> (per_rank sample lists, per-rank total positions, max positions)."""
    n = len(lens)
    per_rank = [list(sample_block(r, world, n)) for r in range(world)]
    rank_pos = [int(sum(int(lens[s]) for s in per_rank[r])) for r in range(world)]
    return per_rank, rank_pos, max(rank_pos) if rank_pos else 0


def merge_results(gathered, bases):
    """Rank-0 merge of per-range survivor lists. gathered[r] = list over phenotypes of tuples
    (name, kmer, row, stat, p, mean_x, mean_y, n_with, presence); bases[r] = global rank of
    range r's first union k-mer. Ranges are ascending, so concatenation keeps k-mer order."""
    out = []
    for j in range(len(gathered[0])):
        parts = [g[j] for g in gathered]
        rows = [np.asarray(p[2], dtype=np.uint64) + np.uint64(bases[r]) for r, p in enumerate(parts)]
        out.append(PhenoResult(
            name=parts[0][0], kmer=np.concatenate([p[1] for p in parts]), row=np.concatenate(rows),
            stat=np.concatenate([p[3] for p in parts]), p=np.concatenate([p[4] for p in parts]),
            mean_x=np.concatenate([p[5] for p in parts]), mean_y=np.concatenate([p[6] for p in parts]),
            n_with=np.concatenate([p[7] for p in parts]), presence=np.concatenate([p[8] for p in parts])))
    return out


def gather_results(res, U_local, rank, world, device, dist):
    """all_gather the per-range U, gather survivors on rank 0 -> merged list (rank 0) or None."""
    import torch
    all_u = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_u, torch.tensor([U_local], dtype=torch.int64, device=device))
    us = [int(x.item()) for x in all_u]
    bases = [int(sum(us[:r])) for r in range(world)]
    payload = [(r.name, r.kmer, r.row, r.stat, r.p, r.mean_x, r.mean_y, r.n_with, r.presence) for r in res]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if rank != 0:
        return None
    return merge_results(gathered, bases)


class _DevView:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def exchange_streams(ka: KmerAssociation, n_samples, rank, world, device):
    """All-gather every rank's packed stream block; import the blocks this rank does not own.
    Returns the bytes received over NVLink by this rank."""
    import torch
    import torch.distributed as dist

    ctx = ka.ctx
    mine = list(sample_block(rank, world, n_samples))
    lens = torch.zeros(n_samples, dtype=torch.int64, device=device)
    first = None
    if mine:
        exported = [ctx.export_stream(s) for s in mine]       # consecutive in this rank's pool
        first = exported[0]
        lens[mine[0]:mine[-1] + 1] = torch.tensor([e[2] for e in exported], dtype=torch.int64, device=device)
    dist.all_reduce(lens)
    lens_h = lens.cpu().numpy()
    per_rank, rank_pos, max_pos = stream_layout(lens_h, world)
    seq_all = torch.empty(world * (max_pos // 4), dtype=torch.uint8, device=device)
    bad_all = torch.empty(world * (max_pos // 8), dtype=torch.uint8, device=device)
    seq_mine = seq_all[rank * (max_pos // 4):(rank + 1) * (max_pos // 4)]
    bad_mine = bad_all[rank * (max_pos // 8):(rank + 1) * (max_pos // 8)]
    my_pos = rank_pos[rank]
    if mine:
        seq_mine[:my_pos // 4] = torch.as_tensor(_DevView(first[0], my_pos // 4), device=device)
        bad_mine[:my_pos // 8] = torch.as_tensor(_DevView(first[1], my_pos // 8), device=device)
    dist.all_gather_into_tensor(seq_all, seq_mine.clone())
    dist.all_gather_into_tensor(bad_all, bad_mine.clone())
    torch.cuda.synchronize(device)
    for r in range(world):
        if r == rank or not per_rank[r]:
            continue
        ctx.import_streams(per_rank[r][0], seq_all.data_ptr() + r * (max_pos // 4),
                           bad_all.data_ptr() + r * (max_pos // 8), [int(lens_h[s]) for s in per_rank[r]])
    return int((max_pos // 4 + max_pos // 8) * (world - 1))


def run_sharded(ka: KmerAssociation, buffers_by_sample, n_samples, k, pheno, binary, weights,
                rank, world, device, cutoff=1, **test_kw):
    """Whole hot path on `world` GPUs. buffers_by_sample: {sample_idx: bytes or (dev_ptr, n)} for the
    samples this rank owns (sample_block(rank, world, n_samples)).
    Returns (U_total, results on rank 0 / None elsewhere, info)."""
    ctx = ka.ctx
    ka.k, ka.n_samples = int(k), int(n_samples)
    ctx.begin(int(k), int(n_samples), int(cutoff))
    mine = list(sample_block(rank, world, n_samples))
    assert sorted(buffers_by_sample) == mine, "this rank must hold exactly its own block of samples"
    if mine:
        ctx.add_samples(mine[0], [buffers_by_sample[s] for s in mine])
    if world == 1:
        U = ka.build()
        res = ka.test(pheno, binary, weights, **test_kw)
        return U, res, {"U_local": U, "nvlink_bytes": 0, "splitters": []}
    import torch
    import torch.distributed as dist

    nvl_bytes = exchange_streams(ka, n_samples, rank, world, device)
    spl = ctx.sample_quantiles(0, world)          # identical on every rank: all hold sample 0
    U_local = ka.build(range_of(rank, spl))
    u = torch.tensor([U_local], dtype=torch.int64, device=device)
    dist.all_reduce(u)
    U_total = int(u.item())
    res = ka.test(pheno, binary, weights, n_union_total=U_total, **test_kw)
    merged = gather_results(res, U_local, rank, world, device, dist)
    return U_total, merged, {"U_local": U_local, "nvlink_bytes": nvl_bytes, "splitters": spl}
