"""Multi-GPU sharding of the hot path: one process per GPU (torch.distributed, NCCL).

Counting is independent per sample and testing is independent per k-mer; the one coupling
is regrouping sample-major data into k-mer-major rows. Rank r ingests the
contiguous block of samples [N*r/G, N*(r+1)/G) and owns one contiguous k-mer RANGE (not a
hash bucket): contiguous ranges keep the global sorted order — a survivor's global rank is
range base + row — which the reference's output column order depends on. Ranges are balanced
with splitters taken from the quantiles of sample 0's sorted k-mer list (same species).

Two exchange routes:
  * "alltoall" (k <= 24, assemblies): every rank extracts packed (k-mer, sample)
    records from ITS OWN samples only, already grouped by destination range
    (k_extract_part), and one NCCL all_to_all_single routes them over NVLink. All per-rank
    work (decode, extract, sort, rows, test) is divided by G.
  * "p2p": the same routing without NCCL: k_extract_part stores each record directly into its
    owner's receive buffer (the peer's sort input buffer, mapped with CUDA IPC) over NVLink, so
    the exchange is fused into the extraction kernel and overlaps with it.
    This is the default for assemblies.
  * "streams" (the route for raw reads / cutoff > 1 / k > 24): the 2-bit packed streams (3 bits per base, 16x
    fewer bytes than the k-mers) are all-gathered and each rank extracts its own range from all
    of them; extraction is then replicated on every rank.

Collectives: broadcast (splitters), all_to_all_single (counts, records) or all_reduce +
all_gather_into_tensor (streams), all_reduce (U, the Bonferroni denominator,
modeling.py:641-644), all_gather (per-range U), gather (survivors).
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


_FIELDS = (("kmer", np.uint64), ("row", np.uint64), ("stat", np.float64), ("p", np.float64),
           ("mean_x", np.float64), ("mean_y", np.float64), ("n_with", np.uint32))


def pack_results(res):
    """Survivors of all phenotypes -> (counts per phenotype, one flat uint8 buffer)."""
    counts = [len(r.kmer) for r in res]
    parts = []
    for r in res:
        for name, dt in _FIELDS:
            parts.append(np.ascontiguousarray(getattr(r, name), dtype=dt).view(np.uint8).reshape(-1))
        parts.append(np.ascontiguousarray(r.presence, dtype=np.uint8).reshape(-1))
    buf = np.concatenate(parts) if parts else np.empty(0, np.uint8)
    return counts, buf


def unpack_results(names, counts, buf, n_samples):
    """Inverse of pack_results -> list of tuples in merge_results' layout."""
    out, off = [], 0
    for name, n in zip(names, counts):
        vals = []
        for _, dt in _FIELDS:
            nb = n * np.dtype(dt).itemsize
            vals.append(buf[off:off + nb].view(dt).copy())
            off += nb
        pres = buf[off:off + n * n_samples].reshape(n, n_samples).copy()
        off += n * n_samples
        out.append((name, vals[0], vals[1], vals[2], vals[3], vals[4], vals[5], vals[6], pres))
    return out


def gather_results(res, U_local, rank, world, device, dist):
    """Per-range U + survivors to rank 0 with two tensor collectives (sizes, then one padded byte
    gather) -> merged list on rank 0, None elsewhere."""
    import torch
    n_samples = res[0].presence.shape[1] if res else 0
    counts, buf = pack_results(res)
    meta = torch.tensor([U_local, len(buf)] + counts, dtype=torch.int64, device=device)
    metas = torch.empty(world * len(meta), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(metas, meta)
    metas = metas.cpu().numpy().reshape(world, -1)
    us = [int(x) for x in metas[:, 0]]
    bases = [int(sum(us[:r])) for r in range(world)]
    cap = max(int(metas[:, 1].max()), 1)
    mine = torch.zeros(cap, dtype=torch.uint8, device=device)
    if len(buf):
        mine[:len(buf)] = torch.from_numpy(buf).to(device)
    if rank == 0:
        got = [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(world)]
        dist.gather(mine, got, dst=0)
        names = [r.name for r in res]
        gathered = []
        for r in range(world):
            b = got[r][:int(metas[r, 1])].cpu().numpy()
            gathered.append(unpack_results(names, [int(x) for x in metas[r, 2:]], b, n_samples))
        return merge_results(gathered, bases)
    dist.gather(mine, None, dst=0)
    return None


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


def exchange_records(ka: KmerAssociation, splitters, rank, world, device):
    """all-to-all of packed records by destination k-mer range. -> (recv tensor, bytes received)."""
    import torch
    import torch.distributed as dist

    ptr, counts = ka.ctx.extract_partition(splitters)
    send_counts = torch.tensor(counts, dtype=torch.int64, device=device)
    recv_counts = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(recv_counts, send_counts)
    rc = [int(x) for x in recv_counts.cpu().tolist()]
    total = int(sum(counts))
    send = (torch.as_tensor(_DevView(ptr, total * 8), device=device).view(torch.int64) if total
            else torch.empty(0, dtype=torch.int64, device=device))
    recv = torch.empty(int(sum(rc)), dtype=torch.int64, device=device)
    dist.all_to_all_single(recv, send, rc, counts)
    torch.cuda.synchronize(device)
    return recv, (int(sum(rc)) - rc[rank]) * 8


def exchange_records_p2p(ka: KmerAssociation, splitters, rank, world, device):
    """The all-to-all fused into the extraction kernel: after the ranks have exchanged their
    per-destination counts, k_extract_part stores every record straight into its owner's receive
    buffer (CUDA IPC mapping of the peer's sort input buffer) over NVLink — no send buffer, no NCCL
    all-to-all, no copy on the receiving side. -> (own receive pointer, records received, bytes in)."""
    import torch
    import torch.distributed as dist

    import os
    import sys
    import time
    ctx = ka.ctx
    tm = [time.time()] if (os.environ.get("PS_DIST_TIMING") and rank == 0) else None

    def lap():
        if tm is not None:
            torch.cuda.synchronize(device)
            tm.append(time.time())

    counts = ctx.partition_count(splitters)                       # records for each destination
    lap()
    mine = torch.tensor(counts, dtype=torch.int64, device=device)
    allc = torch.empty(world * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, mine)
    C = allc.cpu().numpy().reshape(world, world)                  # C[src][dst]
    n_recv = int(C[:, rank].sum())
    base = [int(C[:rank, d].sum()) for d in range(world)]         # my segment inside owner d's buffer
    my_ptr = ctx.recv_buffer(n_recv)
    h = torch.frombuffer(bytearray(ctx.ipc_export(my_ptr)), dtype=torch.uint8).to(device)
    allh = torch.empty(world * 64, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(allh, h)                          # also orders: every buffer is sized before anyone writes
    allh = allh.cpu().numpy().reshape(world, 64)
    ptrs = [my_ptr if d == rank else ctx.ipc_open(allh[d].tobytes()) for d in range(world)]
    lap()
    ctx.partition_write(ptrs, base)                               # returns when the remote stores have landed
    lap()
    dist.barrier()                                                # ... on every rank
    lap()
    if tm is not None:
        sys.stderr.write("[p2p exchange ms] count=%.2f meta=%.2f write=%.2f barrier=%.2f\n" % tuple(
            1e3 * (b - a) for a, b in zip(tm, tm[1:])))
    return my_ptr, n_recv, (n_recv - int(C[rank, rank])) * 8


def run_sharded(ka: KmerAssociation, buffers_by_sample, n_samples, k, pheno, binary, weights,
                rank, world, device, cutoff=1, route="auto", **test_kw):
    """Whole hot path on `world` GPUs. buffers_by_sample: {sample_idx: bytes or (dev_ptr, n)} for the
    samples this rank owns (sample_block(rank, world, n_samples)).
    Returns (U_total, results on rank 0 / None elsewhere, info)."""
    import os
    import time
    timing = os.environ.get("PS_DIST_TIMING") and rank == 0
    marks = []

    def mark(name):
        if timing:
            import torch
            torch.cuda.synchronize(device)
            marks.append((name, time.time()))

    ctx = ka.ctx
    ka.k, ka.n_samples = int(k), int(n_samples)
    mark("start")
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

    mark("ingest")
    is_text = lambda b: not isinstance(b, tuple)
    reads = any(is_text(b) and bytes(b[:1]) == b"@" for b in buffers_by_sample.values())
    if route == "auto":
        # measured on config 2 (250 x 4.3 Mbp), ms per step at 2 / 4 / 8 GPUs: p2p 30.4 / 17.2 / 13.7,
        # streams 30.5 / 21.5 / 15.5, NCCL alltoall 46.3 / 25.0 / 16.1
        flag = torch.tensor([1 if (reads or cutoff > 1 or k > 24) else 0], dtype=torch.int64, device=device)
        dist.all_reduce(flag)
        route = "streams" if int(flag.item()) else "p2p"
    if route in ("alltoall", "p2p") and (reads or cutoff > 1 or k > 24):
        raise ValueError(f"route='{route}' handles assemblies with cutoff 1 and k <= 24 only")
    if route in ("alltoall", "p2p"):
        # rank 0 holds sample 0: its quantiles are the range boundaries for everybody
        spl_t = torch.zeros(world - 1, dtype=torch.int64, device=device)
        if rank == 0:
            q = ctx.sample_quantiles(0, world)
            spl_t = torch.tensor(np.array(q, dtype=np.uint64).view(np.int64), dtype=torch.int64, device=device)
        dist.broadcast(spl_t, src=0)
        spl = [int(x) for x in spl_t.cpu().numpy().view(np.uint64)]
        mark("splitters")
        if route == "p2p":
            ptr, n_recv, nvl_bytes = exchange_records_p2p(ka, spl, rank, world, device)
            mark("exchange")
            U_local = ctx.build_from_records(ptr, n_recv)
        else:
            recv, nvl_bytes = exchange_records(ka, spl, rank, world, device)
            mark("exchange")
            U_local = ctx.build_from_records(recv.data_ptr(), recv.numel())
            del recv
        ka.U = U_local
    else:
        nvl_bytes = exchange_streams(ka, n_samples, rank, world, device)
        mark("exchange")
        spl = ctx.sample_quantiles(0, world)          # identical on every rank: all hold sample 0
        mark("splitters")
        U_local = ka.build(range_of(rank, spl))
    mark("build")
    u = torch.tensor([U_local], dtype=torch.int64, device=device)
    dist.all_reduce(u)
    U_total = int(u.item())
    res = ka.test(pheno, binary, weights, n_union_total=U_total, **test_kw)
    mark("test")
    merged = gather_results(res, U_local, rank, world, device, dist)
    mark("gather")
    if timing:
        import sys
        sys.stderr.write("[dist timing ms] " + " ".join(
            f"{b[0]}={1e3 * (b[1] - a[1]):.1f}" for a, b in zip(marks, marks[1:])) + "\n")
    return U_total, merged, {"U_local": U_local, "nvlink_bytes": nvl_bytes, "splitters": spl, "route": route}
