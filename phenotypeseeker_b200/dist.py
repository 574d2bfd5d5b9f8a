"""Multi-GPU sharding of the hot path: one process per GPU (torch.distributed, NCCL).

Counting is independent per sample and testing is independent per k-mer; the one coupling
is regrouping sample-major data into k-mer-major rows. Rank r ingests the
contiguous block of samples [N*r/G, N*(r+1)/G) and owns one contiguous k-mer RANGE (not a
hash bucket): contiguous ranges keep the global sorted order — a survivor's global rank is
range base + row — which the reference's output column order depends on. Ranges are balanced
with splitters taken once per job from the quantiles of sample 0's sorted k-mer list (same species).

Exchange routes:
  * "pages" (k = 9..16, assemblies and raw reads alike; the default): every GPU owns a receive
    pool of 4 KB pages, mapped into its peers with CUDA IPC once per job. The extraction kernel of
    a sender (k_scatter1) groups its k-mer instances by (owner, top k-mer byte) and appends each
    group straight to pages of the owner's pool — 4-byte records, ~128-byte runs, plain stores over
    NVLink. No counts are exchanged, nothing is extracted twice, there is no send buffer and no
    copy on the receiver: the all-to-all IS the write-out of the extraction kernel, and what the
    receiver finds is already the level-1 partition of its own range. Two stream-ordered barriers
    (tiny all-reduces) fence the exchange; the host is not involved.
  * "streams" (k > 16): the 2-bit packed streams (3 bits per base) are all-gathered and each
    rank extracts its own range from all of them; extraction is then replicated on every rank.

Collectives per step: all_gather (pool sizes; sticky, so IPC handles move only when a pool grows),
2 x all_reduce (barriers), all_reduce (U, the Bonferroni denominator, modeling.py:641-644, plus
the pool-overflow flag), one all_gather of the packed survivors.
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
    (name, kmer, row, stat, p, mean_x, mean_y, n_with, rowbits, n_samples); bases[r] = global rank of
    range r's first union k-mer. Ranges are ascending, so concatenation keeps k-mer order."""
    out = []
    for j in range(len(gathered[0])):
        parts = [g[j] for g in gathered]
        rows = [np.asarray(p[2], dtype=np.uint64) + np.uint64(bases[r]) for r, p in enumerate(parts)]
        out.append(PhenoResult(
            name=parts[0][0], kmer=np.concatenate([p[1] for p in parts]), row=np.concatenate(rows),
            stat=np.concatenate([p[3] for p in parts]), p=np.concatenate([p[4] for p in parts]),
            mean_x=np.concatenate([p[5] for p in parts]), mean_y=np.concatenate([p[6] for p in parts]),
            n_with=np.concatenate([p[7] for p in parts]), rowbits=np.concatenate([p[8] for p in parts]),
            n_samples=parts[0][9]))
    return out


_FIELDS = (("kmer", np.uint64), ("row", np.uint64), ("stat", np.float64), ("p", np.float64),
           ("mean_x", np.float64), ("mean_y", np.float64), ("n_with", np.uint32))


def pack_results(res):
    """Survivors of all phenotypes -> (counts per phenotype, one flat uint8 buffer); matrix rows travel
    bit-packed as they left the GPU (row_words x 4 bytes per survivor)."""
    counts = [len(r.kmer) for r in res]
    parts = []
    for r in res:
        for name, dt in _FIELDS:
            parts.append(np.ascontiguousarray(getattr(r, name), dtype=dt).view(np.uint8).reshape(-1))
        parts.append(np.ascontiguousarray(r.rowbits, dtype=np.uint32).view(np.uint8).reshape(-1))
    buf = np.concatenate(parts) if parts else np.empty(0, np.uint8)
    return counts, buf


def unpack_results(names, counts, buf, n_samples):
    """Inverse of pack_results -> list of tuples in merge_results' layout."""
    out, off = [], 0
    W = (((n_samples + 31) // 32) + 3) // 4 * 4
    for name, n in zip(names, counts):
        vals = []
        for _, dt in _FIELDS:
            nb = n * np.dtype(dt).itemsize
            vals.append(buf[off:off + nb].view(dt).copy())
            off += nb
        rb = buf[off:off + n * W * 4].view(np.uint32).reshape(n, W).copy()
        off += n * W * 4
        out.append((name, vals[0], vals[1], vals[2], vals[3], vals[4], vals[5], vals[6], rb, n_samples))
    return out


_GATHER_CAP = {"bytes": 1 << 16}     # sticky capacity of the survivor all-gather (grows when a rank needs more)


def gather_results(parts, rank, world, device, dist, top_k=None, p_exact=None):
    """Per-range U + survivors of every rank with ONE collective: an all-gather of fixed-capacity
    buffers [header | payload]. parts = this rank's passes, each (U_local, [PhenoResult per column]);
    header = passes x (U_local, payload bytes, survivors per column). Every rank sees every header, so
    all ranks agree when the capacity has to grow (then, and only then, a second round). Ranges ascend
    pass-major, then by rank: that is the order of the merge and of the global row numbers.
    p_exact: final threshold (passes test against a laxer one while U is still growing).
    -> merged list on rank 0, None elsewhere."""
    import torch
    from .pipeline import trim_top
    res0 = parts[0][1]
    n_samples = res0[0].n_samples if res0 else 0
    P = len(res0)
    hdr_rows, bufs = [], []
    for U_local, res in parts:
        counts, buf = pack_results(res)
        hdr_rows.append([U_local, len(buf)] + counts)
        bufs.append(buf)
    buf = np.concatenate(bufs) if bufs else np.empty(0, np.uint8)
    hdr = np.array(hdr_rows, dtype=np.int64).reshape(-1).view(np.uint8)
    while True:
        cap = _GATHER_CAP["bytes"]
        mine = np.zeros(len(hdr) + cap, dtype=np.uint8)
        mine[:len(hdr)] = hdr
        if len(buf) <= cap:
            mine[len(hdr):len(hdr) + len(buf)] = buf
        t = torch.from_numpy(mine).to(device)
        allt = torch.empty(world * len(mine), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allt, t)
        allh = allt.cpu().numpy().reshape(world, -1)
        metas = allh[:, :len(hdr)].copy().view(np.int64).reshape(world, len(parts), -1)
        need = int(metas[:, :, 1].sum(axis=1).max())
        if need <= cap:
            break
        _GATHER_CAP["bytes"] = 2 * need
    if rank != 0:
        return None
    names = [r.name for r in res0]
    gathered, bases, base = [], [], 0
    offs = [0] * world
    for j in range(len(parts)):
        for r in range(world):
            nb = int(metas[r, j, 1])
            payload = allh[r, len(hdr) + offs[r]:len(hdr) + offs[r] + nb]
            offs[r] += nb
            gathered.append(unpack_results(names, [int(x) for x in metas[r, j, 2:2 + P]], payload, n_samples))
            bases.append(base)
            base += int(metas[r, j, 0])
    out = []
    for r in merge_results(gathered, bases):
        if p_exact is not None:
            r = r.take(r.p < p_exact)
        out.append(trim_top(r, top_k))
    return out


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


POOL_BUDGET_BYTES = 38e9      # one level-1 pool per GPU; the level-2 pool is as large again


class PageRoute:
    """Per-job state of the "pages" exchange on one rank: pass count, splitters, the agreed sub-pool size
    and the CUDA IPC mappings of the peers' pools. Everything here is set up once per job and reused by
    every step (a step only all-gathers the per-rank instance counts to see that the job is the same);
    it is redone, collectively, when the job changes or a pool has to grow."""

    def __init__(self, ka, rank, world, device):
        self.ka, self.rank, self.world, self.device = ka, rank, world, device
        self.key = None            # (k, n_samples, per-rank instance counts) the plan was made for
        self.passes = 1            # super-ranges of the k-mer space, one exchange each
        self.splitters = None      # world * passes - 1 ascending k-mer boundaries
        self.pages = 0             # pages per sender in every receive pool: agreed target (sticky maximum)
        self.live_pages = 0        # ... and what the pools are set up for right now
        self.bar = None

    def barrier(self, dist):
        """Stream-ordered barrier: an all-reduce of one word on the library's stream (no host sync)."""
        dist.all_reduce(self.bar)

    def pass_range(self, j):
        """(lo, hi) of super-range j (0 = unbounded) and the world - 1 splitters inside it."""
        W, S = self.world, self.splitters
        lo = 0 if j == 0 else S[j * W - 1]
        hi = 0 if j == self.passes - 1 else S[(j + 1) * W - 1]
        return lo, hi, S[j * W:j * W + W - 1]

    def prepare(self, k, n_samples, dist, torch):
        ctx = self.ka.ctx
        if self.bar is None:
            self.bar = torch.zeros(1, dtype=torch.int32, device=self.device)
        mine = torch.tensor([ctx.instances_upper()], dtype=torch.int64, device=self.device)
        allc = torch.empty(self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(allc, mine)
        counts = tuple(int(x) for x in allc.cpu().tolist())
        key = (k, n_samples, counts)
        if key != self.key:
            # a new job: passes so that a pool fits the budget, then quantile splitters of sample 0 (held by rank 0)
            total = sum(counts)
            self.passes = max(1, int(np.ceil(total * 1.5 * 4 / (self.world * POOL_BUDGET_BYTES))))
            nq = self.world * self.passes
            spl_t = torch.zeros(max(nq - 1, 1), dtype=torch.int64, device=self.device)
            if self.rank == 0:
                q = ctx.sample_quantiles(0, nq)
                spl_t[:len(q)] = torch.tensor(np.array(q, dtype=np.uint64).view(np.int64), dtype=torch.int64,
                                              device=self.device)
            dist.broadcast(spl_t, src=0)
            self.splitters = [int(x) for x in spl_t.cpu().numpy().view(np.uint64)][:nq - 1]
            need = torch.tensor([ctx.route_pages_needed(self.world, self.passes)], dtype=torch.int64, device=self.device)
            allneed = torch.empty(self.world, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(allneed, need)
            self.pages = max(self.pages, int(allneed.max().item()))
            self.key = key
        if self.pages != self.live_pages:
            # pools (re)allocated on every rank: unmap the old ones first, then exchange the new handles
            ctx.ipc_close_all()
            self.barrier(dist)
            torch.cuda.synchronize(self.device)
            pool, meta = ctx.route_setup(self.world, self.rank, self.pass_range(0)[2], self.pages)
            h = np.frombuffer(ctx.ipc_export(pool) + ctx.ipc_export(meta), dtype=np.uint8).copy()
            allh = torch.empty(self.world * 128, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh, torch.from_numpy(h).to(self.device))
            allh = allh.cpu().numpy().reshape(self.world, 128)
            pools = [pool if d == self.rank else ctx.ipc_open(allh[d, :64].tobytes()) for d in range(self.world)]
            metas = [meta if d == self.rank else ctx.ipc_open(allh[d, 64:].tobytes()) for d in range(self.world)]
            ctx.route_peers(pools, metas)
            self.live_pages = self.pages


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
    if world == 1:
        ctx.route_clear()
    if mine:
        ctx.add_samples(mine[0], [buffers_by_sample[s] for s in mine])
    if world == 1:
        U = ka.build()
        res = ka.test(pheno, binary, weights, **test_kw)
        return U, res, {"U_local": U, "nvlink_bytes": 0, "splitters": []}
    import torch
    import torch.distributed as dist

    mark("ingest")
    if route == "auto":
        route = "pages" if 9 <= int(k) <= 16 else "streams"
    if route == "pages" and not 9 <= int(k) <= 16:
        raise ValueError("route='pages' handles k = 9..16")
    stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
    with torch.cuda.stream(stream):
        parts = []            # per pass: (U_local, [PhenoResult])
        if route == "pages":
            pr = getattr(ka, "_page_route", None)
            if pr is None or (pr.rank, pr.world) != (rank, world):
                pr = ka._page_route = PageRoute(ka, rank, world, device)
            pr.prepare(int(k), int(n_samples), dist, torch)
            spl = pr.splitters
            mark("prepare")
            U_seen = 0
            exact = not (binary and test_kw.get("omit_b"))
            for j in range(pr.passes):
                lo, hi, spl_j = pr.pass_range(j)
                if pr.passes > 1:
                    ctx.set_range(lo, hi)
                ctx.route_setup(world, rank, spl_j, pr.pages)      # same sizes: nothing moves, new splitters
                ctx.route_begin()
                pr.barrier(dist)            # every pool is clean before anybody writes into it
                ctx.route_scatter()
                pr.barrier(dist)            # every sender's stores have landed
                U_local, ovf = ctx.route_build()
                ka.U = U_local
                u = torch.tensor([U_local, 1 if ovf else 0], dtype=torch.int64, device=device)
                dist.all_reduce(u)
                U_seen += int(u[0].item())
                if int(u[1].item()):
                    pr.pages *= 2          # same decision on every rank: larger pools from the next step on
                    if pr.passes > 1:
                        ctx.set_range(0, 0)
                    raise RuntimeError("a page pool overflowed while routing k-mers (skewed k-mer ranges); "
                                       "the next step will use pools twice as large")
                # Bonferroni needs the total U: a pass tests against the union seen so far (a laxer threshold),
                # the exact filter follows when every pass is done
                parts.append((U_local, ka.test(pheno, binary, weights, n_union_total=U_seen if exact else None, **test_kw)))
                mark(f"pass{j}")
            if pr.passes > 1:
                ctx.set_range(0, 0)
            U_total = U_seen
            nvl_bytes = None
        else:
            ctx.route_clear()
            nvl_bytes = exchange_streams(ka, n_samples, rank, world, device)
            mark("exchange")
            spl = ctx.sample_quantiles(0, world)          # identical on every rank: all hold sample 0
            U_local = ka.build(range_of(rank, spl))
            mark("build")
            u = torch.tensor([U_local], dtype=torch.int64, device=device)
            dist.all_reduce(u)
            U_total = int(u[0].item())
            parts.append((U_local, ka.test(pheno, binary, weights, n_union_total=U_total, **test_kw)))
            mark("test")
        merged = gather_results(parts, rank, world, device, dist, top_k=test_kw.get("top_k"),
                                p_exact=(float(test_kw.get("pvalue_cutoff", 0.05)) / U_total if U_total else 0.0)
                                if not (binary and test_kw.get("omit_b")) else None)
        mark("gather")
    if timing:
        import sys
        sys.stderr.write("[dist timing ms] " + " ".join(
            f"{b[0]}={1e3 * (b[1] - a[1]):.1f}" for a, b in zip(marks, marks[1:])) + "\n")
    return U_total, merged, {"U_local": sum(p[0] for p in parts), "nvlink_bytes": nvl_bytes, "splitters": spl, "route": route,
                             "passes": len(parts)}
