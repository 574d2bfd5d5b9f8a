"""oracle/kmers.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy/ctypes face of oracle/kmer_oracle.c: per-sample canonical k-mer counts
(glistmaker, modeling.py:303-315), union (glistcompare -u, :351-380), mapping
(glistquery -l, :317-348) and the presence matrix the reference only ever
materialises as text stripes (:677-695).
"""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build_oracle()
        L = ctypes.CDLL(path)
        L.orc_decode.restype = ctypes.c_size_t
        L.orc_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p,
                                 ctypes.POINTER(ctypes.c_int)]
        L.orc_count.restype = ctypes.c_int
        L.orc_count.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int,
                                ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
        L.orc_free.argtypes = [ctypes.c_void_p]
        L.orc_free.restype = None
        _lib = L
    return _lib


def decode(data: bytes):
    """-> (codes u8 array [0..3 base, 4 break], fmt 0/1/2)."""
    L = lib()
    out = np.empty(len(data) + 1, dtype=np.uint8)
    fmt = ctypes.c_int(0)
    m = L.orc_decode(data, len(data), out.ctypes.data, ctypes.byref(fmt))
    return out[:m].copy(), fmt.value


def count_kmers(data: bytes, k: int, cutoff: int = 1):
    """glistmaker restated -> (kmers u64 ascending, counts u32).

    cutoff: keep k-mers with count >= cutoff (documented intent of `-c`; the
    shipped glistmaker 4.2.3 ignores it — SURVEY.md Appendix A4/B2 — so parity
    with the binary holds at cutoff=1 only).
    """
    L = lib()
    pk, pc = ctypes.c_void_p(), ctypes.c_void_p()
    nu, nt = ctypes.c_size_t(), ctypes.c_size_t()
    rc = L.orc_count(data, len(data), k, ctypes.byref(pk), ctypes.byref(pc),
                     ctypes.byref(nu), ctypes.byref(nt))
    if rc != 0:
        raise ValueError("orc_count failed (k out of range?)")
    n = nu.value
    kmers = np.ctypeslib.as_array(ctypes.cast(pk, ctypes.POINTER(ctypes.c_uint64)), (max(n, 1),))[:n].copy()
    counts = np.ctypeslib.as_array(ctypes.cast(pc, ctypes.POINTER(ctypes.c_uint32)), (max(n, 1),))[:n].copy()
    L.orc_free(pk)
    L.orc_free(pc)
    if cutoff > 1:
        keep = counts >= cutoff
        kmers, counts = kmers[keep], counts[keep]
    return kmers, counts


def union(lists):
    """glistcompare -u over all samples -> sorted distinct u64 array."""
    if not lists:
        return np.empty(0, dtype=np.uint64)
    return sorted_unique(np.concatenate(lists))


def sorted_unique(a):
    """Ascending distinct values (sort + neighbour compare; numpy 2.3's hash-based np.unique is ~70x slower
    on tens of millions of u64)."""
    a = np.sort(np.asarray(a, dtype=np.uint64))
    if len(a) == 0:
        return a
    keep = np.empty(len(a), dtype=bool)
    keep[0] = True
    np.not_equal(a[1:], a[:-1], out=keep[1:])
    return a[keep]


def count_many(texts, k, cutoff=1, threads=None):
    """count_kmers over many samples on `threads` host threads (the C oracle runs outside the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    lib()
    with ThreadPoolExecutor(threads or os.cpu_count() or 1) as ex:
        return list(ex.map(lambda t: count_kmers(t, k, cutoff), texts))


def union_and_rows(kmer_lists, threads=None, chunk=24_000_000):
    """glistcompare -u + glistquery -l for MANY samples at full size: (union u64 ascending, rows U x W uint32
    with sample s at bit s % 32 of word s // 32, W = ceil(N/32) rounded up to 4 like the GPU's rows).
    Same result as union() + presence_matrix() + pack_rows(); the k-mer space is cut into ranges of about
    `chunk` instances that are merged independently on host threads (numpy's sort releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    N = len(kmer_lists)
    W = (((N + 31) // 32) + 3) // 4 * 4
    total = sum(len(l) for l in kmer_lists)
    if total == 0:
        return np.empty(0, np.uint64), np.zeros((0, W), np.uint32)
    R = max(1, -(-total // chunk))
    big = max(kmer_lists, key=len)
    cuts = [big[len(big) * r // R] for r in range(1, R)]           # quantiles of the largest list
    bounds = [[0] + [int(np.searchsorted(l, c)) for c in cuts] + [len(l)] for l in kmer_lists]

    def one(r):
        parts = [l[b[r]:b[r + 1]] for l, b in zip(kmer_lists, bounds)]
        u = sorted_unique(np.concatenate(parts))
        rows = np.zeros((len(u), W), dtype=np.uint32)
        for s, p in enumerate(parts):
            if len(p):
                rows[np.searchsorted(u, p), s >> 5] |= np.uint32(1 << (s & 31))
        return u, rows

    with ThreadPoolExecutor(threads or os.cpu_count() or 1) as ex:
        out = list(ex.map(one, range(R)))
    return np.concatenate([o[0] for o in out]), np.concatenate([o[1] for o in out])


def map_counts(u, kmers, counts):
    """glistquery sample.list -l union.list -> count per union k-mer (0 if absent)."""
    idx = np.searchsorted(kmers, u)
    idx_c = np.minimum(idx, max(len(kmers) - 1, 0))
    if len(kmers) == 0:
        return np.zeros(len(u), dtype=np.uint32)
    hit = kmers[idx_c] == u
    return np.where(hit, counts[idx_c], 0).astype(np.uint32)


def presence_matrix(u, sample_lists):
    """U x N uint8 presence (count>0) matrix, sample order = list order."""
    m = np.zeros((len(u), len(sample_lists)), dtype=np.uint8)
    for s, (km, ct) in enumerate(sample_lists):
        m[:, s] = map_counts(u, km, ct) > 0
    return m


def pack_rows(presence):
    """U x N 0/1 -> U x ceil(N/32) uint32, sample s at bit s%32 of word s//32."""
    U, N = presence.shape
    W = (N + 31) // 32
    pad = np.zeros((U, W * 32), dtype=np.uint8)
    pad[:, :N] = presence
    bits = pad.reshape(U, W, 32).astype(np.uint32)
    return (bits << np.arange(32, dtype=np.uint32)).sum(axis=2, dtype=np.uint64).astype(np.uint32)


_B = np.frombuffer(b"ACGT", dtype=np.uint8)


def kmer_to_str(x: int, k: int) -> str:
    return "".join("ACGT"[(int(x) >> (2 * (k - 1 - i))) & 3] for i in range(k))


def str_to_kmer(s: str) -> int:
    v = 0
    for ch in s:
        v = (v << 2) | "ACGT".index(ch)
    return v


# ---------------------------------------------------------------------------
# The shipped binaries (this container, or oracle/_ref/bin on the GPU box)

def run_glistmaker(data: bytes, k: int, suffix=".fa"):
    """Run the real glistmaker|glistquery on `data`; -> (kmers u64, counts u32)."""
    bindir = _build.ref_bin_dir()
    if bindir is None:
        raise RuntimeError("reference binaries not available")
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "in" + suffix)
        with open(p, "wb") as f:
            f.write(data)
        subprocess.run([os.path.join(bindir, "glistmaker"), p, "-o", os.path.join(td, "o"), "-w", str(k)],
                       check=False, capture_output=True)
        lst = os.path.join(td, f"o_{k}.list")
        if not os.path.exists(lst):
            return np.empty(0, np.uint64), np.empty(0, np.uint32)
        txt = subprocess.run([os.path.join(bindir, "glistquery"), lst], capture_output=True, text=True).stdout
    ks, cs = [], []
    for line in txt.splitlines():
        a, b = line.split("\t")
        ks.append(str_to_kmer(a))
        cs.append(int(b))
    return np.array(ks, dtype=np.uint64), np.array(cs, dtype=np.uint32)
