"""GenomeTester4 `.list` files: what `glistmaker` writes and
`glistquery` / `glistcompare` read (modeling.py:309-310, 326-327, 376-379).

Layout (little-endian): 40-byte header — magic "C4TG", u32 version 4, u32 minor 2, u32 word
length, u64 number of words, u64 total frequency, u64 offset of the first record (40) — then
12-byte records (u64 word, u32 count) in ascending word order. Lets users keep working with
the GenomeTester4 tools on lists counted on the GPU, and lets tests diff binary files.
"""
import struct

import numpy as np

_HDR = struct.Struct("<4sIIIQQQ")
_REC = np.dtype([("word", "<u8"), ("count", "<u4")])


def write_list(path, kmers, counts, k):
    kmers = np.asarray(kmers, dtype=np.uint64)
    counts = np.asarray(counts, dtype=np.uint32)
    if len(kmers) != len(counts):
        raise ValueError("kmers and counts differ in length")
    if len(kmers) > 1 and not (kmers[1:] > kmers[:-1]).all():
        raise ValueError("k-mers must be strictly ascending")
    rec = np.empty(len(kmers), dtype=_REC)
    rec["word"], rec["count"] = kmers, counts
    with open(path, "wb") as f:
        f.write(_HDR.pack(b"C4TG", 4, 2, int(k), len(kmers), int(counts.sum(dtype=np.uint64)), 40))
        f.write(rec.tobytes())


def read_list(path):
    """-> (kmers u64, counts u32, k)."""
    with open(path, "rb") as f:
        raw = f.read()
    magic, major, _minor, k, n, _total, start = _HDR.unpack_from(raw, 0)
    if magic != b"C4TG" or major != 4:
        raise ValueError(f"{path}: not a GenomeTester4 v4 list")
    rec = np.frombuffer(raw, dtype=_REC, count=n, offset=start)
    return rec["word"].copy(), rec["count"].copy(), int(k)


def sample_list_name(prefix, k):
    """`glistmaker ... -o PREFIX -w K` writes PREFIX_K.list (modeling.py:309-310)."""
    return f"{prefix}_{k}.list"
