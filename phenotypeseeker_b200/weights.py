"""Mash/GSC sample weights (`-w`): host-side replacement for `Samples.get_mash_sketches` /
`Samples.get_weights` (modeling.py:386-503), for installations without the `mash` binary,
Biopython and ete3. SURVEY.md §8f row 2. The weights are an INPUT VECTOR of the GPU test kernels
(modeling.py:812-822, 734-736); computing them is O(N L) sketching + O(N^2) distances + O(N^3)
neighbour joining on the host and is outside the timed hot path.

What the reference does, step by step, and what is restated here:

  mash sketch -r <sample>         -> sketch(): the 1000 smallest distinct 64-bit hashes
      (mash 2.2: k = 21, MurmurHash3_x64_128 seed 42, first 8 bytes, canonical = the
      alphabetically smaller of a k-mer and its reverse complement, k-mers with a letter outside
      ACGT skipped, case folded).                      PINNED against the shipped binary
      (`mash info -d`), tests/golden/mash.json.
  mash paste + mash dist          -> mash_distance(): merge walk over the two sorted sketches up
      to 1000 union elements, j = shared/denom, d = -ln(2j/(1+j))/k, printed by mash with 6
      significant digits (the reference parses that text, modeling.py:421).        PINNED (same).
  Bio.Phylo DistanceTreeConstructor.nj + phyloxml -> newick (modeling.py:446-458)
                                  -> neighbor_joining(), newick(): restated from the published
      algorithm of Biopython 1.76 (`TreeConstruction.py`), including its tie-breaking order, its
      rooting (the last two nodes are joined by hanging one under the other) and the Newick
      writer's "%1.5f" branch lengths. Biopython is not installed here and not vendored by the
      reference: PARITY UNPINNED for this step.
  GSC weights on the ete3 tree (modeling.py:460-503) -> gsc_weights(): clip branch lengths to
      [1e-9, 1e9], post-order branch sums, pre-order weights w_child = w_parent (dist +
      BranchSum) / parent.BranchSum, leaves x N.        PINNED against the reference's own
      methods run on a minimal ete3 stand-in (tests/test_weights.py).

Reference quirk kept: `mash paste reference.msh K-mer_lists/*.msh` orders the sketches by shell
glob (sorted file names) while `_mash_output_to_distance_matrix` labels the rows with the
data.pheno order (modeling.py:414-428) — the two differ when data.pheno is not sorted by sample
name. `gsc_weights_for_samples(..., keep_glob_quirk=True)` reproduces that; the default labels
the distances correctly.
"""
import math

import numpy as np

MASH_K = 21
MASH_S = 1000
MASH_SEED = 42

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
_C1 = np.uint64(0x87C37B91114253D5)
_C2 = np.uint64(0x4CF5AD432745937F)


def _rotl(x, r):
    return (x << np.uint64(r)) | (x >> np.uint64(64 - r))


def _fmix(k):
    k = k ^ (k >> np.uint64(33))
    k = k * np.uint64(0xFF51AFD7ED558CCD)
    k = k ^ (k >> np.uint64(33))
    k = k * np.uint64(0xC4CEB9FE1A85EC53)
    k = k ^ (k >> np.uint64(33))
    return k


def murmur3_x64_128_h1(data, seed=MASH_SEED):
    """First 64 bits of MurmurHash3_x64_128 of every row of `data` (n x len uint8), vectorised.
    (Appleby's public-domain algorithm; mash hashes the k-mer string with it, Sketch.cpp getHash.)"""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    n, length = data.shape
    h1 = np.full(n, seed, dtype=np.uint64)
    h2 = np.full(n, seed, dtype=np.uint64)
    with np.errstate(over="ignore"):
        def le64(cols):
            v = np.zeros(n, dtype=np.uint64)
            for j in range(cols.shape[1]):
                v |= cols[:, j].astype(np.uint64) << np.uint64(8 * j)
            return v
        nblocks = length // 16
        for b in range(nblocks):
            k1 = le64(data[:, 16 * b:16 * b + 8])
            k2 = le64(data[:, 16 * b + 8:16 * b + 16])
            k1 = k1 * _C1; k1 = _rotl(k1, 31); k1 = k1 * _C2; h1 = h1 ^ k1
            h1 = _rotl(h1, 27); h1 = h1 + h2; h1 = h1 * np.uint64(5) + np.uint64(0x52DCE729)
            k2 = k2 * _C2; k2 = _rotl(k2, 33); k2 = k2 * _C1; h2 = h2 ^ k2
            h2 = _rotl(h2, 31); h2 = h2 + h1; h2 = h2 * np.uint64(5) + np.uint64(0x38495AB5)
        tail = data[:, 16 * nblocks:]
        t = tail.shape[1]
        if t > 8:
            k2 = le64(tail[:, 8:])
            k2 = k2 * _C2; k2 = _rotl(k2, 33); k2 = k2 * _C1; h2 = h2 ^ k2
        if t > 0:
            k1 = le64(tail[:, :min(t, 8)])
            k1 = k1 * _C1; k1 = _rotl(k1, 31); k1 = k1 * _C2; h1 = h1 ^ k1
        h1 = h1 ^ np.uint64(length); h2 = h2 ^ np.uint64(length)
        h1 = h1 + h2; h2 = h2 + h1
        h1 = _fmix(h1); h2 = _fmix(h2)
        h1 = h1 + h2
    return h1


_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b
_VALID = np.zeros(256, dtype=bool)
_VALID[list(b"ACGT")] = True


def _records(text):
    """FASTA / FASTQ bytes -> list of upper-cased sequence byte arrays (one per record)."""
    text = bytes(text)
    recs = []
    if text[:1] == b"@":
        lines = text.split(b"\n")
        for i in range(1, len(lines), 4):
            recs.append(lines[i].strip())
    else:
        for chunk in text.split(b">")[1:]:
            nl = chunk.find(b"\n")
            recs.append(b"" if nl < 0 else chunk[nl + 1:].replace(b"\n", b"").replace(b"\r", b""))
    return [np.frombuffer(r.upper(), dtype=np.uint8) for r in recs if len(r)]


def sketch(text, k=MASH_K, s=MASH_S, seed=MASH_SEED, chunk=1 << 20):
    """`mash sketch -r -k k -s s -S seed`: ascending array of the <= s smallest distinct hashes."""
    best = np.empty(0, dtype=np.uint64)
    for seq in _records(text):
        if len(seq) < k:
            continue
        for c0 in range(0, len(seq) - k + 1, chunk):
            part = seq[c0:c0 + chunk + k - 1]
            win = np.lib.stride_tricks.sliding_window_view(part, k)
            ok = _VALID[part]
            bad = np.concatenate(([0], np.cumsum(~ok)))
            good = (bad[k:] - bad[:-k]) == 0                     # windows made of ACGT only
            win = win[good]
            if not len(win):
                continue
            rc = _COMP[win[:, ::-1]]
            # alphabetical minimum of forward / reverse complement (first differing letter decides)
            diff = win != rc
            first = diff.argmax(axis=1)
            rows = np.arange(len(win))
            use_rc = diff.any(axis=1) & (rc[rows, first] < win[rows, first])
            canon = np.where(use_rc[:, None], rc, win)
            h = np.unique(murmur3_x64_128_h1(canon, seed))
            best = np.union1d(best, h[:s])[:s]
    return best


def mash_distance(a, b, k=MASH_K, s=MASH_S):
    """`mash dist` on two sketches (ascending hash arrays) -> (distance, shared, denom)
    (mash 2.2 CommandDistance.cpp compareSketches)."""
    i = j = common = denom = 0
    na, nb = len(a), len(b)
    while denom < s and i < na and j < nb:
        if a[i] < b[j]:
            i += 1
        elif b[j] < a[i]:
            j += 1
        else:
            i += 1; j += 1; common += 1
        denom += 1
    if denom < s:
        if i < na:
            denom += na - i
        if j < nb:
            denom += nb - j
        if denom > s:
            denom = s
    if denom == 0 or common == denom:
        d = 0.0
    elif common == 0:
        d = 1.0
    else:
        jac = common / denom
        d = -math.log(2.0 * jac / (1.0 + jac)) / k
    return d, common, denom


def printed(x):
    """The value the reference reads back: mash prints distances with 6 significant digits
    (C++ ostream default) and modeling.py:421 parses that text."""
    return float("%g" % x)


def distance_matrix(sketches, k=MASH_K, s=MASH_S):
    n = len(sketches)
    dm = np.zeros((n, n))
    for i in range(n):
        for j in range(i):
            dm[i, j] = dm[j, i] = printed(mash_distance(sketches[i], sketches[j], k, s)[0])
    return dm


class Node:
    __slots__ = ("name", "dist", "children", "up", "BranchSum", "NodeWeight")

    def __init__(self, name=None, dist=0.0):
        self.name, self.dist, self.children, self.up = name, dist, [], None
        self.BranchSum = 0.0
        self.NodeWeight = 0.0

    def add(self, child):
        child.up = self
        self.children.append(child)


def neighbor_joining(names, dm):
    """Biopython 1.76 DistanceTreeConstructor.nj restated (see module docstring: unpinned).
    names: N labels; dm: N x N symmetric. Returns the root Node (branch lengths in .dist)."""
    dm = [list(map(float, row)) for row in np.asarray(dm, dtype=np.float64)]
    clades = [Node(nm) for nm in names]
    n = len(clades)
    if n == 1:
        return clades[0]
    if n == 2:
        c1, c2 = clades[1], clades[0]
        c1.dist = dm[1][0] / 2.0
        c2.dist = dm[1][0] - c1.dist
        root = Node("Inner")
        root.add(c1); root.add(c2)
        return root
    inner = None
    count = 0
    while len(dm) > 2:
        m = len(dm)
        nd = [sum(dm[i]) / (m - 2) for i in range(m)]
        min_dist = dm[1][0] - nd[1] - nd[0]
        mi, mj = 0, 1
        for i in range(1, m):
            for j in range(i):
                t = dm[i][j] - nd[i] - nd[j]
                if min_dist > t:
                    min_dist, mi, mj = t, i, j
        c1, c2 = clades[mi], clades[mj]
        count += 1
        inner = Node("Inner" + str(count))
        inner.add(c1); inner.add(c2)
        c1.dist = (dm[mi][mj] + nd[mi] - nd[mj]) / 2.0
        c2.dist = dm[mi][mj] - c1.dist
        clades[mj] = inner
        del clades[mi]
        for x in range(m):
            if x != mi and x != mj:
                dm[mj][x] = dm[x][mj] = (dm[mi][x] + dm[mj][x] - dm[mi][mj]) / 2.0
        del dm[mi]
        for row in dm:
            del row[mi]
    if clades[0] is inner:
        clades[0].dist = 0.0
        clades[1].dist = dm[1][0]
        clades[0].add(clades[1])
        root = clades[0]
    else:
        clades[0].dist = dm[1][0]
        clades[1].dist = 0.0
        clades[1].add(clades[0])
        root = clades[1]
    return root


def newick(node):
    """Bio.Phylo's Newick writer: names + "%1.5f" branch lengths (the reference converts the
    phyloxml tree to Newick before handing it to ete3, modeling.py:454-458)."""
    label = (node.name or "") + ":%1.5f" % node.dist
    if node.children:
        return "(" + ",".join(newick(c) for c in node.children) + ")" + label
    return label


def parse_newick(text):
    """Minimal Newick reader (names + branch lengths, ete3 format=1 subset) -> root Node."""
    text = text.strip().rstrip(";")
    pos = 0

    def parse():
        nonlocal pos
        node = Node()
        if text[pos] == "(":
            pos += 1
            while True:
                node.add(parse())
                if text[pos] == ",":
                    pos += 1
                    continue
                pos += 1          # ')'
                break
        start = pos
        while pos < len(text) and text[pos] not in ",():":
            pos += 1
        node.name = text[start:pos]
        if pos < len(text) and text[pos] == ":":
            pos += 1
            start = pos
            while pos < len(text) and text[pos] not in ",()":
                pos += 1
            node.dist = float(text[start:pos])
        return node

    return parse()


def gsc_weights(root, normalize="mean1", min_val=1e-9, max_val=1e9):
    """GSC_weights_from_newick + clip_branch_lengths + set_branch_sum + set_node_weight
    (modeling.py:460-503) on a Node tree -> {leaf name: weight}. ete3 gives a root without a branch
    length dist = 0.0, which the clip raises to 1e-9 (unused by the sums)."""
    order, stack = [], [root]
    while stack:                                      # pre-order
        nd = stack.pop()
        order.append(nd)
        stack.extend(reversed(nd.children))
    for nd in order:
        if nd.dist > max_val:
            nd.dist = max_val
        elif nd.dist < min_val:
            nd.dist = min_val
    for nd in reversed(order):                        # children before parents
        total = 0.0
        for ch in nd.children:
            total += ch.BranchSum
            total += ch.dist
        nd.BranchSum = total
    weights = {}
    for nd in order:                                  # parents before children
        if nd.up is None:
            nd.NodeWeight = 1.0
        else:
            nd.NodeWeight = nd.up.NodeWeight * (nd.dist + nd.BranchSum) / nd.up.BranchSum
        if not nd.children:
            weights[nd.name] = nd.NodeWeight
    if normalize == "mean1":
        weights = {k: v * len(weights) for k, v in weights.items()}
    return weights


def gsc_weights_for_samples(names, texts, keep_glob_quirk=False, k=MASH_K, s=MASH_S):
    """names / texts in data.pheno order -> weights in the same order (what `get_weights` leaves in
    Input.samples[name].weight, modeling.py:392-402)."""
    sk = {nm: sketch(t, k, s) for nm, t in zip(names, texts)}
    # shell glob order of K-mer_lists/<name>.msh in the C locale (other locales collate differently)
    order = sorted(names, key=lambda nm: nm + ".msh") if keep_glob_quirk else list(names)
    dm = distance_matrix([sk[nm] for nm in order], k, s)
    root = neighbor_joining(list(names), dm)          # rows are labelled in data.pheno order either way
    w = gsc_weights(parse_newick(newick(root) + ";"))
    return np.array([w[nm] for nm in names], dtype=np.float64)
