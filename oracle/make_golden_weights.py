"""oracle/make_golden_weights.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Golden vectors for phenotypeseeker_b200/weights.py (the `-w` Mash/GSC weights, modeling.py:386-503),
generated from the REAL reference in this container:

  tests/golden/mash.json   small FASTA/FASTQ inputs -> min-hash sketches dumped by the shipped
                           `mash sketch -r` / `mash info -d` (mash 2.2), and the `mash dist` table
                           of `mash paste` of all of them (distance text, shared/denominator)
  tests/golden/gsc.json    Newick trees -> weights returned by the reference's own
                           `Samples.GSC_weights_from_newick` (+ clip_branch_lengths, set_branch_sum,
                           set_node_weight) run on a minimal stand-in for ete3's Tree (ete3 is not
                           installed here; the stand-in only parses Newick and links nodes — the
                           arithmetic and the traversals are the reference's)

Biopython's neighbour joining cannot be run here (not installed, not vendored): no golden for it.

Run:  python -m oracle.make_golden_weights      (needs /root/reference; not needed at test time)
"""
import base64
import json
import os
import subprocess
import tempfile

import numpy as np

from . import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
MASH = os.path.join(ref_shim.REF_ROOT, "bin", "mash")


def _genome(rng, n, gc=0.29):
    p = [(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2]
    return rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n, p=p)


def _fasta(seqs, width=60, lower=False):
    out = bytearray()
    for i, s in enumerate(seqs):
        out += b">rec%d test\n" % i
        b = bytes(s)
        if lower:
            b = b.lower()
        for j in range(0, len(b), width):
            out += b[j:j + width] + b"\n"
    return bytes(out)


def mash_cases():
    rng = np.random.default_rng(20260117)
    anc = _genome(rng, 5000)
    cases = {}
    for i in range(4):                                   # related genomes: shared hashes
        g = anc.copy()
        pos = rng.integers(0, len(g), 40 + 25 * i)
        g[pos] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=len(pos))
        cases[f"rel{i}"] = _fasta([g[:2600], g[2600:]])
    g = anc.copy()
    g[rng.integers(0, len(g), 30)] = ord("N")            # windows with N are skipped
    cases["withN"] = _fasta([g])
    cases["lower"] = _fasta([anc[:3000]], lower=True)    # case is folded
    cases["tiny"] = _fasta([anc[:150], anc[10:15], anc[300:340]])   # < 1000 hashes, one record < k
    cases["other"] = _fasta([_genome(rng, 2500)])        # unrelated: few or no shared hashes
    fq = bytearray()
    for r in range(60):
        st = int(rng.integers(0, 4900))
        read = bytes(anc[st:st + 100])
        fq += b"@r%d\n" % r + read + b"\n+\n" + b"I" * len(read) + b"\n"
    cases["reads"] = bytes(fq)
    return cases


def make_mash():
    cases = mash_cases()
    out = {"k": 21, "s": 1000, "seed": 42, "cases": [], "dist": []}
    with tempfile.TemporaryDirectory() as td:
        names = sorted(cases)
        for nm in names:
            ext = ".fq" if cases[nm][:1] == b"@" else ".fa"
            path = os.path.join(td, nm + ext)
            with open(path, "wb") as f:
                f.write(cases[nm])
            subprocess.run([MASH, "sketch", "-r", path, "-o", os.path.join(td, nm)], check=True,
                           capture_output=True)
            info = json.loads(subprocess.run([MASH, "info", "-d", os.path.join(td, nm + ".msh")], check=True,
                                             capture_output=True, text=True).stdout)
            out["cases"].append({"name": nm, "data_b64": base64.b64encode(cases[nm]).decode(),
                                 "hashes": [str(h) for h in info["sketches"][0]["hashes"]]})
        subprocess.run([MASH, "paste", os.path.join(td, "all")] + [os.path.join(td, nm + ".msh") for nm in names],
                       check=True, capture_output=True)
        txt = subprocess.run([MASH, "dist", os.path.join(td, "all.msh"), os.path.join(td, "all.msh")], check=True,
                             capture_output=True, text=True).stdout
        for line in txt.strip().split("\n"):
            a, b, d, _, frac = line.split("\t")
            out["dist"].append([os.path.basename(a).rsplit(".", 1)[0], os.path.basename(b).rsplit(".", 1)[0], d, frac])
    with open(os.path.join(GOLD, "mash.json"), "w") as f:
        json.dump(out, f)
    print("mash.json:", len(out["cases"]), "sketches,", len(out["dist"]), "distances")


class StubTree:
    """The slice of ete3.Tree the reference's GSC code touches (modeling.py:464-503)."""

    def __init__(self, newick=None, format=1):
        self.name, self.dist, self.children, self.up = "", 0.0, [], None
        if newick is not None:
            from phenotypeseeker_b200.weights import parse_newick
            text = open(newick).read() if os.path.exists(newick) else newick
            self._adopt(parse_newick(text))

    def _adopt(self, node):
        self.name, self.dist = node.name, node.dist
        for ch in node.children:
            c = StubTree()
            c._adopt(ch)
            c.up = self
            self.children.append(c)

    def get_children(self):
        return self.children

    def traverse(self, strategy="levelorder"):
        queue = [self]
        while queue:
            nd = queue.pop(0)
            yield nd
            queue.extend(nd.children)

    def iter_leaves(self):
        for nd in self.traverse():
            if not nd.children:
                yield nd


def make_gsc():
    from phenotypeseeker_b200 import weights as W
    m = ref_shim.load_modeling()
    m.Tree = StubTree
    rng = np.random.default_rng(7)
    trees = ["((a:0.10000,b:0.20000)Inner1:0.05000,c:0.30000)Inner2:0.00000;",
             "((a:0.00000,b:0.00000)Inner1:0.00000,c:0.00100)Inner2:0.00000;",       # zero lengths -> clipped to 1e-9
             "(((s1:0.00120,s2:0.00080)Inner1:0.00045,s3:0.00210)Inner2:0.00000,s4:0.00300)Inner3:0.00000;"]
    for n in (5, 9, 20, 64):
        pts = rng.random((n, 3))
        dm = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)) * 0.01
        names = [f"S{i:03d}" for i in range(n)]
        trees.append(W.newick(W.neighbor_joining(names, dm)) + ";")
    out = []
    with tempfile.TemporaryDirectory() as td:
        for t in trees:
            path = os.path.join(td, "tree_newick.txt")
            with open(path, "w") as f:
                f.write(t)
            for norm in ("mean1", "sum1"):
                w = m.Samples.GSC_weights_from_newick(path, normalize=norm)
                out.append({"newick": t, "normalize": norm, "weights": {k: float(v) for k, v in w.items()}})
    with open(os.path.join(GOLD, "gsc.json"), "w") as f:
        json.dump(out, f)
    print("gsc.json:", len(out), "cases")


if __name__ == "__main__":
    make_mash()
    make_gsc()
