"""--kmerDB (modeling.py:361-372): the host logic that cuts the union down to the k-mers of a database.
Runs on CPU against a stand-in context (numpy arrays behind the four calls restrict_to makes); the calls
themselves — ps_get_union, ps_get_rows, ps_load_matrix — are exercised on the GPU by tests/test_gpu_parity.py."""
import numpy as np

from oracle import kmers as ok
from oracle.kmers import pack_rows
from phenotypeseeker_b200.pipeline import KmerAssociation, unpack_rows
from phenotypeseeker_b200 import synth


class StandInContext:
    def __init__(self, union, rows, n_samples):
        self.u, self.rows, self.n = union, rows, n_samples
        self.U = len(union)
        self.loaded = None

    def row_words(self):
        return self.rows.shape[1]

    def get_union(self, first=0, count=None):
        return self.u[first:first + (self.U - first if count is None else count)].copy()

    def get_rows(self, first=0, count=None):
        return self.rows[first:first + (self.U - first if count is None else count)].copy()

    def load_matrix(self, rows, kmers=None):
        self.loaded = (np.array(rows), np.array(kmers))
        self.u, self.rows, self.U = np.array(kmers), np.array(rows), len(kmers)


def test_restrict_to_equals_glistcompare_intersection():
    ds = synth.config(0, tiny=True)
    k = 13
    lists = [ok.count_kmers(f, k) for f in ds.files]
    u = ok.union([l[0] for l in lists])
    pres = ok.presence_matrix(u, lists)
    db_text = ds.files[0][:20000] + b">extra\nACGTACGTTTGACCAGTAGGATCCAAGT\n"
    db = ok.count_kmers(db_text, k)[0]
    ctx = StandInContext(u, pack_rows(pres), ds.n_samples)
    ka = KmerAssociation(ctx=ctx)
    ka.U = len(u)
    new_u = ka.restrict_to(db, chunk=1000)                   # several row chunks
    expect = np.intersect1d(u, db)
    assert 0 < new_u == len(expect) < len(u)
    rows, kmers = ctx.loaded
    assert np.array_equal(kmers, expect)
    assert np.array_equal(unpack_rows(rows, ds.n_samples), pres[np.isin(u, db)])
    # empty database / disjoint database -> empty feature vector (the reference then stops: nothing left to test)
    ctx2 = StandInContext(u, pack_rows(pres), ds.n_samples)
    ka2 = KmerAssociation(ctx=ctx2)
    assert ka2.restrict_to(np.empty(0, np.uint64)) == 0 and ctx2.loaded[0].shape == (0, ctx2.row_words())
