"""The drop-in boundary wired into the REAL, unmodified reference CLI (this container only:
/root/reference is absent on the GPU box, where this test skips).

`modeling_gpu.install()` patches the five methods of modeling.py:1644-1686; here the GPU engine
behind it is replaced by an oracle-backed stand-in with the same interface, so the test checks the
wiring itself: argument plumbing, ML_df layout, column order, no_results, and that the unchanged
`get_ML_df` then writes byte-identical files. The CUDA engine is checked against the same golden
files in tests/test_outputs.py::test_gpu_path_reproduces_reference_files."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import GOLD, ROOT

def _ref_present():
    sys.path.insert(0, ROOT)
    from oracle import build
    return build.ref_python_root() is not None and build.ref_bin_dir() is not None


pytestmark = pytest.mark.skipif(not _ref_present(), reason="reference not present (neither /root/reference nor oracle/_ref)")

DRIVER = textwrap.dedent("""
    import sys, os, json
    import numpy as np
    sys.path.insert(0, {root!r})
    from oracle import ref_shim, kmers as ok, stats as ostats
    from phenotypeseeker_b200 import synth, modeling_gpu as mg
    from phenotypeseeker_b200.pipeline import PhenoResult

    class OracleEngine:                      # same interface as pipeline.KmerAssociation
        def __init__(self): self.k = None; self.U = 0
        def count(self, files, k, cutoff=1):
            self.k = k; self.lists = [ok.count_kmers(f, k, cutoff) for f in files]; self.n = len(files)
        def count_files(self, paths, k, cutoff=1):
            self.count([open(p, "rb").read() for p in paths], k, cutoff)
        def fits_in_one_build(self): return True
        def build(self):
            self.u = ok.union([l[0] for l in self.lists]); self.pres = ok.presence_matrix(self.u, self.lists)
            self.U = len(self.u); return self.U
        def test(self, pheno, binary, weights=None, min_samples=2, max_samples=None, pvalue_cutoff=0.05,
                 omit_b=False, n_union_total=None, pheno_names=None, top_k=None):
            ph = np.asarray(pheno, float).reshape(self.n, -1); out = []
            w = np.ones(self.n) if weights is None else np.asarray(weights, float)
            for j in range(ph.shape[1]):
                if binary:
                    r = ostats.chi2_rows(self.pres, np.where(np.isnan(ph[:, j]), -1, ph[:, j]).astype(np.int8), w, min_samples, max_samples)
                    thr = pvalue_cutoff if omit_b else pvalue_cutoff / self.U
                else:
                    r = ostats.welch_rows(self.pres, ph[:, j], w, min_samples, max_samples)
                    thr = pvalue_cutoff / self.U
                keep = r["tested"] & (r["p"] < thr); z = np.zeros(int(keep.sum()))
                out.append(PhenoResult(pheno_names[j], self.u[keep], np.nonzero(keep)[0].astype(np.uint64), r["stat"][keep],
                                       r["p"][keep], r["mean_x"][keep] if "mean_x" in r else z,
                                       r["mean_y"][keep] if "mean_y" in r else z, r["n_with"][keep], self.pres[keep]))
            return out

    case = json.load(open({case!r}))
    ds = synth.config(case["config"], tiny=True, n_samples=case["n_samples"], genome_len=case["genome_len"])
    ph, _ = ds.write("in")
    m = ref_shim.load_modeling()
    mg._STATE["ka"] = OracleEngine(); mg._STATE["pid"] = os.getpid()
    mg.install(m)
    ref_shim.run_cli(["modeling", ph] + case["args"], os.getcwd())
""")


@pytest.mark.parametrize("tag", ["chi2", "ttest"])
def test_patched_reference_cli_writes_identical_files(tag, tmp_path):
    gold_dir = os.path.join(GOLD, f"cli_{tag}")
    case_path = os.path.join(gold_dir, "case.json")
    with open(case_path) as f:
        case = json.load(f)
    script = tmp_path / "drive.py"
    script.write_text(DRIVER.format(root=ROOT, case=case_path))
    r = subprocess.run([sys.executable, str(script)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    for fn in case["files"]:
        p = tmp_path / fn
        assert p.exists(), (fn, r.stderr[-2000:])
        with open(os.path.join(gold_dir, fn), "rb") as f:
            assert p.read_bytes() == f.read(), fn
    assert "took" in (tmp_path / "log.txt").read_text()
    assert not (tmp_path / "K-mer_lists").exists() or not any(
        fn.endswith(".list") for fn in os.listdir(tmp_path / "K-mer_lists"))   # no glistmaker ran


GPU_DRIVER = textwrap.dedent("""
    import sys, os, json
    sys.path.insert(0, {root!r})
    from oracle import ref_shim
    from phenotypeseeker_b200 import synth, modeling_gpu as mg
    case = json.load(open({case!r}))
    ds = synth.config(case["config"], tiny=True, n_samples=case["n_samples"], genome_len=case["genome_len"])
    ph, _ = ds.write("in")
    m = ref_shim.load_modeling()
    mg.install(m)                              # the real CUDA engine behind the reference's own methods
    ref_shim.run_cli(["modeling", ph] + case["args"], os.getcwd())
    print("LAUNCHES", mg._ka().ctx.launch_count())
""")


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["chi2", "ttest"])
def test_unmodified_cli_with_cuda_engine_writes_identical_files(tag, tmp_path):
    """The whole drop-in on the GPU box: the reference's unmodified CLI and orchestrator (installed under
    oracle/_ref by oracle/build.py), `modeling_gpu.install()`, libpskmer.so doing stages 1-3 — and the files
    `get_ML_df` writes must equal, byte for byte, what the all-CPU reference wrote (tests/golden/cli_*)."""
    gold_dir = os.path.join(GOLD, f"cli_{tag}")
    case_path = os.path.join(gold_dir, "case.json")
    with open(case_path) as f:
        case = json.load(f)
    script = tmp_path / "drive.py"
    script.write_text(GPU_DRIVER.format(root=ROOT, case=case_path))
    r = subprocess.run([sys.executable, str(script)], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert "LAUNCHES" in r.stdout and int(r.stdout.split("LAUNCHES")[1].split()[0]) > 0, r.stderr[-2000:]
    for fn in case["files"]:
        p = tmp_path / fn
        assert p.exists(), (fn, r.stderr[-2000:])
        with open(os.path.join(gold_dir, fn), "rb") as f:
            assert p.read_bytes() == f.read(), fn
