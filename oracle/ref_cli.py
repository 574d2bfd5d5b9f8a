"""oracle/ref_cli.py — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

Runs the reference's own, UNMODIFIED `phenotypeseeker modeling` CLI (modeling.py:1625-1709, through
oracle/ref_shim.py) up to the end of its hot path — k-mer lists, feature vector, mapping, per-k-mer
tests — and reports U, the survivors per phenotype and the wall-clock of exactly that stretch. The
run is stopped where `modeling()` leaves the hot path (`Input.pop_phenos_out_of_kmers`, :1686); the
sklearn stage does not run. bench.py's CPU reference arm calls this in a subprocess per step.

  python -m oracle.ref_cli <data.pheno> <workdir> <threads> [--weights w.json] [extra CLI args ...]

The result goes to <workdir>/refcli.json (the reference forks Manager processes that outlive the run and
keep inherited pipes open, so callers must not wait on stdout: use `run()` below).

--weights: sample weights to inject instead of the reference's Mash/GSC computation (mash output needs
Biopython + ete3, which this image lacks): SURVEY.md 8d allows the same synthetic weights on both arms.
For continuous phenotypes `statsmodels.ttest_ind` is the restatement of oracle/stats.py (see ref_shim).
"""
import json
import os
import sys
import time


class _HotPathDone(BaseException):
    pass


def main(argv):
    pheno, workdir, threads = argv[0], argv[1], int(argv[2])
    rest = argv[3:]
    weights = None
    if rest and rest[0] == "--weights":
        with open(rest[1]) as f:
            weights = json.load(f)
        rest = rest[2:]
    from . import ref_shim
    m = ref_shim.load_modeling()
    t = {}

    def stop(cls=None):
        t["end"] = time.time()
        raise _HotPathDone()

    m.Input.pop_phenos_out_of_kmers = classmethod(stop)
    if weights is not None:
        def get_weights(cls):
            for s, w in zip(m.Input.samples.values(), weights):
                s.weight = float(w)
        m.Samples.get_mash_sketches = lambda self: None
        m.Samples.get_weights = classmethod(get_weights)
        rest = ["-w"] + rest
    os.makedirs(workdir, exist_ok=True)
    import runpy
    old = os.getcwd()
    os.chdir(workdir)
    sys.argv = ["phenotypeseeker", "modeling", pheno, "-nt", str(threads)] + rest
    t["start"] = time.time()
    try:
        runpy.run_path(os.path.join(ref_shim.REF_ROOT, "scripts", "phenotypeseeker"), run_name="__main__")
    except _HotPathDone:
        pass
    finally:
        os.chdir(old)
    out = {"U": int(m.phenotypes.no_kmers_to_analyse), "seconds": t["end"] - t["start"], "threads": threads,
           "survivors": {name: int(p.ML_df.shape[1]) for name, p in m.Input.phenotypes_to_analyse.items()}}
    with open(os.path.join(workdir, "refcli.json"), "w") as f:
        json.dump(out, f)
    os._exit(0)          # the reference leaves Manager processes behind


def run(pheno, workdir, threads, extra_args=(), weights=None, timeout=3600):
    """Run main() in a subprocess of its own session, wait for IT (not for its pipes), then remove the
    Manager processes the reference left behind. -> dict(U, seconds, threads, survivors)."""
    import signal
    import subprocess
    os.makedirs(workdir, exist_ok=True)
    cmd = [sys.executable, "-m", "oracle.ref_cli", pheno, workdir, str(threads)]
    if weights is not None:
        wp = os.path.join(workdir, "weights.json")
        with open(wp, "w") as f:
            json.dump([float(w) for w in weights], f)
        cmd += ["--weights", wp]
    cmd += list(extra_args)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(workdir, "refcli.log"), "w") as log:
        proc = subprocess.Popen(cmd, cwd=root, stdout=log, stderr=log, stdin=subprocess.DEVNULL, start_new_session=True)
        try:
            proc.wait(timeout=timeout)
        finally:
            try:
                os.killpg(proc.pid, signal.SIGKILL)      # exactly the session this call started
            except ProcessLookupError:
                pass
    res = os.path.join(workdir, "refcli.json")
    if not os.path.exists(res):
        with open(os.path.join(workdir, "refcli.log")) as f:
            raise RuntimeError("reference CLI did not finish its hot path: " + f.read()[-1500:])
    with open(res) as f:
        return json.load(f)


if __name__ == "__main__":
    main(sys.argv[1:])
