"""oracle/ref_shim.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Imports the UNMODIFIED reference (`/root/reference/PhenotypeSeeker/modeling.py`)
in this container so that golden vectors can be generated from the real code
(SURVEY.md Appendix C). On the GPU box, where /root/reference does not exist,
the copy that oracle/build.py installed under oracle/_ref/python is imported
instead (bench.py's CPU reference arm and the drop-in test use it there).

Four imports of modeling.py are not installed here (matplotlib, Bio, ete3,
statsmodels) and its `pkg_resources.require` pins fail against this image's
numpy/scipy/pandas/sklearn; they are stubbed before import.
`statsmodels.stats.weightstats.ttest_ind` is replaced by the restatement in
oracle/stats.py (the one piece of the reference's arithmetic that cannot be run
here — "restated", see that module's header).
"""
import os
import sys
import traceback
import types
import warnings

from . import build as _build

REF_ROOT = _build.ref_python_root() or "/root/reference"    # the source tree here, the installed copy on the GPU box


def available():
    return os.path.exists(os.path.join(REF_ROOT, "PhenotypeSeeker", "modeling.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_modeling():
    """-> the reference's `PhenotypeSeeker.modeling` module (real code)."""
    if "PhenotypeSeeker.modeling" in sys.modules:
        return sys.modules["PhenotypeSeeker.modeling"]
    if not available():
        raise RuntimeError("reference not present")
    from . import stats as _stats

    warnings.filterwarnings("ignore")
    mpl = _stub("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = _stub("matplotlib.pyplot")
    bio = _stub("Bio")
    bio.Phylo = _stub("Bio.Phylo")
    bio.Phylo.TreeConstruction = _stub("Bio.Phylo.TreeConstruction",
                                       DistanceTreeConstructor=object, _DistanceMatrix=object)
    _stub("ete3", Tree=object)
    sm = _stub("statsmodels")
    sm.stats = _stub("statsmodels.stats")

    def ttest_ind(x, y, usevar="unequal", weights=(None, None)):
        assert usevar == "unequal"
        return _stats.ttest_ind_weighted(x, y, weights[0], weights[1])

    sm.stats.weightstats = _stub("statsmodels.stats.weightstats", ttest_ind=ttest_ind)
    try:
        import pkg_resources
    except Exception:  # pragma: no cover
        pkg_resources = _stub("pkg_resources")
    pkg_resources.require = lambda *a, **k: None
    os.environ["PATH"] = (_build.ref_bin_dir() or os.path.join(REF_ROOT, "bin")) + os.pathsep + os.environ.get("PATH", "")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import PhenotypeSeeker.modeling as modeling  # noqa
    return modeling


class FakeSample:
    """The three attributes conduct_* reads from a Samples object."""

    def __init__(self, name, pheno, weight):
        self.name = name
        self.phenotypes = pheno
        self.weight = weight


def run_cli(argv, cwd):
    """Run the real CLI (`scripts/phenotypeseeker`) with argv in cwd.

    The sklearn section after get_ML_df crashes on this image's sklearn
    (modeling.py:1316); everything the hot path writes exists by then.
    """
    import runpy
    load_modeling()
    old_argv, old_cwd = sys.argv, os.getcwd()
    os.chdir(cwd)
    sys.argv = ["phenotypeseeker"] + list(argv)
    try:
        runpy.run_path(os.path.join(REF_ROOT, "scripts", "phenotypeseeker"), run_name="__main__")
    except SystemExit:
        pass
    except Exception as e:  # sklearn-era crash after the hot path
        sys.stderr.write(f"[ref_shim] reference stopped after hot path: {type(e).__name__}: {e}\n")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
