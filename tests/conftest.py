import base64
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_json(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_kmer_lists():
    cases = load_json("kmer_lists.json")
    for c in cases:
        c["data"] = base64.b64decode(c["data_b64"])
        c["kmers"] = np.array(c["kmers"], dtype=np.uint64)
        c["counts"] = np.array(c["counts"], dtype=np.uint32)
    return cases


@pytest.fixture(scope="session")
def golden_stage3():
    return load_json("stage3.json")


@pytest.fixture(scope="session")
def ctx():
    """One libpskmer context on cuda:0 for the whole GPU session (no CPU fallback)."""
    from phenotypeseeker_b200._native import Context
    c = Context(0)
    yield c
    c.close()


def stage3_inputs(case):
    """golden stage-3 case -> (presence U x N, pheno array, weights or None)."""
    pres = np.array(case["presence"], dtype=np.uint8)
    if case["kind"] == "chi2":
        ph = np.array([-1 if p == "NA" else int(p) for p in case["pheno"]], dtype=np.int8)
    else:
        ph = np.array([np.nan if p == "NA" else float(p) for p in case["pheno"]], dtype=np.float64)
    w = np.array(case["weights"], dtype=np.float64)
    return pres, ph, w
