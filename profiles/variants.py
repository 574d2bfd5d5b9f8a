"""Build tuning variants of libpskmer.so with -D overrides into phenotypeseeker_b200/_variants/
(git-ignored; they travel to the GPU box). Select one with PSKMER_LIB=<path>.

  python profiles/variants.py name:-DPP_THREADS=256:-DPP_MIN_BLOCKS=4 ...
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phenotypeseeker_b200 import build as b

out = os.path.join(ROOT, "phenotypeseeker_b200", "_variants")
os.makedirs(out, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, *defs = spec.split(":")
    flags = [f for f in b.NVCC_FLAGS if not f.startswith("--use_fast_math")]
    cmd = [b._nvcc()] + flags + defs + ["-o", os.path.join(out, f"libpskmer_{name}.so"), os.path.join(b.CSRC, "ps_api.cu")]
    procs.append((name, subprocess.Popen(cmd)))
for name, p in procs:
    print(name, "rc", p.wait())
