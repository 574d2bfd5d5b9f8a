"""Host check of the word-wide (SWAR) FASTA emit used by k_decode_write_fasta: the per-chunk bit strings and
their placement (phenotypeseeker_b200/csrc/ps_decode_bits.h, free of CUDA-only constructs) against a
byte-at-a-time transducer on random FASTA-like text — headers mid-text, lower case, IUPAC, CR/TAB, bytes >= 128,
ignored prefixes, every output alignment. The GPU tests then check the whole kernel against glistmaker goldens."""
import os
import subprocess

from conftest import ROOT


def test_swar_fasta_emit_equals_byte_transducer(tmp_path):
    exe = str(tmp_path / "decode_bits_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "decode_bits_check.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
