"""oracle/build.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Builds the checker:
  * oracle/_build/libkmer_oracle.so  from oracle/kmer_oracle.c (gcc -O2)
  * oracle/_ref/bin/{glistmaker,glistcompare,glistquery,gmer_counter}
    when /root/reference is present (this container only).
  * oracle/_ref/python/{PhenotypeSeeker/*.py,scripts/phenotypeseeker}: the reference's own Python,
    installed unmodified (same condition).

About oracle/_ref: the reference ships GenomeTester4 only as prebuilt x86-64
ELF binaries (no source anywhere under /root/reference — SURVEY.md §2), so
there is nothing to compile. The recipe therefore does what the reference's
own install.sh:11 does (`cp bin/* <venv>/bin`): it installs the unmodified
binaries into oracle/_ref/bin (git-ignored, travels to the GPU box). They are
used (a) to pin the C restatement and (b) as the CPU reference arm of bench.py.
No reference SOURCE is copied anywhere.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REF = os.path.join(HERE, "_ref")
LIB = os.path.join(BUILD, "libkmer_oracle.so")
REF_SRC_BIN = "/root/reference/bin"
REF_TOOLS = ("glistmaker", "glistcompare", "glistquery", "gmer_counter")


def build_oracle(force=False):
    src = os.path.join(HERE, "kmer_oracle.c")
    os.makedirs(BUILD, exist_ok=True)
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(src)):
        return LIB
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", LIB, src])
    return LIB


def install_ref_tools():
    """Install the shipped GenomeTester4 binaries into oracle/_ref/bin."""
    if not os.path.isdir(REF_SRC_BIN):
        return None
    dst = os.path.join(REF, "bin")
    os.makedirs(dst, exist_ok=True)
    for t in REF_TOOLS:
        s = os.path.join(REF_SRC_BIN, t)
        d = os.path.join(dst, t)
        if os.path.exists(s) and not (os.path.exists(d)
                                      and os.path.getsize(d) == os.path.getsize(s)):
            shutil.copyfile(s, d)
            os.chmod(d, 0o755)
    return dst


REF_SRC_ROOT = "/root/reference"


def install_ref_python():
    """Install the reference's Python (package + CLI script, unmodified) into oracle/_ref/python — what
    the reference's install.sh does with `pip install .`, into a private, git-ignored directory that
    travels to the GPU box next to the binaries. Lets the CPU reference arm of bench.py and the drop-in
    test run the REAL modeling.py there (through oracle/ref_shim.py's stubs for the four modules this
    image lacks). Nothing under oracle/_ref is product code or enters the history."""
    src_pkg = os.path.join(REF_SRC_ROOT, "PhenotypeSeeker")
    src_cli = os.path.join(REF_SRC_ROOT, "scripts", "phenotypeseeker")
    if not (os.path.isdir(src_pkg) and os.path.exists(src_cli)):
        return None
    dst = os.path.join(REF, "python")
    os.makedirs(os.path.join(dst, "PhenotypeSeeker"), exist_ok=True)
    os.makedirs(os.path.join(dst, "scripts"), exist_ok=True)
    for fn in os.listdir(src_pkg):
        if fn.endswith(".py"):
            shutil.copyfile(os.path.join(src_pkg, fn), os.path.join(dst, "PhenotypeSeeker", fn))
    shutil.copyfile(src_cli, os.path.join(dst, "scripts", "phenotypeseeker"))
    return dst


def ref_python_root():
    """Root holding PhenotypeSeeker/modeling.py and scripts/phenotypeseeker, or None."""
    for d in (REF_SRC_ROOT, os.path.join(REF, "python")):
        if os.path.exists(os.path.join(d, "PhenotypeSeeker", "modeling.py")):
            return d
    return None


def ref_bin_dir():
    """Directory holding the reference's native tools, or None."""
    for d in (os.path.join(REF, "bin"), REF_SRC_BIN):
        if os.path.exists(os.path.join(d, "glistmaker")):
            return d
    return None


def build_all():
    build_oracle()
    install_ref_tools()
    install_ref_python()


if __name__ == "__main__":
    build_all()
    print(LIB, ref_bin_dir())
