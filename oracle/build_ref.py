"""TEST / BENCH INFRASTRUCTURE.  Stages the UNMODIFIED reference sources the hot path needs into the git-ignored
`oracle/_ref/` so that the real reference can run where `/root/reference` does not exist (the GPU box):
`oracle/_ref/` is listed in .gitignore (never enters history) but not in .gpurunignore, so it travels with the
snapshot exactly like the built .so files.

    python oracle/build_ref.py            # in the build container (needs $S2AG_REFERENCE or /root/reference)

What is staged: `processor_v2.py`, `net/`, `utils/`, `torchlight/torchlight/` -- byte-for-byte copies; third-party
modules that are absent from the image (librosa, lmdb, ...) are stubbed at import time by oracle/ref_loader.py,
nothing is patched on disk.  Used by `bench.py --impl reference` (the reference's own CPU arm),
`--impl reference-gpu` (context arm: stock PyTorch/cuDNN on the same B200) and oracle/gen_golden.py.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
WHAT = ["processor_v2.py", "net", "utils", os.path.join("torchlight", "torchlight")]


def build(src=None, quiet=False):
    src = src or os.environ.get("S2AG_REFERENCE", "/root/reference")
    if not os.path.isdir(src):
        if os.path.isdir(DST):
            return DST  # already staged (GPU box): use as is
        raise FileNotFoundError("reference checkout not found at %s and %s is not staged" % (src, DST))
    os.makedirs(DST, exist_ok=True)
    for rel in WHAT:
        s, d = os.path.join(src, rel), os.path.join(DST, rel)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(s, d)
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(src + "\n")
    if not quiet:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print("staged %d reference files into %s" % (n, DST))
    return DST


if __name__ == "__main__":
    build(sys.argv[1] if len(sys.argv) > 1 else None)
