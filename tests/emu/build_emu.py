"""TEST INFRASTRUCTURE ONLY.  Compiles the SIMT kernel sources with g++ against tests/emu/cuda_emu.h
(a fiber-based CUDA execution-model emulator) into tests/emu/libs2ag_emu.so so kernel LOGIC can be
checked on the CPU-only dev container.  The product path never loads this library."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
CSRC = os.path.join(ROOT, "speech2affective_gestures_b200", "csrc")
LIB = os.path.join(HERE, "libs2ag_emu.so")
# kernels that use sm_100a-only PTX (tcgen05/TMA) cannot be emulated; they are checked on the GPU
# against the SIMT kernels instead.
SKIP_PREFIX = ("umma_",)


def build(force=False):
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "s2ag.h")]
    jobs, objs = [], []
    for s in sorted(os.listdir(CSRC)):
        if not s.endswith(".cu") or s.startswith(SKIP_PREFIX):
            continue
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in [src] + headers):
            jobs.append((src, obj))

    def run(job):
        src, obj = job
        cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-fPIC", "-DS2AG_EMU", "-I", HERE, "-I", os.path.join(ROOT, "include"), "-Wno-unknown-pragmas",
               "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed for %s:\n%s" % (src, r.stderr[-6000:]))

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    if force or jobs or not os.path.exists(LIB):
        r = subprocess.run(["g++", "-shared", "-o", LIB] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
