"""Builds soundsym_b200/libsoundsym_b200.so (in-tree) with nvcc for sm_100a.

  python -m soundsym_b200.build [--force] [--verbose]

Every .cu under csrc/ is compiled to an object with
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17
and linked into one shared library with a static cudart (so the library does not depend on torch's runtime).
exact.cu is compiled with --fmad=false: its f64 kernels must round multiply and add separately, like the CPU path.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libsoundsym_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--cudart", "static"]
PER_FILE = {"exact.cu": ["--fmad=false"], "segment.cu": ["--fmad=false"]}


def _newer(a, deps):
    return (not os.path.exists(a)) or any(os.path.getmtime(d) > os.path.getmtime(a) for d in deps)


def build(force=False, verbose=False, variant=None, defs=()):
    """variant / defs: development A/B builds, e.g. variant="v7", defs=["-DSS_TC_SLOTS=4"] writes
    libsoundsym_b200.v7.so next to the product library (select it with SS_B200_LIB=<path>)."""
    OBJ = os.path.join(HERE, "csrc", "_obj" + ("_" + variant if variant else ""))
    LIB = os.path.join(HERE, "libsoundsym_b200%s.so" % ("." + variant if variant else ""))
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "soundsym_b200.h"))
    objs, procs = [], []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            cmd = [NVCC] + ARCH + COMMON + list(defs) + PER_FILE.get(s, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s ----\n%s\n" % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _newer(LIB, objs):
        subprocess.check_call([NVCC] + ARCH + ["--cudart", "static", "-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    _variant = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")), None)
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=_variant,
                defs=[a for a in sys.argv if a.startswith("-D")]))
