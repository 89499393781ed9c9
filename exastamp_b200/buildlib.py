"""Build libxsb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libxsb200.so")
SOURCES = ["xsb_core.cu", "xsb_nbr.cu", "xsb_pair.cu", "xsb_eam.cu", "xsb_ghost.cu", "xsb_assign.cu", "xsb_snap.cu", "xsb_thermo.cu", "xsb_ncclwin.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def nccl_device_include():
    """include directory of the NCCL >= 2.28 headers with the device API (nccl_device.h ships with the nvidia-nccl wheel that
    torch loads); None when absent: the peer-memory ghost transport is then compiled out and ncclSend/ncclRecv is used"""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations or []):
            inc = os.path.join(d, "include")
            if os.path.exists(os.path.join(inc, "nccl_device.h")):
                return inc
    except Exception:
        pass
    return None


def nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "xsb200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """compile + link when a source is newer than the library.  Several processes may call this at once (one rank per GPU
    under torchrun): an exclusive file lock serialises them, the link goes to a temporary name and is renamed into place, so
    no process ever dlopens a half-written file."""
    if not force and not stale():
        return LIB
    import fcntl
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            if force or stale():             # somebody else may have built it while we waited
                _build_locked(verbose)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)
    return LIB


def _build_locked(verbose):
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(o)
        extra = []
        if s == "xsb_ncclwin.cu" and nccl_device_include():
            extra = ["-DXSB_HAVE_NCCL_DEVICE=1", "-I" + nccl_device_include()]
        cmd = [nvcc()] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (s, out))
        if p.returncode != 0:
            failed = True
    with open(os.path.join(HERE, "build", "nvcc.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed, see exastamp_b200/build/nvcc.log")
    tmp = LIB + ".%d.tmp" % os.getpid()
    link = [nvcc(), "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    subprocess.check_call(link)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
