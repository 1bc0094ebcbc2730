"""In-tree build of libcrm_b200.so (sm_100a only) with nvcc.  `python -m cellregmap_b200.build [--force] [-v]`.

Every csrc/*.cu is one translation unit, compiled in parallel into build/ and linked into
cellregmap_b200/libcrm_b200.so (git-ignored; it travels to the GPU box with the tree)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(ROOT, "build", "crm_b200")
LIB = os.path.join(HERE, "libcrm_b200.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _units():
    return [f for f in sorted(os.listdir(CSRC)) if f.endswith(".cu") or f.endswith(".cpp")]


def _deps():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "crm_b200.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _deps())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.path.join(CUDA_HOME, "bin", "nvcc")
    os.makedirs(OBJDIR, exist_ok=True)
    newest_header = max(os.path.getmtime(s) for s in _deps() if not (s.endswith(".cu") or s.endswith(".cpp")))

    def compile_one(unit):
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJDIR, os.path.splitext(unit)[0] + ".o")
        stamp = max(os.path.getmtime(src), newest_header)
        if unit == "kernels_fit_null.cu":
            stamp = max(stamp, os.path.getmtime(os.path.join(CSRC, "kernels_fit_g.cu")))
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > stamp:
            return obj
        if unit.endswith(".cpp"):      # host-only translation units (worker pool, narrowing loops): plain g++
            cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-pthread", "-I" + os.path.join(CUDA_HOME, "include"), "-c", src, "-o", obj]
        else:
            cmd = [nvcc] + ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
                                   "-c", src, "-o", obj]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, _units()))
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcusolver", "-lcublas", "-lcublasLt", "-Xcompiler", "-pthread", "-Xlinker", "-rpath," + os.path.join(CUDA_HOME, "lib64")]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
