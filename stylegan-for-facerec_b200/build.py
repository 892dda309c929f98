"""Build libsg2_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m build  (from the package dir)   or   python __graft_entry__.py build
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libsg2_b200.so")
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defs=(), tag=""):
    """defs / tag: a kernel-variant build (extra -D flags) into libsg2_b200_<tag>.so for same-box A/B runs
    through SG2_B200_LIB; the default build takes neither."""
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "sg2_b200.h")]
    objdir = os.path.join(CSRC, "build" + ("_" + tag if tag else ""))
    out = OUT if not tag else OUT[:-3] + "_" + tag + ".so"
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + list(defs) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout[-4000:]}\n{r.stderr[-4000:]}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _stale(out, objs):
        cmd = [nvcc, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout[-4000:]}\n{r.stderr[-4000:]}")
    return out


if __name__ == "__main__":
    # python build.py [--force] [--tag NAME -DX=1 ...]
    argv = sys.argv[1:]
    tag = argv[argv.index("--tag") + 1] if "--tag" in argv else ""
    print(build(force="--force" in argv, verbose=True, defs=[a for a in argv if a.startswith("-D")], tag=tag))
