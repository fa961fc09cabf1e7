"""Builds reface_b200/librefaceb200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m reface_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librefaceb200.so")
OBJ = os.path.join(HERE, "_build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xptxas", "-v", "--expt-relaxed-constexpr"]
SOURCES = ["engine.cu", "unet.cu", "vae.cu", "clip.cu", "arcface.cu", "parse.cu", "paste.cu", "attn_flash.cu", "capi.cu"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "reface_b200.h")]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest(deps):
        return OUT
    os.makedirs(OBJ, exist_ok=True)

    def cc(src):
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        return src, obj, r

    objs, logs = [], []
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        for src, obj, r in ex.map(cc, srcs):
            logs.append((src, r.stderr))
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed on {src}")
            objs.append(obj)
    with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
        for src, log in logs:
            f.write(f"==== {src}\n{log}\n")
    r = subprocess.run([NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    if verbose:
        for src, log in logs:
            print(f"==== {src}\n{log}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
