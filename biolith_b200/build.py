"""In-tree build of libbiolith_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m biolith_b200.build            # incremental
    python -m biolith_b200.build --force

The shared object lands next to this file (git-ignored, but it travels to the GPU box).
"""

from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libbiolith_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden", "--expt-relaxed-constexpr", *os.environ.get("BL_EXTRA_NVCC_FLAGS", "").split(),
]
SOURCES = ["api.cu", "pack.cu", "occu.cu", "occu_chain.cu", "occu_signed.cu", "occu_small.cu", "occu_rn.cu", "occu_rn2.cu", "occu_re.cu", "obs_loglik.cu", "multi.cu", "occu_cop.cu", "nmixture.cu", "occu_cs.cu", "comm.cu", "nuts.cu", "xla.cu", "microbench.cu"]


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _deps_hash(src: str) -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    files = [os.path.join(CSRC, src)] + sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))
    ) + [os.path.join(HERE, "..", "include", "biolith_b200.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _compile(src: str, force: bool) -> tuple[str, bool]:
    obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
    stamp = obj + ".sha"
    want = _deps_hash(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
        return obj, False
    cmd = [NVCC, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(want)
    return obj, True


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in results]
    if any(changed for _, changed in results) or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[biolith_b200.build] linked {LIB} ({os.path.getsize(LIB) / 1e6:.1f} MB)")
    elif verbose:
        print(f"[biolith_b200.build] up to date: {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
