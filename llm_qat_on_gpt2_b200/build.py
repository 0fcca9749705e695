"""Build libspq_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m llm_qat_on_gpt2_b200.build [--force]

The library is a plain C-ABI shared object (include/spq_b200.h); it does not link against
torch.  It is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libspq_b200.so")
STAMP = os.path.join(HERE, ".libspq_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    files = sources() + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]
    files.append(os.path.join(INCLUDE, "spq_b200.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != _digest()


def build(force=False, verbose=False, out=None, extra_flags=()):
    """out / extra_flags: a side build for same-box A/B runs (e.g. out="libspq_epi8.so", extra_flags=["-DSPQ_EPI_WARPS=8"],
    selected at run time with SPQ_LIB=<path>); the default build is the one every import loads."""
    side = out is not None
    if not side and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build", os.path.basename(out)[:-3]) if side else os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"== {os.path.basename(src)}\n{log}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libspq_b200.so")
    target = os.path.join(HERE, out) if side else LIB
    link = [nvcc, "-shared", "-o", target] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link of libspq_b200.so failed")
    if side:
        return target
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB


if __name__ == "__main__":
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=outs[0] if outs else None,
                 extra_flags=[a for a in sys.argv[1:] if a.startswith("-D")])
    print(path)
