"""Build libcdq.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python chess-deep-q_b200/build.py [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcdq.so")
SOURCES = ["abi.cu", "eval_kernel.cu", "playout_kernel.cu", "conv_stream_kernel.cu", "pipe_kernel.cu", "fc_kernel.cu", "fc_split_kernel.cu", "fc_bwd_kernel.cu", "conv_bwd2_kernel.cu", "train_kernels.cu", "dp_adam_kernel.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--shared", "-Xptxas", "-v", "-diag-suppress", "177",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "cdq.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: build an A/B variant, e.g. build(defines=["CDQ_ALL_LANES_POLL"], out="libcdq_alt.so")."""
    if out is None and not force and not stale():
        return OUT
    target = os.path.join(HERE, out) if out else OUT
    cmd = [nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", target] + sources()
    env = dict(os.environ)
    env.pop("CC", None)  # this image exports a CC wrapper that nvcc's host compile does not need
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libcdq.so")
    if verbose:
        print(log)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
