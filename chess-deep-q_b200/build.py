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
SOURCES = ["abi.cu", "eval_kernel.cu", "playout_kernel.cu", "tree_kernel.cu", "conv_stream_kernel.cu", "pipe_kernel.cu", "fc_kernel.cu", "fc_split_kernel.cu", "fc_bwd_kernel.cu", "conv_bwd2_kernel.cu", "train_kernels.cu", "dp_adam_kernel.cu"]
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
    """Every .cu is compiled to its own object (in parallel, rebuilt only when it or a header changed) and the objects
    are linked into the shared library.  defines/out: build an A/B variant, e.g.
    build(defines=["CDQ_ALL_LANES_POLL"], out="libcdq_alt.so") -- variants use their own object directory."""
    from concurrent.futures import ThreadPoolExecutor
    if out is None and not force and not stale():
        return OUT
    target = os.path.join(HERE, out) if out else OUT
    objdir = os.path.join(HERE, "build", "obj" if not out else "obj_" + os.path.splitext(out)[0])
    os.makedirs(objdir, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)  # this image exports a CC wrapper that nvcc's host compile does not need
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [os.path.join(HERE, "..", "include", "cdq.h"), __file__]
    newest_header = max(os.path.getmtime(h) for h in headers)
    compile_flags = [f for f in NVCC_FLAGS if f != "--shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_header):
            return obj, 0, ""
        cmd = [nvcc()] + compile_flags + ["-D" + d for d in defines] + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        return obj, res.returncode, " ".join(cmd) + "\n" + res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, sources()))
    log = "".join(r[2] for r in results)
    failed = any(r[1] != 0 for r in results)
    if not failed:
        cmd = [nvcc(), "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", target] + [r[0] for r in results]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        failed = res.returncode != 0
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(log)
    if failed:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libcdq.so")
    if verbose:
        print(log)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
