"""Builds opticalflowscivis_b200/libofsv.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libofsv.so")
SOURCES = ["api.cu", "warp.cu", "warp_bwd.cu", "warp3d_slab.cu", "upflow_ops.cu", "upflow_bwd.cu", "ifnet_glue.cu", "conv_simt.cu", "conv_tc.cu", "conv_stack.cu", "conv_bwd.cu", "conv_halo_ring.cu", "block_stage.cu", "block_stage_hfast.cu", "metrics.cu", "adamw.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
]


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ofsv.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link libofsv.so next to this file.  No-op when up to date.
    Safe under torchrun: ranks serialise on a lock file, objects and the library are written under temporary names and moved into
    place, and staleness is re-checked once the lock is held (the first rank builds, the others find it done)."""
    if not force and not _stale():
        return SO
    import fcntl
    import tempfile
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    with open(os.path.join(objdir, ".lock"), "w") as lockf:
        fcntl.flock(lockf, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return SO
            tmpdir = tempfile.mkdtemp(prefix="obj.", dir=objdir)
            try:
                objs, procs = [], []
                for s in SOURCES:
                    o = os.path.join(tmpdir, s.replace(".cu", ".o"))
                    cmd = [nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
                    if verbose:
                        cmd.insert(1, "-Xptxas=-v")
                    procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
                    objs.append(o)
                failed = False
                for s, p in procs:
                    out, _ = p.communicate()
                    if p.returncode != 0 or verbose:
                        sys.stderr.write(f"--- nvcc {s} ---\n{out}\n")
                    failed |= p.returncode != 0
                if failed:
                    raise RuntimeError("nvcc failed")
                tmp_so = os.path.join(tmpdir, "libofsv.so")
                subprocess.check_call([nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp_so, *objs,
                                       "-lcudart_static", "-ldl", "-lrt", "-lpthread"])
                os.replace(tmp_so, SO)
            finally:
                shutil.rmtree(tmpdir, ignore_errors=True)
        finally:
            fcntl.flock(lockf, fcntl.LOCK_UN)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
