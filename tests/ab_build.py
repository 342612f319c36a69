"""Builds a VARIANT of libofsv.so with extra nvcc flags for A/B timing experiments on the GPU box:
    python tests/ab_build.py <tag> [-DFLAG ...]   ->  opticalflowscivis_b200/libofsv_<tag>.so   (select it with OFSV_LIB=<path>)
Probe builds (-DOFSV_STACK_PROBE, -DOFSV_SLAB_PROBE) can skip work and produce wrong results: timing only, never shipped."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import build as b  # noqa: E402

tag, flags = sys.argv[1], sys.argv[2:]
objdir = os.path.join(b.HERE, "build", "ab_" + tag)
os.makedirs(objdir, exist_ok=True)
procs, objs = [], []
for s in b.SOURCES:
    o = os.path.join(objdir, s.replace(".cu", ".o"))
    procs.append(subprocess.Popen([b.nvcc(), *b.NVCC_FLAGS, *flags, "-c", os.path.join(b.CSRC, s), "-o", o]))
    objs.append(o)
assert all(p.wait() == 0 for p in procs)
out = os.path.join(b.HERE, f"libofsv_{tag}.so")
subprocess.check_call([b.nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"])
print(out)
