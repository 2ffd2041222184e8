"""Development aid: build an A/B variant of libpcl_b200.so with extra nvcc flags.

  python tools/build_variant.py NAME -DPCL_EMD_THREADS=1024 ...
writes pointcloud_b200/_build/variants/NAME.so; run anything with PCL_LIB_OVERRIDE=<that path> to use it."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_b200 import build as B  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
vdir = os.path.join(B.HERE, "_build", "variants")
odir = os.path.join(vdir, name + "_obj")
os.makedirs(odir, exist_ok=True)
procs = []
for s in B.SOURCES:
    obj = os.path.join(odir, s.replace(".cu", ".o"))
    procs.append((obj, subprocess.Popen([B._nvcc(), "-ccbin", "/usr/bin/g++", *B.NVCC_FLAGS, *extra, "-c", os.path.join(B.CSRC, s), "-o", obj])))
objs = []
for obj, p in procs:
    if p.wait():
        raise SystemExit("nvcc failed")
    objs.append(obj)
out = os.path.join(vdir, name + ".so")
subprocess.run([B._nvcc(), "-ccbin", "/usr/bin/g++", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs], check=True)
print(out)
