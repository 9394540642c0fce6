"""Build an experimental variant of the library next to the product one: only encoder_tc.cu is recompiled with the given
-D flags, the other objects are taken from the product build.   python tools/build_variant.py NAME -DFOO=1 ...
-> point-cloud-audio_b200/csrc/libvar_NAME.so   (select it with PCAUDIO_B200_LIB=<path>)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

name, defs = sys.argv[1], sys.argv[2:]
g.build()
objdir = os.path.join(g.CSRC, "build")
obj = os.path.join(objdir, f"encoder_tc_{name}.o")
subprocess.run(["/usr/local/cuda/bin/nvcc"] + g.NVCC_FLAGS + defs + ["-c", "-o", obj, os.path.join(g.CSRC, "encoder_tc.cu")], check=True, cwd=g.CSRC)
objs = [os.path.join(objdir, s[:-3] + ".o") for s in g.SOURCES if s != "encoder_tc.cu"] + [obj]
out = os.path.join(g.CSRC, f"libvar_{name}.so")
subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", out] + objs, check=True, cwd=g.CSRC)
print(out)
