"""A/B helper: builds a second copy of the library with extra -D flags (e.g. -DGDRN_PDL_EW_NO_TRIGGER=1) into
gdr_net_b200/lib/libgdrn_b200_<tag>.so; select it with GDRN_LIB_PATH.  Usage: python tools/build_variant.py <tag> <flag>..."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gdr_net_b200 import build as B
tag, extra = sys.argv[1], sys.argv[2:]
bdir = os.path.join(B.HERE, "_build_" + tag)
os.makedirs(bdir, exist_ok=True)
objs = []
def one(src):
    obj = os.path.join(bdir, src[:-3] + ".o")
    subprocess.check_call([B.NVCC] + B.FLAGS + extra + ["-I", B.CSRC, "-c", os.path.join(B.CSRC, src), "-o", obj])
    return obj
with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(one, B._sources()))
out = os.path.join(B.LIBDIR, f"libgdrn_b200_{tag}.so")
subprocess.check_call([B.NVCC, "-shared", "-o", out] + objs + ["-cudart", "shared", "-Xlinker", "-rpath,/usr/local/cuda/lib64"])
print(out)
