"""Builds a variant of libb200mpm.so with extra nvcc flags (e.g. -DG2P_PREFETCH=0) next to the real one, for A/B
timing in ONE gpurun call:  python tools/build_variant.py nopf -DG2P_PREFETCH=0
then on the GPU box:        B200MPM_LIB=wgsparkl_b200/_variants/lib_nopf.so python tools/run_config.py cube1m"""
import concurrent.futures
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wgsparkl_b200 import build as B  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
csrc = B.CSRC
if flags and flags[0].startswith("--csrc="):  # sources from another tree (e.g. a `git archive` of an older commit)
    csrc, flags = flags[0][7:], flags[1:]
out_dir = os.path.join(B.HERE, "_variants")
obj_dir = os.path.join(out_dir, "obj_" + name)
os.makedirs(obj_dir, exist_ok=True)


def cc(src):
    obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
    res = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + flags + ["-I", os.path.join(ROOT, "include"), "-c", os.path.join(csrc, src), "-o", obj], capture_output=True, text=True)
    if res.returncode != 0:
        raise SystemExit(res.stderr)
    open(obj + ".ptxas.log", "w").write(res.stderr)
    return obj


with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
    objs = list(ex.map(cc, B.SOURCES))
lib = os.path.join(out_dir, "lib_%s.so" % name)
subprocess.check_call([B._nvcc(), "-shared", "-o", lib] + objs + ["-lcudart", "-ldl"])
print(lib)
