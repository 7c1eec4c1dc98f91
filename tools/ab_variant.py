"""Build a kernel variant of the current tree into factory_marl_b200/csrc/build/variants/<name>.so for A/B timing on the GPU
box (tools/ab_run.sh; the product always loads the in-tree libfm_b200.so).
usage: python tools/ab_variant.py <name> [extra nvcc flags ...]"""
import os, subprocess, sys
sys.path.insert(0, ".")
from factory_marl_b200 import lib
name, flags = sys.argv[1], sys.argv[2:]
d = os.path.join(os.path.dirname(lib.SO_PATH), "csrc", "build", "variants", name)
only = os.environ.get("FM_AB_ONLY")          # e.g. FM_AB_ONLY=fm_step_shape_2_14.cu: rebuild just the benchmarked block shape
out = lib.build(force=True, extra_flags=flags, out=d + ".so", objdir=d, only=only.split(",") if only else None)
for o in sorted(os.listdir(d)):
    if o.startswith("fm_step_shape"):
        r = subprocess.run(["cuobjdump", "-res-usage", os.path.join(d, o)], capture_output=True, text=True).stdout
        print(o, [t for t in r.split() if t.startswith(("REG:", "STACK:"))][:2])
print(out)
