"""Build a tuning variant of the library:  python scripts/build_variant.py NAME -DSS_E_STREAM=1408 ...
-> stellarscope_b200/variants/libss_NAME.so (git-ignored; select it with STELLARSCOPE_B200_LIB=<path>)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stellarscope_b200 import build as b

name, defs = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(b.HERE, 'variants')
os.makedirs(os.path.join(out_dir, name), exist_ok=True)
objs, procs = [], []
for s in b.SOURCES:
    o = os.path.join(out_dir, name, s.replace('.cu', '.o'))
    objs.append(o)
    procs.append(subprocess.Popen([b._nvcc()] + b.NVCC_FLAGS + defs + ['-c', os.path.join(b.CSRC, s), '-o', o]))
assert all(p.wait() == 0 for p in procs)
lib = os.path.join(out_dir, f'libss_{name}.so')
subprocess.check_call([b._nvcc(), '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
for o in objs:
    os.remove(o)
print(lib)
