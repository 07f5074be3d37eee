"""Checkpoint benchmark (host only, no GPU needed): the reference's format -- pickle of the
whole object, Stellarscope.save / load, model.py:1259-1289 -- against the side-car, on an object of the shape
load_alignment leaves behind (N reads, 10k barcodes, UMIs), plus the time until the hot path has its row arrays.
  python tests/tools/checkpoint_bench.py [reads_in_millions=5]"""
import os
import pickle
import sys
import tempfile
import time
import types
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.sparse as sp

from stellarscope_b200 import checkpoint, model

N = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 5_000_000
C, K = 10000, 28001
rng = np.random.default_rng(0)
cell = rng.integers(0, C, N)
umi = rng.integers(0, 4 ** 6, N)
st = types.SimpleNamespace()
st.opts = types.SimpleNamespace(ignore_umi=False, umi_counts=True, pooling_mode='individual')
st.read_index, st.read_bcode_map, st.read_umi_map = {}, {}, {}
st.bcode_ridx_map, st.bcode_umi_map = defaultdict(set), defaultdict(list)
t0 = time.perf_counter()
for r, (c, u) in enumerate(zip(cell.tolist(), umi.tolist())):
    q = f'A00123:45:HXXXXXXX:1:{1101 + r % 500}:{r}:{r % 9973}'
    bc, um = f'BC{c:012d}-1', f'U{u:011d}'
    st.read_index[q] = r
    st.read_bcode_map[q] = bc
    st.read_umi_map[q] = um
    st.bcode_ridx_map[bc].add(r)
    st.bcode_umi_map[bc].append(um)
print(f'object with {N} reads built in {time.perf_counter() - t0:.1f} s')
lens = rng.integers(2, 6, N)
ptr = np.zeros(N + 1, dtype=np.int32); np.cumsum(lens, out=ptr[1:])
st.raw_scores = sp.csr_matrix((rng.integers(140, 260, ptr[-1]).astype(np.uint16), rng.integers(0, K, ptr[-1]).astype(np.int32), ptr),
                              shape=(N, K))
st.shape = (N, K)
st.feat_index = {f'L{j}': j for j in range(K)}
st.filtlist, st.celltypes, st.ctype_bcode_map, st.bcode_ctype_map = {}, [], defaultdict(set), {}
st.corrected = st.umi_dups = None
st.reassignments = {}
d = tempfile.mkdtemp(prefix='ss_ckpt_')
a, b = os.path.join(d, 'ref.pickle'), os.path.join(d, 'ours.pickle')


def hot_arrays(o):
    seg, keys = model.pooling_row_segments(o, o.opts)
    bc_list = sorted(checkpoint.row_table(o).bcodes) if checkpoint.row_table(o) is not None else sorted(o.bcode_ridx_map)
    rc, ru = model._row_outputs(o, {x: i for i, x in enumerate(bc_list)})
    return seg, rc, ru


t0 = time.perf_counter()
with open(a, 'wb') as fh:
    pickle.dump(st, fh)
t_ref_save = time.perf_counter() - t0
t0 = time.perf_counter(); checkpoint.save(st, b); t_save = time.perf_counter() - t0
t0 = time.perf_counter()
with open(a, 'rb') as fh:
    ref = pickle.load(fh)
t_ref_load = time.perf_counter() - t0
t0 = time.perf_counter(); ref_arr = hot_arrays(ref); t_ref_hot = time.perf_counter() - t0
t0 = time.perf_counter(); ours = checkpoint.load(b); t_load = time.perf_counter() - t0
t0 = time.perf_counter(); our_arr = hot_arrays(ours); t_hot = time.perf_counter() - t0
same = np.array_equal(ref_arr[0], our_arr[0]) and np.array_equal(ref_arr[1], our_arr[1])
t0 = time.perf_counter(); full = dict(ours.read_index) == ref.read_index and dict(ours.read_bcode_map) == ref.read_bcode_map; t_mat = time.perf_counter() - t0
sz = lambda p: os.path.getsize(p) / 1e6
print(f'reference (pickle):  save {t_ref_save:6.2f} s   load {t_ref_load:6.2f} s   row arrays of the hot path {t_ref_hot:6.2f} s   file {sz(a):.0f} MB')
print(f'side-car:            save {t_save:6.2f} s   load {t_load:6.2f} s   row arrays of the hot path {t_hot:6.2f} s   files {sz(b) + sz(checkpoint.sidecar_path(b)):.0f} MB')
print(f'resume until the fit can start: {t_ref_load + t_ref_hot:.2f} s -> {t_load + t_hot:.2f} s (x{(t_ref_load + t_ref_hot) / (t_load + t_hot):.1f}); '
      f'save x{t_ref_save / t_save:.1f}')
print(f'materialising read_index + read_bcode_map on demand and comparing with the unpickled ones: {t_mat:.2f} s, equal={full}; row arrays equal={same}')
sys.exit(0 if (same and full) else 1)
