"""MTX writer benchmark: scipy.io.mmwrite -- what the reference calls, model.py:1418 --
against the device formatter on count matrices of configs[4] shape (28,001 loci x 100k cells).
  python tests/tools/mtx_bench.py [entries_in_millions=20] [real=0]"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.io
import scipy.sparse as sp
import torch

from stellarscope_b200 import mtx
from stellarscope_b200.engine import Engine

n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 20_000_000
real = len(sys.argv) > 2 and sys.argv[2] == '1'
K, Cn = 28001, 100000
rng = np.random.default_rng(0)
r, c = rng.integers(0, K, n), rng.integers(0, Cn, n)
v = (rng.integers(1, 9, n) / rng.integers(1, 5, n)) if real else rng.integers(1, 60, n).astype(np.int32)
m = sp.coo_matrix((v, (r, c)), shape=(K, Cn)).tocsc()
m.sum_duplicates(); m.sort_indices()
m = m.astype(np.float64 if real else np.int32)
d = tempfile.mkdtemp(prefix='ssmtx_')
a, b = os.path.join(d, 'a.mtx'), os.path.join(d, 'b.mtx')
eng = Engine(0)
mtx.mmwrite(b, m[:, :100].tocsc(), comment='warm', engine=eng)          # pinned ring, context
# every writer writes its file twice and the faster pass counts (the first pass into a new file also pays the
# file system's block allocation)
t_gpu = t_res = t_ref = 1e9
for rep in range(2):
    t0 = time.perf_counter(); mtx.mmwrite(b, m, comment='x', engine=eng); t_gpu = min(t_gpu, time.perf_counter() - t0)
coo = m.tocoo()
rd, cd, vd = (torch.from_numpy(x).cuda() for x in (coo.row, coo.col, coo.data.astype(np.float64)))
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); txt = eng.mtx_text(rd, cd, vd, not real); ev1.record(); torch.cuda.synchronize()
t_dev = ev0.elapsed_time(ev1) * 1e-3
for rep in range(2):
    t0 = time.perf_counter(); mtx.mmwrite(b, m, comment='x', engine=eng, device_coo=(rd, cd, vd)); t_res = min(t_res, time.perf_counter() - t0)
for rep in range(2):
    t0 = time.perf_counter(); scipy.io.mmwrite(a, m, comment='x'); t_ref = min(t_ref, time.perf_counter() - t0)
t0 = time.perf_counter()
with open(b, 'rb') as fh:
    blob = fh.read()
with open(b + '.copy', 'wb', buffering=0) as fh:
    fh.write(blob)
with open(b + '.copy', 'wb', buffering=0) as fh:
    t0 = time.perf_counter(); fh.write(blob); t_file = time.perf_counter() - t0
size = os.path.getsize(b)
same = open(a, 'rb').read() == open(b, 'rb').read() if not real else (scipy.io.mmread(b).tocsc() != m).nnz == 0
print(f'{"real" if real else "integer"} matrix {K} x {Cn}, {m.nnz} entries, file {size / 1e6:.0f} MB')
print(f'  scipy.io.mmwrite            {t_ref:8.3f} s  {m.nnz / t_ref / 1e6:8.1f} M entries/s')
print(f'  mtx.mmwrite (H2D+format+D2H+file) {t_gpu:8.3f} s  {m.nnz / t_gpu / 1e6:8.1f} M entries/s  x{t_ref / t_gpu:.1f}')
print(f'  mtx.mmwrite, triples resident (format+D2H+file; the output_report route) {t_res:8.3f} s  '
      f'{m.nnz / t_res / 1e6:8.1f} M entries/s  x{t_ref / t_res:.1f}')
print(f'  (writing the finished {size / 1e6:.0f} MB from host memory into a file alone: {t_file:.3f} s)')
print(f'  device part (measure + scan + write, triples resident) {t_dev * 1e3:8.2f} ms  '
      f'{(64 * m.nnz + txt.numel()) / t_dev / 1e9:8.1f} GB/s algorithmic (64 B/entry over the three passes + the text written once)')
print('  identical' if same else '  MISMATCH')
sys.exit(0 if same else 1)
