"""Score-matrix construction: device path vs the reference's dok_matrix loop.

    python tests/tools/scores_bench.py [n_records]

Synthetic loader output (one record per fragment x locus mapping; 25 % of the (read, locus) pairs repeated).
Reports the device time of ``ss_scores_build`` (CUDA events, records resident), the wall time of the drop-in
``loader.mapping_to_matrix`` (C walk of the tuple list + H2D + device + D2H + scipy object) and the CPU
restatements on a bounded sample: the reference's literal loop (dict + max, model.py:966-969) and the
vectorised numpy port.
"""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import scores_oracle as so
from stellarscope_b200 import loader
from stellarscope_b200.engine import Engine

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
rng = np.random.default_rng(1)
n_reads, n_feats = max(1, n // 4), 28001
read = np.unique(rng.integers(0, n_reads, n), return_inverse=True)[1]  # every read has a record
read = np.sort(read).astype(np.int32)                                  # name-collated BAM: reads arrive grouped
n_reads = int(read[-1]) + 1
feat = (1 + (read.astype(np.int64) * 7919 + rng.integers(0, 12, n)) % (n_feats - 1)).astype(np.int32)
alnscore = rng.integers(60, 200, n).astype(np.int32)
alnlen = rng.integers(30, 152, n).astype(np.int32)
print(f'{n} records, {n_reads} reads, {n_feats} features', flush=True)
eng = Engine(0)
lib, h = eng.lib, eng.h
d = [torch.from_numpy(a).cuda() for a in (read, feat, alnscore, alnlen)]
row_ptr = torch.empty(n_reads + 1, dtype=torch.int32, device='cuda')
col = torch.empty(n, dtype=torch.int32, device='cuda')
sc = torch.empty(n, dtype=torch.uint16, device='cuda')
info = np.zeros(4, dtype=np.int64)
from stellarscope_b200 import _lib
best = None
for i in range(4):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rc = lib.ss_scores_build(h, eng._stream(), n, *(eng._p(t) for t in d), 60, n_reads, n_feats, eng._p(row_ptr), eng._p(col),
                             eng._p(sc), info.ctypes.data_as(_lib.c_i64p))
    e.record()
    torch.cuda.synchronize()
    assert rc == 0
    t = s.elapsed_time(e)
    best = t if best is None or (i > 0 and t < best) else best
nnz = int(info[0])
print(f'ss_scores_build: {best:.2f} ms ({n / best / 1e3:.0f} M records/s; {n * 16 / best / 1e6:.0f} GB/s of record bytes), '
      f'nnz {nnz}, reads without feature {int(info[3])}', flush=True)
# drop-in call from the loader's tuples (bounded: building 5-tuples of Python objects is the slow part here)
m = min(n, 2_000_000)
qn = [f'r{i}' for i in range(int(read[m - 1]) + 1)]
fn = ['__no_feature'] + [f'L{j}' for j in range(1, n_feats)]
maps = [('PM', qn[r], fn[f], s_, l_) for r, f, s_, l_ in zip(read[:m].tolist(), feat[:m].tolist(), alnscore[:m].tolist(), alnlen[:m].tolist())]
st = types.SimpleNamespace(read_index={q: i for i, q in enumerate(qn)}, feat_index={f: i for i, f in enumerate(fn)},
                           opts=types.SimpleNamespace(no_feature_key='__no_feature'))
for i in range(3):
    t0 = time.perf_counter()
    loader.mapping_to_matrix(st, maps, 60)
    dt = time.perf_counter() - t0
print(f'loader.mapping_to_matrix (drop-in, from {m} tuples + dicts): {dt * 1e3:.1f} ms ({m / dt / 1e6:.1f} M records/s)', flush=True)
t0 = time.perf_counter()
ref, _ = so.mapping_to_matrix(read[:m], feat[:m], alnscore[:m], alnlen[:m], 60, len(qn), n_feats)
dtv = time.perf_counter() - t0
same = (np.array_equal(st.raw_scores.indptr, ref.indptr) and np.array_equal(st.raw_scores.indices, ref.indices)
        and np.array_equal(st.raw_scores.data, ref.data))
ml = min(m, 300_000)
t0 = time.perf_counter()
so.mapping_to_matrix_loop(read[:ml], feat[:ml], alnscore[:ml], alnlen[:ml], 60, len(qn), n_feats)
dtl = time.perf_counter() - t0
print(f'CPU: numpy port {m / dtv / 1e6:.2f} M records/s ({m} records), literal dict loop {ml / dtl / 1e6:.2f} M records/s '
      f'({ml} records; the reference\'s dok_matrix is slower still); device result equals the port: {same}', flush=True)
eng.close()
