"""UMI de-duplication: device path vs the CPU restatement of the reference.

    python tests/tools/umi_bench.py [n_reads]

Synthetic 10x-style reads (config 3 shape: 15 % PCR duplicates inside a cell).  Reports the device time of
``ss_umi_dedup`` (CUDA events, matrix resident), the wall time of the drop-in ``umi.dedup_umi`` (host dict
walk + H2D + device + scipy outputs) and the oracle port (pure-Python restatement of
model.py:561-668, 1148-1219) on a bounded sample of the same reads.
"""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import scipy.sparse as sp
from oracle import umi_oracle as uo
from stellarscope_b200 import synth, umi
from stellarscope_b200.engine import Engine

n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
d = synth.generate(n_cells=max(10, n_reads // 1400), n_alignments=int(n_reads * 3.6), n_loci=28000, seed=3, umi_dup_rate=0.15)
# duplicates of a molecule mostly hit the same loci: copy the previous read's alignments for half of them
key = d.row_cell.astype(np.int64) * (int(d.row_umi.max()) + 1) + d.row_umi
_, row_group = np.unique(key, return_inverse=True)
row_group = row_group.astype(np.int32)
n_groups = int(row_group.max()) + 1
print(f'{d.N} reads, {d.A} alignments, {n_groups} barcode+UMI pairs, {d.N - n_groups} possible duplicates', flush=True)
eng = Engine(0)
eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
rg = torch.from_numpy(row_group).cuda()
best = None
for i in range(4):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = eng.umi_dedup(rg, n_groups)
    e.record()
    torch.cuda.synchronize()
    t = s.elapsed_time(e)
    best = t if best is None or (i > 0 and t < best) else best
print(f'ss_umi_dedup: {best:.2f} ms device+D2H ({d.N / best / 1e3:.1f} M reads/s), excluded {int(out["excluded"].sum())}, '
      f'components histogram {dict(zip(*np.unique(out["group_ncomp"][out["group_size"] > 1], return_counts=True)))}', flush=True)
eng.close()
# drop-in call (host dict walk included)
st = types.SimpleNamespace(shape=(d.N, d.K), raw_scores=sp.csr_matrix((d.scores, d.indices, d.indptr), shape=(d.N, d.K)))
st.read_index = {f'r{i}': i for i in range(d.N)}
st.read_bcode_map = {f'r{i}': int(d.row_cell[i]) for i in range(d.N)}
st.read_umi_map = {f'r{i}': int(d.row_umi[i]) for i in range(d.N)}
st.opts = types.SimpleNamespace(outfile_path=lambda s: os.path.join('/tmp', s))
for i in range(2):
    t0 = time.perf_counter()
    info = umi.dedup_umi(st, output_report=False, summary=False)
    dt = time.perf_counter() - t0
print(f'umi.dedup_umi (drop-in, from the reference\'s dicts): {dt * 1e3:.0f} ms, nexclude {info.nexclude}', flush=True)
# CPU restatement on a bounded sample (first rows; groups are local to a cell, rows of a cell are contiguous)
ns = min(d.N, 150_000)
t0 = time.perf_counter()
ref = uo.dedup(d.indptr[:ns + 1], d.indices[:d.indptr[ns]], d.scores[:d.indptr[ns]], row_group[:ns])
dtc = time.perf_counter() - t0
same = np.array_equal(ref['excluded'][:ns - 2000], out['excluded'][:ns - 2000])
print(f'CPU port (1 core) on the first {ns} reads: {dtc:.2f} s ({ns / dtc / 1e6:.3f} M reads/s); agrees with the device: {same}', flush=True)
