"""Fragment-major tile form (stellarscope_b200/layout_fm.py): format specification for the next streaming kernel,
checked on the CPU -- one fused E+M pass computed lane by lane from the FM records equals the pass computed from the
definition (E-step model.py:207-210, M-step :236), for tiles with all row / column group sizes, with rows and
columns that overflow the lane budget, and the records account for every alignment exactly once."""
import numpy as np
import pytest

import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import layout_fm
from stellarscope_b200 import layout_host


def _layout(seed, n_cells, n_alignments, max_k=32, n_loci=300):
    from stellarscope_b200 import synth
    d = synth.generate(n_cells=n_cells, n_alignments=n_alignments, n_loci=n_loci, seed=seed, max_k=max_k)
    return layout_host.build(d.indptr, d.indices, d.scores, d.row_cell, d.n_cells, d.K)


@pytest.mark.parametrize('threads,fr,fc', [(512, 2, 4), (256, 4, 6), (64, 1, 1), (32, 2, 2)])
def test_fm_pass_equals_definition(threads, fr, fc):
    L = _layout(seed=3, n_cells=6, n_alignments=14000)
    rng = np.random.default_rng(threads)
    seen_generic_rows = seen_generic_cols = 0
    for t in range(L.n_tiles):
        s = layout_fm.tile_slices(L, t)
        fm = layout_fm.build_fm(L, t, threads, fr, fc)
        # every alignment of a covered row sits in exactly one lane, with its own offsets
        n = fm.rows['meta'] & 0xff
        used = [(int(o) & 0xffff) >> 3 for rec, k in zip(fm.rows, n) for o in rec['off'][:k]]
        cpos = [int(o) >> 19 for rec, k in zip(fm.rows, n) for o in rec['off'][:k]]
        gen = sum(int(s['rlen'][r]) for r in fm.row_generic)
        assert len(cpos) + gen == s['nent'] and len(set(cpos)) == len(cpos)
        assert all(c < s['ncol'] for c in used)
        x = rng.random(s['ncol']) * 1e-4 + 1e-9
        T_fm = layout_fm.iterate_fm(fm, s, x)
        T_def = layout_fm.iterate_direct(s, x)
        assert np.allclose(T_fm, T_def, rtol=1e-12, atol=0), t
        seen_generic_rows += len(fm.row_generic)
        seen_generic_cols += len(fm.col_generic)
    if threads <= 64:
        assert seen_generic_rows > 0 and seen_generic_cols > 0     # the overflow paths were exercised


def test_fm_long_rows_and_hot_columns():
    """Rows of up to 32 alignments (groups of 8 lanes) and columns with hundreds of rows (groups of 32 lanes, and
    the > 128 class that stays generic)."""
    L = _layout(seed=5, n_cells=2, n_alignments=7000, n_loci=40)
    groups = set()
    for t in range(L.n_tiles):
        s = layout_fm.tile_slices(L, t)
        fm = layout_fm.build_fm(L, t, 512, 2, 4)
        groups |= set((fm.rows['meta'] >> 8).tolist())
        x = np.full(s['ncol'], 1.0 / 41) ** 2
        assert np.allclose(layout_fm.iterate_fm(fm, s, x), layout_fm.iterate_direct(s, x), rtol=1e-12, atol=0)
        assert ((fm.cols['cm'] >> 4) & 63).max() >= 8
    assert {1, 2}.issubset(groups)


def test_fm_record_sizes():
    assert layout_fm.ROW_REC.itemsize == 64 and layout_fm.COL_REC.itemsize == 8
    L = _layout(seed=1, n_cells=3, n_alignments=3000)
    fm = layout_fm.build_fm(L, 0, 512, 2, 4)
    assert fm.rows.nbytes == 64 * 2 * 512 and fm.cols.nbytes == 8 * 4 * 512
