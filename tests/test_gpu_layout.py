"""Device layout builder (ss_layout_build) against the numpy format specification."""
import numpy as np
import pytest

from stellarscope_b200 import layout_host, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

EXACT = [('tile_hdr', np.int64), ('e_lcol', np.uint16), ('e_cpos', np.uint16), ('e_opos', np.int32),
         ('r_len', np.uint16), ('r_orow', np.int32), ('c_ptr', np.uint16), ('c_scol', np.int32),
         ('c_nuniq', np.int32), ('j_ptr', np.uint16), ('seg_tile0', np.int32), ('seg_ntiles', np.int32),
         ('seg_col0', np.int32), ('seg_ncol', np.int32), ('seg_smax', np.int32), ('seg_nobs', np.int32),
         ('seg_namb', np.int32), ('seg_nuniq', np.int32), ('seg_A', np.int64), ('seg_Aamb', np.int64),
         ('sc_gcol', np.int32), ('sc_nuniq', np.int32), ('sc_inamb', np.int8), ('u_row', np.int32),
         ('u_pos', np.int32)]
# floating sums may be accumulated in a different (fixed) order on the device
CLOSE = [('e_q', 1e-15), ('r_w', 1e-15), ('c_pisum0', 1e-13), ('seg_wmax', 1e-15), ('seg_wtot', 1e-12),
         ('seg_wamb', 1e-12), ('seg_logq_uniq', 1e-12), ('sc_pisum0', 1e-13)]


def _check(d, row_seg, S):
    from stellarscope_b200.engine import Engine
    hl = layout_host.build(d.indptr, d.indices, d.scores, row_seg, S, d.K)
    eng = Engine()
    eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
    eng.build_layout(row_seg, S)
    info = eng.layout_info()
    assert info['n_tiles'] == hl.n_tiles and info['chunk'] == hl.chunk and info['max_row_len'] == hl.maxlen
    assert info['n_segcols'] == len(hl.sc_gcol) and info['n_unique_rows'] == len(hl.u_row)
    for name, dt in EXACT:
        got = eng.export_layout_array(name, dt)
        exp = np.asarray(getattr(hl, name)).astype(dt).ravel()
        assert got.shape == exp.shape, name
        assert np.array_equal(got, exp), name
    for name, rtol in CLOSE:
        got = eng.export_layout_array(name, np.float64)
        exp = np.asarray(getattr(hl, name), dtype=np.float64).ravel()
        assert got.shape == exp.shape, name
        assert np.allclose(got, exp, rtol=rtol, atol=0), name
    eng.close()


def test_build_individual():
    d = H.small_synth(seed=1)
    _check(d, d.row_cell, d.n_cells)


def test_build_pseudobulk_multi_tile():
    d = H.small_synth(seed=2, n_alignments=20000)
    _check(d, np.zeros(d.N, dtype=np.int32), 1)


def test_build_celltype_unsorted_segments():
    d = H.small_synth(seed=3, n_cells=30, n_alignments=15000, n_celltypes=4)
    _check(d, d.cell_type[d.row_cell].astype(np.int32), d.n_celltypes)


def test_build_rows_outside_models_and_empty_rows():
    d = H.small_synth(seed=4, umi_dup_rate=0.2)
    dm = synth.apply_row_mask(d, synth.umi_dedup_mask(d))       # zeroed rows
    seg = d.row_cell.copy()
    seg[seg % 3 == 0] = -1                                       # cells without a model
    _check(dm, seg, d.n_cells)


def test_build_all_unique_and_single_row_cells():
    d = H.small_synth(seed=5, unique_frac=1.0, n_cells=8, n_alignments=300)
    _check(d, d.row_cell, d.n_cells)
    d = H.small_synth(seed=6, n_cells=50, n_alignments=150)      # many one-row cells
    _check(d, d.row_cell, d.n_cells)


def test_build_long_rows():
    d = H.small_synth(seed=7, n_loci=3000, n_alignments=30000, max_k=32)
    _check(d, d.row_cell, d.n_cells)
