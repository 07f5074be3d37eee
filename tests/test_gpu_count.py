"""ss_count_cells (sort-free per-cell counts, csrc/ss_count.cu) against ss_count (sort + run-length) and numpy:
contiguous and scattered cells, dropped rows, empty cells, cells with more distinct loci than the touched-bin
list holds (dense scan of the bins)."""
import numpy as np
import pytest

from stellarscope_b200 import synth

pytestmark = pytest.mark.gpu


def _numpy_counts(d, flag, row_cell, n_cells):
    rows = np.repeat(np.arange(d.N), np.diff(d.indptr))
    keep = (flag != 0) & (row_cell[rows] >= 0) & (row_cell[rows] < n_cells)
    out = np.zeros((d.K, n_cells), dtype=np.int64)
    np.add.at(out, (d.indices[keep], row_cell[rows][keep]), 1)
    return out


@pytest.mark.parametrize('case', ['contiguous', 'scattered', 'dropped', 'wide_cells'])
def test_count_cells_matches_sort_path(case):
    import torch
    from stellarscope_b200.engine import Engine
    rng = np.random.default_rng(3)
    if case == 'wide_cells':
        # few huge cells over many loci: > 8192 distinct loci per cell -> dense scan of the bins
        d = synth.generate(n_cells=3, n_alignments=400_000, n_loci=28_000, seed=1)
    else:
        d = synth.generate(n_cells=300, n_alignments=200_000, n_loci=3000, seed=2)
    n_cells = d.n_cells + 2                                    # two output columns without any row
    row_cell = d.row_cell.copy()
    if case == 'scattered':
        row_cell = rng.permutation(row_cell)                   # rows of a cell all over the matrix (BAM order)
    if case == 'dropped':
        row_cell[rng.random(d.N) < 0.2] = -1                   # reads of barcodes outside the filter list
        row_cell[:50] = -1
    flag_h = (rng.random(d.A) < 0.6).astype(np.uint8)
    eng = Engine(0)
    eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
    flag = torch.from_numpy(flag_h).to(eng.device)
    ptr, loci, cnt = eng.count_cells(flag, row_cell, n_cells)
    exp = _numpy_counts(d, flag_h, row_cell, n_cells)
    assert ptr[0] == 0 and ptr[-1] == len(loci) == len(cnt) == np.count_nonzero(exp)
    got = np.zeros_like(exp)
    cells = np.repeat(np.arange(n_cells), np.diff(ptr))
    got[loci, cells] = cnt
    assert np.array_equal(got, exp)
    for c in range(n_cells):                                   # loci ascending inside every cell (CSC order)
        assert np.all(np.diff(loci[ptr[c]:ptr[c + 1]]) > 0)
    if case == 'wide_cells':
        assert np.diff(ptr).max() > 8192
    cell2, col2, val2 = eng.count(flag, np.where(row_cell < n_cells, row_cell, -1))
    assert np.array_equal(cell2, cells) and np.array_equal(col2, loci) and np.array_equal(val2.astype(np.int64), cnt)
    # nothing selected / no rows with a cell
    z = torch.zeros_like(flag)
    ptr0, loci0, cnt0 = eng.count_cells(z, row_cell, n_cells)
    assert ptr0[-1] == 0 and len(loci0) == 0
    ptr1, loci1, _ = eng.count_cells(flag, np.full(d.N, -1, dtype=np.int32), n_cells)
    assert ptr1[-1] == 0 and len(loci1) == 0
    eng.close()
