"""MatrixMarket output of the count matrices.

Replaces ``scipy.io.mmwrite(out_mtx, _counts, comment=mtx_meta(_rmode))`` at the end of the reference's
``Stellarscope.output_report`` (stellarscope/utils/model.py:1418; metadata comment :1347-1368).  scipy's
writer formats 8-20 M entries per second on the host; for the atlas configuration (70-210 M entries per
reassignment mode) that is 10-25 s per mode, more than the EM fit, the reassignment and the aggregation
together.  Here the text is produced on the device (``ss_mtx_measure`` / ``ss_mtx_write``,
``csrc/ss_mtx.cu``) and streamed into the file through the pinned staging ring.

The writer itself is third-party code outside the reference checkout: ``scipy.io.mmwrite`` (the reference asks for
``scipy>=1.2.1``, setup.py:48, unpinned; 1.18.1 in this image, whose writer is the C++ ``fast_matrix_market`` since
1.12).  Its format is the published MatrixMarket coordinate format: banner ``%%MatrixMarket matrix coordinate
<field> general``, ``%`` comment lines, a ``rows cols entries`` line, one ``row col value`` line (1-based) per stored
entry.  Parity is anchored on the reference's call site (tests/test_mtx.py: files of the real ``output_report``,
scipy's own output, goldens written by scipy).

The file has the same header, comment block, size line, entry order and -- for integer matrices -- the same
bytes as scipy's.  Real values (best_average, ``--umi_counts``) are written as the correctly rounded
17-significant-digit decimal (parses back to the same double) where scipy prints the shortest round-trip
form; the matrices are equal.  The symmetry field is always ``general`` (scipy inspects square matrices for
symmetry; a features x barcodes matrix is a general matrix).
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse

_engine_cache = {}


def _default_engine():
    from .engine import Engine
    if 'e' not in _engine_cache:
        _engine_cache['e'] = Engine()
    return _engine_cache['e']


def header(n_rows, n_cols, nnz, field, comment=''):
    """Banner, comment block and size line exactly as scipy.io.mmwrite writes them (one ``%`` line per line
    of ``comment``)."""
    lines = ['%%MatrixMarket matrix coordinate {} general'.format(field)]
    lines += ['%' + ln for ln in str(comment).split('\n')]
    lines.append('{} {} {}'.format(int(n_rows), int(n_cols), int(nnz)))
    return ('\n'.join(lines) + '\n').encode()


def mmwrite(target, a, comment='', engine=None, device_coo=None):
    """``scipy.io.mmwrite(target, a, comment=comment)`` for a sparse count matrix.  ``target``: path (``.mtx``
    is appended when the name has no extension, like scipy) or an open binary file.  ``a``: scipy sparse matrix,
    integer (-> ``integer``) or floating (-> ``real``).  Entries are written in the storage order of
    ``a.tocoo()`` -- column-major for the CSC matrices of ``output_report``.  ``device_coo``: (row, col, value)
    device tensors holding the entries of ``a`` in that order already (what ``count_by_cell`` leaves on the
    matrix as ``_ss_dev_coo``); without it the triples are uploaded."""
    if not scipy.sparse.issparse(a):
        raise TypeError('mmwrite: a scipy sparse matrix is expected (the count matrices of output_report)')
    if np.issubdtype(a.dtype, np.integer) or a.dtype == np.bool_:
        integer, field = True, 'integer'
    elif np.issubdtype(a.dtype, np.floating):
        integer, field = False, 'real'
    else:
        raise TypeError(f'mmwrite: unsupported dtype {a.dtype}')
    eng = engine if engine is not None and getattr(engine, 'h', None) else _default_engine()
    nnz = int(a.nnz)
    # the body is complete on the device before the file is touched: an unsupported value leaves no partial file
    csc = getattr(a, '_ss_dev_csc', None)
    if device_coo is None and csc is not None and int(csc[1].numel()) == nnz:
        # CSC arrays left on the device by ss_count_cells: expand the column pointers into one cell index per entry
        import torch
        ptr, loci, cnt = csc
        cells = torch.repeat_interleave(torch.arange(ptr.numel() - 1, device=ptr.device, dtype=torch.int32),
                                        (ptr[1:] - ptr[:-1]), output_size=nnz)
        device_coo = (loci, cells, cnt.to(torch.float64))
    if device_coo is not None and all(int(t.numel()) == nnz for t in device_coo):
        text = eng.mtx_text(device_coo[0], device_coo[1], device_coo[2], integer)
    else:
        coo = a.tocoo()
        nnz = int(coo.nnz)
        text = eng.mtx_text(coo.row, coo.col, coo.data, integer)
    from .engine import drain_to_file, write_all
    if hasattr(target, 'write'):
        write_all(target, memoryview(header(a.shape[0], a.shape[1], nnz, field, comment)))
        drain_to_file(text, target)
        return
    path = os.fspath(target)
    if not os.path.splitext(path)[1]:
        path += '.mtx'
    with open(path, 'wb', buffering=0) as fh:
        write_all(fh, memoryview(header(a.shape[0], a.shape[1], nnz, field, comment)))
        drain_to_file(text, fh)
