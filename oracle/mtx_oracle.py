"""TEST INFRASTRUCTURE -- CPU restatement of the MatrixMarket output of the count matrices.

The reference writes them with ``scipy.io.mmwrite(out_mtx, _counts, comment=mtx_meta(_rmode))``
(stellarscope/utils/model.py:1418; metadata comment :1347-1368).  scipy is third-party (``scipy>=1.2.1``,
setup.py:48, unpinned; 1.18.1 here); its output follows the published MatrixMarket coordinate format, restated
below in plain Python:

    %%MatrixMarket matrix coordinate <integer|real> general
    %<comment line>            one per line of the comment
    <rows> <cols> <entries>
    <row> <col> <value>        1-based, one per stored entry, in the storage order of ``a.tocoo()``

Integers print as decimal integers (byte-identical to scipy's file).  Reals: scipy prints the shortest decimal
that round-trips ("1.4E1"); this restatement -- like the device formatter it checks, csrc/ss_mtx_fmt.cuh --
prints the exact binary value rounded half-even to 17 significant digits in fixed notation without trailing zeros
(integers from 2^52 on: all digits), which parses back to the same double.  Only tests may import this module.
"""
import decimal

import numpy as np


def real_text(x):
    x = float(x)
    if x == 0:
        return '-0' if np.signbit(x) else '0'
    if abs(x) >= 2.0 ** 52:
        return str(int(x))
    with decimal.localcontext() as ctx:
        ctx.prec = 17
        ctx.rounding = decimal.ROUND_HALF_EVEN
        r = +decimal.Decimal(x)
    s = format(r, 'f')
    return s.rstrip('0').rstrip('.') if '.' in s else s


def mtx_bytes(a, comment=''):
    """The whole file for a scipy sparse matrix ``a`` (integer or floating dtype)."""
    coo = a.tocoo()
    integer = np.issubdtype(a.dtype, np.integer)
    lines = ['%%MatrixMarket matrix coordinate {} general'.format('integer' if integer else 'real')]
    lines += ['%' + ln for ln in str(comment).split('\n')]
    lines.append(f'{a.shape[0]} {a.shape[1]} {coo.nnz}')
    fmt = (lambda v: str(int(v))) if integer else real_text
    lines += [f'{int(r) + 1} {int(c) + 1} {fmt(v)}' for r, c, v in zip(coo.row, coo.col, coo.data)]
    return ('\n'.join(lines) + '\n').encode()
