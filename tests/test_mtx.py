"""MTX writer: scipy.io.mmwrite as called by the reference (model.py:1418) is the
oracle.  CPU part: the line formatter shared by host and device (csrc/ss_mtx_fmt.cuh) against exact decimal
arithmetic, the header against scipy's.  GPU part: whole files against scipy's writer / reader."""
import ctypes as C
import decimal
import io
import os
import subprocess
import tempfile

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')


@pytest.fixture(scope='module')
def fmt():
    d = tempfile.mkdtemp(prefix='ssmtx_')
    so = os.path.join(d, 'mtx_fmt_host.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', os.path.join(HERE, 'tools', 'mtx_fmt_host.cpp'),
                           '-o', so])
    lib = C.CDLL(so)
    lib.ssmtx_put_real.argtypes = [C.c_double, C.c_char_p]
    lib.ssmtx_real_supported.argtypes = [C.c_double]
    lib.ssmtx_put_line.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_char_p]
    lib.ssmtx_line_len.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int]
    return lib


def _real(lib, x):
    buf = C.create_string_buffer(128)
    n = lib.ssmtx_put_real(x, buf)
    return buf.raw[:n].decode()


def _expected_17(x):
    """Exact binary value of x, rounded half-even to 17 significant digits, fixed notation, zeros stripped
    (integers from 2^52 on: all their digits)."""
    if x == 0:
        return '-0' if np.signbit(x) else '0'
    if abs(x) >= 2.0 ** 52:
        return str(int(x))
    d = decimal.Decimal(x)
    with decimal.localcontext() as ctx:
        ctx.prec = 17
        ctx.rounding = decimal.ROUND_HALF_EVEN
        r = +d
    s = format(r, 'f')
    if '.' in s:
        s = s.rstrip('0').rstrip('.')
    return s


def test_real_digits_known(fmt):
    for x, want in [(0.0, '0'), (1.0, '1'), (10.0, '10'), (14.0, '14'), (0.5, '0.5'), (-0.25, '-0.25'),
                    (1 / 3, '0.33333333333333331'), (14 / 3, '4.666666666666667'), (0.1, '0.10000000000000001'),
                    (2.0 ** 53, '9007199254740992'), (2.0 ** 62, '4611686018427387904'),
                    (2.0 ** -64, '0.000000000000000000054210108624275222'),
                    (9.999999999999999999, '10'), (0.99999999999999989, '0.99999999999999989'),
                    (123456789.125, '123456789.125')]:
        assert _real(fmt, x) == want, (x, _real(fmt, x), want)


def test_real_digits_random(fmt):
    rng = np.random.default_rng(5)
    xs = np.concatenate([
        rng.random(4000), rng.random(2000) * 1e6, 1.0 / rng.integers(1, 2049, 2000),
        rng.integers(1, 50, 2000) / rng.integers(1, 33, 2000),
        np.ldexp(rng.random(3000) + 0.5, rng.integers(-63, 63, 3000)),
        -np.ldexp(rng.random(500) + 0.5, rng.integers(-63, 63, 500)),
        np.nextafter(np.arange(1, 200, dtype=np.float64), 0), np.nextafter(np.arange(1, 200, dtype=np.float64), 1e9),
        # halfway cases of the 17th digit: k + 0.5 ulp patterns around 2^53 / 10^16
        np.arange(2.0 ** 52, 2.0 ** 52 + 300) / 2.0, np.arange(2.0 ** 52, 2.0 ** 52 + 300) / 8.0,
    ])
    for x in xs:
        x = float(x)
        assert fmt.ssmtx_real_supported(x)
        s = _real(fmt, x)
        assert float(s) == x, (x, s)
        assert s == _expected_17(x), (x, s, _expected_17(x))


def test_real_range(fmt):
    for x in (np.inf, -np.inf, np.nan, 2.0 ** 63, 2.0 ** -65, 5e-324, 1e300):
        assert not fmt.ssmtx_real_supported(float(x))
    for x in (0.0, -0.0, 2.0 ** -64, np.nextafter(2.0 ** 63, 0)):
        assert fmt.ssmtx_real_supported(float(x))


def test_lines(fmt):
    buf = C.create_string_buffer(128)
    for r, c, v, integer, want in [(0, 0, 1.0, 1, '1 1 1\n'), (62, 0, 30.0, 1, '63 1 30\n'),
                                   (28000, 99999, 2147483647.0, 1, '28001 100000 2147483647\n'),
                                   (2147483646, 2147483646, -17.0, 1, '2147483647 2147483647 -17\n'),
                                   (5, 7, 0.5, 0, '6 8 0.5\n'), (5, 7, 3.0, 0, '6 8 3\n')]:
        n = fmt.ssmtx_put_line(r, c, v, integer, buf)
        assert buf.raw[:n].decode() == want
        assert fmt.ssmtx_line_len(r, c, v, integer) == len(want)
    rng = np.random.default_rng(9)
    for _ in range(3000):                                   # integers across all digit counts, both signs
        r, c = (int(x) for x in rng.integers(0, 2 ** 31 - 1, 2) >> rng.integers(0, 31, 2))
        v = float(int(rng.integers(-2 ** 52, 2 ** 52)) >> int(rng.integers(0, 52)))
        n = fmt.ssmtx_put_line(r, c, v, 1, buf)
        assert buf.raw[:n].decode() == f'{r + 1} {c + 1} {int(v)}\n'
        assert fmt.ssmtx_line_len(r, c, v, 1) == n
        x = float(rng.integers(1, 50) / rng.integers(1, 33))
        assert fmt.ssmtx_line_len(r, c, x, 0) == fmt.ssmtx_put_line(r, c, x, 0, buf)
    assert fmt.ssmtx_max_line() >= 11 + 1 + 11 + 1 + 58 + 1


def test_header_matches_scipy():
    from stellarscope_b200 import mtx
    comment = 'PN: stellarscope\n VN: 1.5\n pooling_mode: individual\n reassign_mode: best_exclude\n command: x y z'
    for dtype, field in ((np.int32, 'integer'), (np.float64, 'real')):
        m = sp.csc_matrix((np.array([3], dtype=dtype), (np.array([2]), np.array([4]))), shape=(7, 5))
        f = io.BytesIO()
        scipy.io.mmwrite(f, m, comment=comment)
        assert mtx.header(7, 5, 1, field, comment) + b'3 5 3\n' == f.getvalue()


def test_header_comment_edge_cases_match_scipy():
    from stellarscope_b200 import mtx
    m = sp.csc_matrix((np.array([3], dtype=np.int32), (np.array([2]), np.array([4]))), shape=(7, 5))
    for comment in ('', 'one', 'a\nb', 'a\n\nb', 'trailing\n'):
        f = io.BytesIO()
        scipy.io.mmwrite(f, m, comment=comment)
        assert mtx.header(7, 5, 1, 'integer', comment) + b'3 5 3\n' == f.getvalue(), repr(comment)


def test_mmwrite_rejects_what_it_does_not_write():
    """Argument errors come before any device work (no GPU needed)."""
    from stellarscope_b200 import mtx
    with pytest.raises(TypeError):
        mtx.mmwrite('x.mtx', np.ones((2, 2)))                          # dense: not a count matrix of output_report
    with pytest.raises(TypeError):
        mtx.mmwrite('x.mtx', sp.csc_matrix(np.ones((2, 2), dtype=np.complex128)))
    assert not os.path.exists('x.mtx')


def test_write_all_handles_short_writes():
    from stellarscope_b200.engine import write_all

    class Short:                                    # an unbuffered sink that takes at most 5 bytes per call
        def __init__(self):
            self.got = bytearray()

        def write(self, mv):
            k = min(5, len(mv))
            self.got += bytes(mv[:k])
            return k
    sink = Short()
    data = bytes(range(256)) * 3
    write_all(sink, memoryview(data))
    assert bytes(sink.got) == data
    buf = io.BytesIO()
    write_all(buf, memoryview(data))
    assert buf.getvalue() == data

    class Stuck:
        def write(self, mv):
            return 0
    with pytest.raises(OSError):
        write_all(Stuck(), memoryview(b'abc'))


def test_golden_files_are_scipy():
    """The committed goldens are what scipy.io.mmwrite (the reference's writer) produces for the committed
    triples (tests/tools/make_golden_mtx.py)."""
    z = np.load(os.path.join(GOLDEN, 'mtx_cases.npz'))
    for name, dtype in (('int', np.int32), ('real', np.float64)):
        m = sp.csc_matrix((z[f'{name}_data'].astype(dtype), z[f'{name}_indices'], z[f'{name}_indptr']),
                          shape=tuple(z['shape']))
        f = io.BytesIO()
        scipy.io.mmwrite(f, m, comment=str(z['comment']))
        with open(os.path.join(GOLDEN, f'mtx_{name}.mtx'), 'rb') as fh:
            assert fh.read() == f.getvalue()


# ---- GPU -------------------------------------------------------------------------------------------------

def _random_counts(rng, K, Cn, n, real):
    r = rng.integers(0, K, n)
    c = rng.integers(0, Cn, n)
    if real:
        v = rng.integers(1, 40, n) / rng.integers(1, 33, n)
    else:
        v = rng.integers(1, 100000, n).astype(np.int32)
    m = sp.coo_matrix((v, (r, c)), shape=(K, Cn)).tocsc()
    m.sum_duplicates()
    m.sort_indices()
    return m.astype(np.float64 if real else np.int32)


@pytest.mark.gpu
def test_gpu_mtx_integer_bytes_equal_scipy(tmp_path):
    from stellarscope_b200 import mtx
    rng = np.random.default_rng(1)
    comment = 'PN: stellarscope\n VN: t\n reassign_mode: best_exclude'
    for K, Cn, n in ((101, 37, 1), (101, 37, 900), (28001, 1200, 300000), (3, 2500, 5000)):
        m = _random_counts(rng, K, Cn, n, real=False)
        a, b = str(tmp_path / 'a.mtx'), str(tmp_path / 'b.mtx')
        scipy.io.mmwrite(a, m, comment=comment)
        mtx.mmwrite(b, m, comment=comment)
        with open(a, 'rb') as fa, open(b, 'rb') as fb:
            assert fa.read() == fb.read(), (K, Cn, n)
    # an empty matrix: header only (scipy labels an empty integer matrix "real"; the content is what counts)
    e = sp.csc_matrix((101, 37), dtype=np.int32)
    mtx.mmwrite(str(tmp_path / 'e.mtx'), e, comment=comment)
    back = scipy.io.mmread(str(tmp_path / 'e.mtx'))
    assert back.shape == (101, 37) and back.nnz == 0


@pytest.mark.gpu
def test_gpu_mtx_real_content_equal(tmp_path):
    from stellarscope_b200 import mtx
    rng = np.random.default_rng(2)
    for K, Cn, n in ((101, 37, 900), (28001, 1200, 200000)):
        m = _random_counts(rng, K, Cn, n, real=True)
        b = str(tmp_path / 'b.mtx')
        mtx.mmwrite(b, m, comment='x')
        back = scipy.io.mmread(b).tocsc()
        back.sort_indices()
        assert back.shape == m.shape and back.nnz == m.nnz
        assert (back.indptr == m.indptr).all() and (back.indices == m.indices).all()
        assert (back.data == m.data).all()                 # bit-exact round trip
        with open(b, 'rb') as fh:
            assert fh.readline() == b'%%MatrixMarket matrix coordinate real general\n'


@pytest.mark.gpu
def test_gpu_mtx_golden(tmp_path):
    from stellarscope_b200 import mtx
    z = np.load(os.path.join(GOLDEN, 'mtx_cases.npz'))
    m = sp.csc_matrix((z['int_data'].astype(np.int32), z['int_indices'], z['int_indptr']), shape=tuple(z['shape']))
    b = str(tmp_path / 'g.mtx')
    mtx.mmwrite(b, m, comment=str(z['comment']))
    with open(b, 'rb') as fb, open(os.path.join(GOLDEN, 'mtx_int.mtx'), 'rb') as fg:
        assert fb.read() == fg.read()
    mr = sp.csc_matrix((z['real_data'], z['real_indices'], z['real_indptr']), shape=tuple(z['shape']))
    mtx.mmwrite(b, mr, comment=str(z['comment']))
    g = scipy.io.mmread(os.path.join(GOLDEN, 'mtx_real.mtx')).tocsc()
    o = scipy.io.mmread(b).tocsc()
    assert (g != o).nnz == 0 and (o.data == g.data).all()


@pytest.mark.gpu
def test_gpu_mtx_rejects_unsupported(tmp_path):
    from stellarscope_b200 import mtx
    from stellarscope_b200.engine import EngineError
    m = sp.csc_matrix(np.array([[1.0, 0.0], [0.0, 1e300]]))
    with pytest.raises(EngineError):
        mtx.mmwrite(str(tmp_path / 'x.mtx'), m, comment='x')
    assert not os.path.exists(str(tmp_path / 'x.mtx'))


@pytest.mark.gpu
def test_gpu_mtx_csr_and_large_staging(tmp_path):
    """CSR input (row-major order like scipy's) and a body large enough for the chunked copy-out."""
    from stellarscope_b200 import mtx
    rng = np.random.default_rng(3)
    m = _random_counts(rng, 28001, 4000, 1500000, real=False).tocsr()
    a, b = str(tmp_path / 'a.mtx'), str(tmp_path / 'b.mtx')
    scipy.io.mmwrite(a, m, comment='c')
    mtx.mmwrite(b, m, comment='c')
    with open(a, 'rb') as fa, open(b, 'rb') as fb:
        assert fa.read() == fb.read()


@pytest.mark.skipif(not os.path.isdir('/root/reference/stellarscope'), reason='reference checkout not present')
def test_header_and_metadata_equal_the_real_output_report(tmp_path, monkeypatch):
    """The real ``Stellarscope.output_report`` (unmodified reference, CPU) writes TE_counts*.mtx; banner, metadata
    comment (mtx_meta, model.py:1347-1368) and size line of this package's writer must be the same bytes."""
    import sys
    from oracle import make_golden, refshim
    from stellarscope_b200 import model, mtx, synth
    ref = refshim.load()
    monkeypatch.setattr(sys, 'argv', ['stellarscope', 'assign', '--pooling_mode', 'individual', 'aln.bam', 'te.gtf'])
    d = synth.generate(n_cells=5, n_alignments=900, n_loci=60, seed=31, unique_frac=0.1)
    for extra in (dict(), dict(ignore_umi=True, umi_counts=True), dict(samfile='aln.bam', gtffile='te.gtf')):
        opts = refshim.RefOpts(reassign_mode=['best_exclude', 'best_average'], pooling_mode='individual',
                               version='1.4.2', **extra)
        opts.outfile_path = lambda suffix: str(tmp_path / suffix)
        st, names = make_golden.make_st(d, opts, False)
        with refshim.cli_fp_policy():
            tl, poolinfo = ref._fit_pooling_model(st, opts)
            ref.Stellarscope.reassign(st, tl)
            ref.Stellarscope.output_report(st, tl)
        for i, mode in enumerate(opts.reassign_mode):
            fn = tmp_path / ('TE_counts.mtx' if i == 0 else f'TE_counts.{mode}.mtx')
            raw = fn.read_bytes()
            lines = raw.split(b'\n')
            n_head = next(k for k, ln in enumerate(lines) if not ln.startswith(b'%')) + 1
            real_head = b'\n'.join(lines[:n_head]) + b'\n'
            K, C, nnz = (int(x) for x in lines[n_head - 1].split())
            field = lines[0].split()[3].decode()
            assert field == ('real' if (mode == 'best_average' or opts.umi_counts) else 'integer')
            assert mtx.header(K, C, nnz, field, model._mtx_meta(st, mode)) == real_head, (extra, mode)
