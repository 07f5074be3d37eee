"""Import the *real* reference (``/root/reference``) in the build container.

TEST INFRASTRUCTURE, build container only: ``/root/reference`` does not exist
on the GPU box, so nothing that runs there may import this module.  It is used
by ``oracle/make_golden.py`` (writes ``tests/golden/*.npz``) and by the
``-m "not gpu"`` tests that validate the numpy restatement against the
reference when the checkout is present.

The reference imports pysam / intervaltree / future / past / its Cython
extension at module top level (model.py:7-45, sparse_plus.py:11-12); none of
them is touched by the EM / reassign code, so stub modules are enough
(SURVEY.md Appendix C).
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get('STELLARSCOPE_REFERENCE', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'stellarscope'))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """Returns the reference's ``stellarscope.utils.model`` module."""
    if 'stellarscope.utils.model' in sys.modules:
        return sys.modules['stellarscope.utils.model']
    if not available():
        raise RuntimeError(f'reference checkout not found at {REFERENCE_ROOT}')
    for name, attrs in (
        ('pysam', dict(AlignedSegment=object, AlignmentFile=object, IteratorRow=object,
                       set_verbosity=lambda v: 0, FSECONDARY=256)),
        ('intervaltree', dict(Interval=object, IntervalTree=object)),
    ):
        if name not in sys.modules:
            _stub(name, **attrs)
    if 'future' not in sys.modules:
        f = _stub('future')
        f.standard_library = _stub('future.standard_library', install_aliases=lambda: None)
    if 'past' not in sys.modules:
        p = _stub('past')
        p.utils = _stub('past.utils', old_div=lambda a, b: a / b)
    _stub('stellarscope.utils.calignment', AlignedPair=object)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import stellarscope.utils.model as model  # noqa: E402
    return model


class cli_fp_policy:
    """``np.seterr`` exactly as the CLI sets it (__main__.py:30)."""

    def __enter__(self):
        self._prev = np.seterr(divide='raise', over='raise', under='raise', invalid='raise')

    def __exit__(self, *exc):
        np.seterr(**self._prev)


class RefOpts:
    """Minimal ``opts`` object for TelescopeLikelihood / _fit_pooling_model."""

    def __init__(self, **kw):
        self.em_epsilon = 1e-7
        self.max_iter = 100
        self.pi_prior = 0
        self.theta_prior = 200000
        self.use_likelihood = False
        self.pooling_mode = 'pseudobulk'
        self.reassign_mode = ['best_exclude']
        self.conf_prob = 0.9
        self.ignore_umi = True
        self.umi_counts = False
        self.nproc = 1
        self.rng = np.random.default_rng(0)
        self.version = '0+unknown'
        self.__dict__.update(kw)
