"""The C-ABI library loads and exports every symbol include/stellarscope_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from stellarscope_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'stellarscope_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ss_[a-z_0-9]+)\s*\(', src)) - {'ss_allreduce_fn'})


def test_library_builds_and_loads():
    lib = _lib.load()
    assert os.path.exists(_lib.lib_path())
    assert lib.ss_abi_version() == _lib.ABI_VERSION


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 15
    raw = ctypes.CDLL(_lib.lib_path())
    for n in names:
        assert hasattr(raw, n), f'{n} declared in the header but not exported'
        assert n in _lib.SIGNATURES, f'{n} has no ctypes signature in _lib.py'
    for n in _lib.SIGNATURES:
        assert n in names, f'{n} bound in _lib.py but not declared in the header'


def test_struct_layouts_match_header():
    # ss_layout_info: 12 int32 + 8 int64; ss_host_layout: 7 int32 (+pad) + 6 int64 + 33 pointers
    assert ctypes.sizeof(_lib.LayoutInfo) == 12 * 4 + 8 * 8
    assert ctypes.sizeof(_lib.HostLayoutC) == 8 * 4 + 6 * 8 + 33 * 8


def test_null_handle_is_an_error_not_a_crash():
    lib = _lib.load()
    assert lib.ss_destroy(None) == _lib.SS_ERR_ARG
    assert lib.ss_last_error(None) == b'null handle'
    assert lib.ss_launch_count(None) == 0


def test_engine_refuses_to_run_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from stellarscope_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine()
