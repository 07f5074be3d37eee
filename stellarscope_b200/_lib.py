"""ctypes binding of libstellarscope_b200.so (include/stellarscope_b200.h).

There is no CPU fallback: if the shared library is missing it is built with
nvcc; if that fails, importing the engine fails.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))

SS_OK, SS_ERR_ARG, SS_ERR_CUDA, SS_ERR_STATE, SS_ERR_MODEL, SS_ERR_CAPACITY, SS_ERR_NOT_CANONICAL = range(7)
ABI_VERSION = 1

MODES = {  # order of TelescopeLikelihood.REASSIGN_MODES (reference model.py:92-100)
    'best_exclude': 0, 'best_conf': 1, 'best_random': 2, 'best_average': 3,
    'initial_unique': 4, 'initial_random': 5, 'total_hits': 6,
}

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u16p = C.POINTER(C.c_uint16)
c_u8p = C.POINTER(C.c_uint8)
c_i8p = C.POINTER(C.c_int8)
c_f64p = C.POINTER(C.c_double)


class LayoutInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        'n_rows', 'n_cols', 'n_segments', 'n_tiles', 'max_row_len', 'chunk',
        'n_unique_rows', 'n_resident_tiles', 'n_stream_tiles', 'n_stream_segments',
        'n_cluster_tiles', 'n_cluster_segments')] + \
        [(n, C.c_int64) for n in (
            'nnz', 'n_segcols', 'tot_ent', 'tot_row', 'tot_col', 'tot_jds',
            'nnz_ambiguous', 'rows_ambiguous')]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class HostLayoutC(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in ('N', 'K', 'S', 'n_tiles', 'maxlen', 'chunk', 'n_uniq')] +
        [(n, C.c_int64) for n in ('A', 'n_segcols', 'tot_ent', 'tot_row', 'tot_col', 'tot_jds')] +
        [('tile_hdr', c_i64p),
         ('e_lcol', c_u16p), ('e_cpos', c_u16p), ('e_q', c_f64p), ('e_opos', c_i32p),
         ('r_len', c_u16p), ('r_w', c_f64p), ('r_orow', c_i32p),
         ('c_ptr', c_u16p), ('c_scol', c_i32p), ('c_pisum0', c_f64p), ('c_nuniq', c_i32p),
         ('j_ptr', c_u16p),
         ('seg_tile0', c_i32p), ('seg_ntiles', c_i32p), ('seg_col0', c_i32p), ('seg_ncol', c_i32p),
         ('seg_smax', c_i32p), ('seg_nobs', c_i32p), ('seg_namb', c_i32p), ('seg_nuniq', c_i32p),
         ('seg_wmax', c_f64p), ('seg_wtot', c_f64p), ('seg_wamb', c_f64p), ('seg_logq_uniq', c_f64p),
         ('seg_A', c_i64p), ('seg_Aamb', c_i64p),
         ('sc_gcol', c_i32p), ('sc_pisum0', c_f64p), ('sc_nuniq', c_i32p), ('sc_inamb', c_i8p),
         ('u_row', c_i32p), ('u_pos', c_i32p)])


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, c_f64p, C.c_int64, C.c_int32)

# name -> (restype, argtypes); every symbol include/stellarscope_b200.h declares
SIGNATURES = {
    'ss_abi_version': (C.c_int, []),
    'ss_create': (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    'ss_destroy': (C.c_int, [C.c_void_p]),
    'ss_release': (C.c_int, [C.c_void_p]),
    'ss_last_error': (C.c_char_p, [C.c_void_p]),
    'ss_layout_build': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    'ss_layout_import': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(HostLayoutC)]),
    'ss_layout_get_info': (C.c_int, [C.c_void_p, C.POINTER(LayoutInfo)]),
    'ss_layout_export': (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, c_i64p]),
    'ss_em_fit': (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_double, C.c_double, C.c_int32]),
    'ss_em_get_results': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_void_p] * 7),
    'ss_em_get_params': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_void_p] * 7),
    'ss_em_emit_z': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'ss_reassign': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'ss_reassign_csr': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_void_p, c_i64p]),
    'ss_count': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                           c_i64p]),
    'ss_reassign_pick': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'ss_count_cells': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, c_i64p]),
    'ss_umi_dedup': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'ss_scores_build': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                 C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, c_i64p]),
    'ss_mtx_measure': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                c_i64p]),
    'ss_mtx_write': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                              C.c_void_p]),
    'ss_set_allreduce': (C.c_int, [C.c_void_p, ALLREDUCE_FN, C.c_void_p]),
    'ss_peer_buffer_bytes': (C.c_int64, [C.c_int32]),
    'ss_set_peers': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_int64]),
    'ss_em_get_dense_params': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'ss_launch_count': (C.c_int64, [C.c_void_p]),
    'ss_lane_classes': (C.c_int, [c_i32p, C.c_int32]),
    'ss_set_timing': (C.c_int, [C.c_void_p, C.c_int32]),
    'ss_get_timings': (C.c_int, [C.c_void_p, c_f64p, C.c_int32]),
}

_lib = None


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load the shared library (building it first when it is missing or stale
    and nvcc is available).  Raises if it cannot be loaded -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and _build.needs_build():
        try:
            _build.build()
        except Exception:
            if not os.path.exists(path):
                raise
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.ss_abi_version() != ABI_VERSION:
        raise RuntimeError('libstellarscope_b200.so ABI version mismatch')
    _lib = lib
    return lib
