"""Thin host wrapper over the C ABI: owns the handle and the torch device
buffers that are handed to it (PyTorch is only the allocator / stream provider).

All arithmetic happens in libstellarscope_b200.so; nothing here computes the
model on the CPU.  Using this module without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib
from ._lib import MODES


class EngineError(RuntimeError):
    """Failure inside the native library (status code in ``.status``)."""

    def __init__(self, status, msg):
        super().__init__(f'[ss status {status}] {msg}')
        self.status = status


class NotCanonicalError(EngineError):
    """SS_ERR_NOT_CANONICAL -- ss_layout_build found zero scores or unsorted / duplicate column
    indices; the caller canonicalises the matrix on the host (what the reference does on
    construction, model.py:113-114) and builds again."""


class ModelError(EngineError):
    """SS_ERR_MODEL -- the class of failures the reference raises as
    ``StellarscopeError`` (model.py:267-272, 454-457)."""


class LazyTrace:
    """(S, max_iter) trace that lives on the device (a private copy, so later fits do not change it)
    until somebody indexes it; behaves like the numpy array afterwards."""

    def __init__(self, dev_tensor, S, MI):
        self._dev, self._shape, self._host = dev_tensor, (S, MI), None

    def numpy(self):
        if self._host is None:
            self._host = self._dev.cpu().numpy().reshape(self._shape)
            self._dev = None
        return self._host

    shape = property(lambda self: self._shape)

    def __getitem__(self, k):
        return self.numpy()[k]

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return self._shape[0]

    def __getstate__(self):                   # picklable (multi-GPU gather of the per-model records)
        return {'_dev': None, '_shape': self._shape, '_host': self.numpy()}


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _HostStager:
    """Device -> host copies of large results.  A plain ``.cpu()`` lands in freshly allocated pageable memory at
    2-4 GB/s: the destination's page faults and the driver's staging copy run on one thread.  Here the result is
    cut into chunks that travel through a small ring of pinned buffers (allocated once per process) and are
    copied into the destination array by a few worker threads, so the PCIe transfer and the page faults of
    several chunks overlap.  The returned arrays are ordinary numpy arrays (no pinned memory is handed out)."""
    CHUNK = 8 << 20
    SLOTS = 8
    MIN_BYTES = 4 << 20          # below this a plain copy is as fast

    def __init__(self):
        self.ring = None
        self.lock = threading.Lock()
        self.pool = None

    def _ensure(self):
        if self.ring is None:
            self.ring = torch.empty(self.SLOTS * self.CHUNK, dtype=torch.uint8, pin_memory=True)
            self.ring_np = self.ring.numpy()
            self.events = [torch.cuda.Event() for _ in range(self.SLOTS)]
            import os
            try:
                nthr = int(os.environ.get('SS_HOST_THREADS', self.SLOTS))
            except ValueError:
                nthr = self.SLOTS
            self.pool = ThreadPoolExecutor(max_workers=max(2, min(self.SLOTS, nthr)), thread_name_prefix='ss_d2h')

    def fetch(self, t):
        n = t.numel() * t.element_size()
        if not t.is_cuda or n < self.MIN_BYTES:
            return t.cpu().numpy()
        with self.lock, torch.cuda.device(t.device):
            self._ensure()
            src = t.contiguous().reshape(-1).view(torch.uint8)
            out = np.empty(tuple(t.shape), dtype=_TORCH_NP[t.dtype])
            dst = out.reshape(-1).view(np.uint8)
            stream = torch.cuda.current_stream(t.device)
            pending = [None] * self.SLOTS
            C_ = self.CHUNK

            def drain(slot, off, m, ev):
                ev.synchronize()
                dst[off:off + m] = self.ring_np[slot * C_:slot * C_ + m]

            for i, off in enumerate(range(0, n, C_)):
                slot = i % self.SLOTS
                if pending[slot] is not None:
                    pending[slot].result()                       # the slot has been copied out
                m = min(C_, n - off)
                self.ring[slot * C_:slot * C_ + m].copy_(src[off:off + m], non_blocking=True)
                self.events[slot].record(stream)
                pending[slot] = self.pool.submit(drain, slot, off, m, self.events[slot])
            for f in pending:
                if f is not None:
                    f.result()
            return out


    def drain_to_file(self, t, fh):
        """Device bytes -> open binary file, in order, through the pinned ring: up to SLOTS chunk copies are in
        flight while the oldest one is written (no host copy of the whole text is ever made)."""
        src = t.contiguous().reshape(-1).view(torch.uint8)
        n = src.numel()
        if n == 0:
            return 0
        with self.lock, torch.cuda.device(t.device):
            self._ensure()
            stream = torch.cuda.current_stream(t.device)
            C_ = self.CHUNK
            offs = list(range(0, n, C_))

            def issue(i):
                slot, off = i % self.SLOTS, offs[i]
                m = min(C_, n - off)
                self.ring[slot * C_:slot * C_ + m].copy_(src[off:off + m], non_blocking=True)
                self.events[slot].record(stream)

            for i in range(min(self.SLOTS, len(offs))):
                issue(i)
            for i, off in enumerate(offs):
                slot = i % self.SLOTS
                m = min(C_, n - off)
                self.events[slot].synchronize()
                write_all(fh, memoryview(self.ring_np[slot * C_:slot * C_ + m]))
                if i + self.SLOTS < len(offs):
                    issue(i + self.SLOTS)                    # the slot is free again
        return n

    def push(self, a, device, filler=None, shape=None, dtype=None):
        """Host array -> new device tensor through the ring (the mirror image of ``fetch``): worker threads copy
        chunks into the pinned slots while earlier slots are on their way.  For uploads too large to pin as a
        whole -- pinning a multi-GB matrix costs more than the transfer and the pinned block would stay cached.
        With ``filler(dst_bytes, byte_offset, n_bytes)`` the chunks are PRODUCED into the slots instead of copied
        from ``a`` (``shape`` / ``dtype`` of the result given): the rows a rank cuts out of the host matrix go
        straight into the staging ring, the compacted copy never exists on the host."""
        if filler is None:
            a = np.ascontiguousarray(a)
            shape, dtype = tuple(a.shape), a.dtype
            src = a.reshape(-1).view(np.uint8)
        dtype = np.dtype(dtype)
        n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        with self.lock, torch.cuda.device(device):
            self._ensure()
            t = torch.empty(tuple(shape), dtype=_NP_TORCH[dtype.type], device=device)
            dst = t.reshape(-1).view(torch.uint8)
            stream = torch.cuda.current_stream(device)
            C_ = self.CHUNK
            offs = list(range(0, n, C_))

            def fill(slot, off, m, ev):
                if ev is not None:
                    ev.synchronize()                             # the slot's previous chunk has left
                if filler is not None:
                    filler(self.ring_np[slot * C_:slot * C_ + m], off, m)
                else:
                    self.ring_np[slot * C_:slot * C_ + m] = src[off:off + m]

            fills = [None] * self.SLOTS
            for i in range(min(self.SLOTS, len(offs))):
                fills[i] = self.pool.submit(fill, i, offs[i], min(C_, n - offs[i]), None)
            for i, off in enumerate(offs):
                slot = i % self.SLOTS
                m = min(C_, n - off)
                fills[slot].result()
                dst[off:off + m].copy_(self.ring[slot * C_:slot * C_ + m], non_blocking=True)
                self.events[slot].record(stream)
                j = i + self.SLOTS
                if j < len(offs):
                    fills[slot] = self.pool.submit(fill, slot, offs[j], min(C_, n - offs[j]), self.events[slot])
            stream.synchronize()                                 # the ring is free for the next call
            return t


_TORCH_NP = {torch.uint8: np.uint8, torch.int8: np.int8, torch.int16: np.int16, torch.uint16: np.uint16,
             torch.int32: np.int32, torch.int64: np.int64, torch.float32: np.float32, torch.float64: np.float64,
             torch.bool: np.bool_}
_NP_TORCH = {v: k for k, v in _TORCH_NP.items()}
_stager = _HostStager()
PINNED_UPLOAD_LIMIT = 256 << 20      # set_matrix_rows: smaller arrays are gathered on the host in one piece
RING_UPLOAD_MIN = 1 << 20            # uploads from this size on go through the pinned staging ring


def write_all(fh, mv):
    """``fh.write`` until every byte is out: an unbuffered file object may accept fewer bytes than it is given."""
    while len(mv):
        n = fh.write(mv)
        if n is None or n >= len(mv):
            return
        if n <= 0:
            raise OSError('write returned no progress')
        mv = mv[n:]


def drain_to_file(t, fh):
    """Write the bytes of a device tensor to an open binary file (chunked through the pinned ring)."""
    return _stager.drain_to_file(t, fh)


def to_host(t):
    """Device tensor -> numpy array (large results through the pinned ring of ``_HostStager``)."""
    if t.dtype not in _TORCH_NP:
        return t.cpu().numpy()
    return _stager.fetch(t)


def _np(x, dtype):
    return np.ascontiguousarray(np.asarray(x), dtype=dtype)


class Engine:
    """One handle on one GPU."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError('stellarscope_b200 needs a CUDA device (there is no CPU fallback)')
        self.lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.init()
            torch.zeros(1, device=self.device)          # make sure the primary context exists
            rc = self.lib.ss_create(self.device.index, C.byref(h))
        if rc != 0:
            raise EngineError(rc, 'ss_create failed')
        self.h = h
        self._keep = {}
        self.N = self.K = self.S = self.A = 0
        self.max_iter = 0

    # -- plumbing ------------------------------------------------------------
    _pool = {}          # device index -> idle engines (handle, internal streams and events are reused)

    @classmethod
    def acquire(cls, device=None):
        """An idle pooled engine of the device, or a new one.  Creating a handle costs ~2 ms (streams,
        events); a fit of 10^4 cells takes as long."""
        if not torch.cuda.is_available():
            raise RuntimeError('stellarscope_b200 needs a CUDA device (there is no CPU fallback)')
        idx = torch.cuda.current_device() if device is None else torch.device('cuda', device).index
        idle = cls._pool.get(idx)
        if idle:
            return idle.pop()
        return cls(idx)

    def release(self):
        """Return the engine to the pool (its device buffers are dropped, the handle stays)."""
        if getattr(self, 'h', None) is None:
            return
        if getattr(self, 'sharded', False):
            self.close()                     # sharded handles carry peer buffers / hooks: do not reuse
            return
        self.lib.ss_release(self.h)
        self._keep = {}
        self._params_cache = None
        Engine._pool.setdefault(self.device.index, []).append(self)

    def close(self):
        if getattr(self, 'h', None):
            self.lib.ss_destroy(self.h)
            self.h = None
            self._keep = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.ss_last_error(self.h).decode()
            cls = {_lib.SS_ERR_MODEL: ModelError, _lib.SS_ERR_NOT_CANONICAL: NotCanonicalError}.get(rc, EngineError)
            raise cls(rc, msg)

    def _to_dev(self, arr, dtype, pinned=True):
        a = _np(arr, dtype)
        if a.nbytes >= RING_UPLOAD_MIN and a.dtype.type in _NP_TORCH:
            # through the persistent pinned ring: pin_memory() allocates page-locked memory on every call
            # (~1 ms per array, more than the transfer of a few MB; for a multi-GB matrix more than the upload)
            return _stager.push(a, self.device)
        return torch.from_numpy(a).to(self.device)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def launch_count(self):
        return int(self.lib.ss_launch_count(self.h))

    def enable_timing(self, on=True):
        self._check(self.lib.ss_set_timing(self.h, int(bool(on))))

    def last_timings(self):
        """Device time (ms) of the kernel groups of the last ss_em_fit (synchronises):
        prepare, streaming kernel, resident kernels, emit, finalize."""
        thr = (C.c_int32 * 32)()
        nc = self.lib.ss_lane_classes(thr, 32)
        n = 5 + nc + 15
        buf = (C.c_double * n)()
        self._check(self.lib.ss_get_timings(self.h, buf, n))
        out = {'prepare': buf[0], 'em': buf[1], 'emit': buf[3], 'finalize': buf[4], 'stream': buf[5 + nc]}
        for i in range(nc):
            out[f'lane{i}_{thr[i]}'] = buf[5 + i]
        out['resident'] = max(buf[5:5 + nc])
        out['cluster'] = max(buf[5 + nc + 1:5 + nc + 15])
        out['cluster_heavy'] = max(buf[5 + nc + 1:5 + nc + 8])
        out['cluster_light'] = max(buf[5 + nc + 8:5 + nc + 15])
        return {k: float(v) for k, v in out.items()}

    # -- one model, rows sharded over ranks (pseudobulk; csrc/ss_dense.cu) ------------
    def set_sharding(self, rank, world, group=None, peers=True, K=None):
        """This handle fits a block of rows of ONE model shared with ``world - 1`` other ranks.

        Installs the all-reduce hook (``torch.distributed`` / NCCL on the caller's process group) for
        the build-time sums and -- with ``peers`` -- exchange buffers in NVLink-mapped symmetric
        memory, so that the per-iteration exchange happens inside the persistent fit kernel.
        ``world == 1`` selects the same dense code path on a single GPU.  Call before
        ``build_layout``."""
        rank, world = int(rank), int(world)
        self.sharded = True
        self.shard_rank, self.shard_world, self.shard_group = rank, world, group
        self.peer_mode = False
        ptr_arr = None
        nbytes = 0
        if world > 1:
            import torch.distributed as dist
            dev = self.device

            class _DevBuf:                                  # zero-copy view of the library's device buffer
                def __init__(self, ptr, n):
                    self.__cuda_array_interface__ = {'shape': (n,), 'typestr': '<f8', 'data': (ptr, False),
                                                     'version': 3, 'strides': None}

            def hook(_user, stream, buf, count, op):
                try:
                    t = torch.as_tensor(_DevBuf(C.cast(buf, C.c_void_p).value, int(count)), device=dev)
                    sptr = int(stream or 0)
                    cur = torch.cuda.current_stream(dev)
                    ctx = (torch.cuda.stream(torch.cuda.ExternalStream(sptr, device=dev))
                           if sptr != cur.cuda_stream else _null_ctx())
                    with torch.cuda.device(dev), ctx:
                        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM, group=group)
                    self.n_allreduce += 1
                    return 0
                except Exception:                           # never let an exception cross the C boundary
                    import traceback
                    traceback.print_exc()
                    return 1
            self.n_allreduce = 0
            self._hook = _lib.ALLREDUCE_FN(hook)
            self._check(self.lib.ss_set_allreduce(self.h, self._hook, None))
            if peers:
                if K is None:
                    raise ValueError('set_sharding(peers=True) needs the number of loci K')
                try:
                    import torch.distributed._symmetric_memory as symm
                    nbytes = int(self.lib.ss_peer_buffer_bytes(int(K)))
                    with torch.cuda.device(dev):
                        buf = symm.empty(nbytes // 8, dtype=torch.float64, device=dev)
                        buf.zero_()
                        g = group if group is not None else dist.group.WORLD
                        hdl = symm.rendezvous(buf, g)
                        ptrs = [int(p) for p in hdl.buffer_ptrs]
                        torch.cuda.synchronize(dev)
                        dist.barrier(group=group)           # every rank's buffer is zero before anybody signals
                    self._keep['symm_buf'], self._keep['symm_hdl'] = buf, hdl
                    ptr_arr = (C.c_void_p * world)(*ptrs)
                    self.peer_mode = True
                except Exception as e:                      # no peer access: per-iteration NCCL from the host loop
                    import logging
                    logging.getLogger(__name__).warning('symmetric memory unavailable (%s): using the NCCL hook '
                                                        'for the per-iteration exchange', e)
                    ptr_arr, nbytes = None, 0
        if world == 1 and not peers:
            nbytes = -1                                     # single rank, host-loop variant (tests)
        self._check(self.lib.ss_set_peers(self.h, rank, world, ptr_arr, nbytes))

    def sync_ranks(self):
        """Keep the ranks of a sharded fit in step (the peers kernel spins on its peers)."""
        if getattr(self, 'sharded', False) and self.shard_world > 1:
            import torch.distributed as dist
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.shard_group)

    def dense_params_sharded(self):
        """pi, theta over all K loci of the row-sharded model (synchronises)."""
        dev = self.device
        with torch.cuda.device(dev):
            pi = torch.empty(self.K, dtype=torch.float64, device=dev)
            th = torch.empty(self.K, dtype=torch.float64, device=dev)
            self._check(self.lib.ss_em_get_dense_params(self.h, self._stream(), self._p(pi), self._p(th)))
            return pi.cpu().numpy(), th.cpu().numpy()

    # -- layout --------------------------------------------------------------
    def set_matrix(self, indptr, indices, scores, K):
        """Upload the N x K uint16 CSR score matrix (host arrays) to the device."""
        self.N = len(indptr) - 1
        self.K = int(K)
        self.A = int(indptr[-1])
        if self.A > 2 ** 31 - 1 or self.N > 2 ** 31 - 2:
            # scipy switches to int64 index arrays at 2^31 entries; the device layout keeps int32 row pointers
            raise OverflowError(f'{self.A} alignments / {self.N} rows: the device layout uses int32 row pointers '
                                f'(at most 2^31 - 1 alignments); shard the rows over more GPUs')
        with torch.cuda.device(self.device):
            self._keep['row_ptr'] = self._to_dev(indptr, np.int32)
            self._keep['col_idx'] = self._to_dev(indices, np.int32)
            self._keep['score'] = self._to_dev(scores, np.uint16)
        self.h2d_bytes = 4 * (self.N + 1) + 6 * self.A

    def set_matrix_rows(self, indptr, indices, scores, rows, K):
        """Upload the given rows (ascending global ids) of a HOST CSR matrix as the engine's matrix: a rank of a
        sharded run holds only its own rows.  The rows are cut out while they are staged -- worker
        threads copy each chunk of the compacted arrays straight from the host matrix into the pinned ring
        (``gather_chunk``, csrc/ss_hostutil.c) -- so the host never holds a compacted copy."""
        from . import hostutil
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        indptr = np.ascontiguousarray(indptr)
        indices, scores = np.ascontiguousarray(indices), np.ascontiguousarray(scores)
        if indices.dtype != np.int32 or scores.dtype != np.uint16:
            return self.set_matrix(*hostutil.gather_rows(indptr, indices, scores, rows), K)     # converting route
        ptr = hostutil.row_extents(indptr, rows)
        self.N, self.K, self.A = len(rows), int(K), int(ptr[-1])
        if self.A > 2 ** 31 - 1:
            raise OverflowError(f'{self.A} alignments: the device layout uses int32 row pointers')
        lib = hostutil.load()
        with torch.cuda.device(self.device):
            self._keep['row_ptr'] = self._to_dev(ptr, np.int32)
            for name, arr in (('col_idx', indices), ('score', scores)):
                isz = arr.dtype.itemsize
                if self.A * isz <= PINNED_UPLOAD_LIMIT // 8:
                    out = np.empty(self.A, dtype=arr.dtype)                    # small: one gather, one copy
                    if self.A:
                        lib.gather_chunk(indptr, arr, rows, ptr, 0, self.A, out)
                    self._keep[name] = self._to_dev(out, arr.dtype.type)
                else:
                    def filler(dst, off, m, arr=arr, isz=isz):
                        lib.gather_chunk(indptr, arr, rows, ptr, off // isz, (off + m) // isz, dst)
                    self._keep[name] = _stager.push(None, self.device, filler=filler, shape=(self.A,), dtype=arr.dtype)
        self.h2d_bytes = 4 * (self.N + 1) + 6 * self.A

    def set_matrix_device(self, row_ptr, col_idx, score, K):
        """Use device tensors that are already resident (int32, int32, uint16)."""
        self.N = row_ptr.numel() - 1
        self.K = int(K)
        self.A = col_idx.numel()
        self._keep['row_ptr'], self._keep['col_idx'], self._keep['score'] = row_ptr, col_idx, score
        self.h2d_bytes = 0

    def build_layout(self, row_segment, n_segments):
        """ss_layout_build on the uploaded matrix; ``row_segment`` host or device int32[N]."""
        with torch.cuda.device(self.device):
            if isinstance(row_segment, torch.Tensor) and row_segment.is_cuda:
                rs = row_segment
            else:
                rs = self._to_dev(row_segment, np.int32)
                self.h2d_bytes += 4 * self.N
            self._keep['row_seg'] = rs
            k = self._keep
            rc = self.lib.ss_layout_build(self.h, self._stream(), self.N, self.K, self.A,
                                          self._p(k['row_ptr']), self._p(k['col_idx']), self._p(k['score']),
                                          self._p(rs), int(n_segments))
        self._check(rc)
        self.S = int(n_segments)

    def import_layout(self, hl):
        """Install a host-built layout (stellarscope_b200.layout_host.HostLayout). Test/debug."""
        c = _lib.HostLayoutC()
        keep = []

        def ptr(arr, dtype, ctype):
            a = _np(arr, dtype)
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(ctype))
        c.N, c.K, c.S, c.n_tiles, c.maxlen, c.chunk = hl.N, hl.K, hl.S, hl.n_tiles, hl.maxlen, hl.chunk
        c.n_uniq = len(hl.u_row)
        c.A, c.n_segcols = hl.A, len(hl.sc_gcol)
        c.tot_ent, c.tot_row, c.tot_col, c.tot_jds = len(hl.e_q), len(hl.r_w), len(hl.c_scol), len(hl.j_ptr)
        c.tile_hdr = ptr(hl.tile_hdr, np.int64, C.c_int64)
        for name, dt, ct in (
                ('e_lcol', np.uint16, C.c_uint16), ('e_cpos', np.uint16, C.c_uint16),
                ('e_q', np.float64, C.c_double), ('e_opos', np.int32, C.c_int32),
                ('r_len', np.uint16, C.c_uint16), ('r_w', np.float64, C.c_double),
                ('r_orow', np.int32, C.c_int32), ('c_ptr', np.uint16, C.c_uint16),
                ('c_scol', np.int32, C.c_int32), ('c_pisum0', np.float64, C.c_double),
                ('c_nuniq', np.int32, C.c_int32), ('j_ptr', np.uint16, C.c_uint16),
                ('seg_tile0', np.int32, C.c_int32), ('seg_ntiles', np.int32, C.c_int32),
                ('seg_col0', np.int32, C.c_int32), ('seg_ncol', np.int32, C.c_int32),
                ('seg_smax', np.int32, C.c_int32), ('seg_nobs', np.int32, C.c_int32),
                ('seg_namb', np.int32, C.c_int32), ('seg_nuniq', np.int32, C.c_int32),
                ('seg_wmax', np.float64, C.c_double), ('seg_wtot', np.float64, C.c_double),
                ('seg_wamb', np.float64, C.c_double), ('seg_logq_uniq', np.float64, C.c_double),
                ('seg_A', np.int64, C.c_int64), ('seg_Aamb', np.int64, C.c_int64),
                ('sc_gcol', np.int32, C.c_int32), ('sc_pisum0', np.float64, C.c_double),
                ('sc_nuniq', np.int32, C.c_int32), ('sc_inamb', np.int8, C.c_int8),
                ('u_row', np.int32, C.c_int32), ('u_pos', np.int32, C.c_int32)):
            setattr(c, name, ptr(getattr(hl, name), dt, ct))
        with torch.cuda.device(self.device):
            rc = self.lib.ss_layout_import(self.h, self._stream(), C.byref(c))
        self._check(rc)
        self.N, self.K, self.S, self.A = hl.N, hl.K, hl.S, hl.A

    def layout_info(self):
        info = _lib.LayoutInfo()
        self._check(self.lib.ss_layout_get_info(self.h, C.byref(info)))
        return info.as_dict()

    def export_layout_array(self, name, dtype):
        n = C.c_int64(0)
        self._check(self.lib.ss_layout_export(self.h, name.encode(), None, C.byref(n)))
        out = np.empty(n.value // np.dtype(dtype).itemsize, dtype=dtype)
        if n.value:
            self._check(self.lib.ss_layout_export(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), C.byref(n)))
        return out

    # -- fit -----------------------------------------------------------------
    def fit(self, em_epsilon=1e-7, max_iter=100, pi_prior=0.0, theta_prior=200000.0, use_likelihood=False):
        """ss_em_fit: enqueue EM for every segment (asynchronous)."""
        self.max_iter = int(max_iter)
        self.use_likelihood = bool(use_likelihood)
        self._params_cache = None
        self.sync_ranks()
        with torch.cuda.device(self.device):
            rc = self.lib.ss_em_fit(self.h, self._stream(), float(em_epsilon), int(max_iter), float(pi_prior),
                                    float(theta_prior), int(bool(use_likelihood)))
        self._check(rc)

    def results(self, traces=True):
        """Per-segment results as numpy arrays (synchronises).  The per-iteration traces (S x max_iter
        doubles -- 8 MB for 10^4 models) stay on the device: ``out['diffs']`` / ``out['lnls']`` are
        fetched on first access (``LazyTrace``)."""
        S, MI, dev = self.S, self.max_iter, self.device
        with torch.cuda.device(dev):
            i32 = torch.empty(4 * max(S, 1), dtype=torch.int32, device=dev)
            lnl = torch.empty(max(S, 1), dtype=torch.float64, device=dev)
            iters, flags, nobs, nfeats = (i32[k * S:(k + 1) * S] for k in range(4))
            diffs = torch.empty(S * MI, dtype=torch.float64, device=dev) if traces else None
            lnls = torch.empty(S * MI, dtype=torch.float64, device=dev) if traces and self.use_likelihood else None
            rc = self.lib.ss_em_get_results(self.h, self._stream(), self._p(iters), self._p(flags), self._p(lnl),
                                            self._p(diffs), self._p(lnls), self._p(nobs), self._p(nfeats))
            self._check(rc)
            h32 = i32.cpu().numpy()
            out = dict(iters=h32[0:S], flags=h32[S:2 * S], nobs=h32[2 * S:3 * S], nfeats=h32[3 * S:4 * S],
                       lnl=lnl[:S].cpu().numpy())
            if (out['flags'] & 8).any():
                raise EngineError(_lib.SS_ERR_CUDA, 'row-sharded fit: a peer rank did not answer (timeout)')
            if diffs is not None:
                out['diffs'] = LazyTrace(diffs, S, MI)
            if lnls is not None:
                out['lnls'] = LazyTrace(lnls, S, MI)
        return out

    def params_device(self):
        """Per segment-column parameters as DEVICE tensors (pi / theta / gcol over all segment columns) plus the
        small per-segment arrays on the host; callers that need a few models slice before copying."""
        S, dev = self.S, self.device
        nc = self.layout_info()['n_segcols']
        with torch.cuda.device(dev):
            col0 = torch.empty(max(S, 1), dtype=torch.int32, device=dev)
            ncol = torch.empty(max(S, 1), dtype=torch.int32, device=dev)
            gcol = torch.empty(max(nc, 1), dtype=torch.int32, device=dev)
            pi = torch.empty(max(nc, 1), dtype=torch.float64, device=dev)
            theta = torch.empty(max(nc, 1), dtype=torch.float64, device=dev)
            pi_rest = torch.empty(max(S, 1), dtype=torch.float64, device=dev)
            th_rest = torch.empty(max(S, 1), dtype=torch.float64, device=dev)
            rc = self.lib.ss_em_get_params(self.h, self._stream(), self._p(col0), self._p(ncol), self._p(gcol),
                                           self._p(pi), self._p(theta), self._p(pi_rest), self._p(th_rest))
            self._check(rc)
            return dict(seg_col0=col0[:S].cpu().numpy(), seg_ncol=ncol[:S].cpu().numpy(), gcol=gcol[:nc], pi=pi[:nc],
                        theta=theta[:nc], pi_rest=pi_rest[:S].cpu().numpy(), theta_rest=th_rest[:S].cpu().numpy())

    def params(self):
        """Per segment-column parameters (synchronises)."""
        p = self.params_device()
        for k in ('gcol', 'pi', 'theta'):
            p[k] = to_host(p[k])
        return p

    def dense_params(self, seg):
        """Full-length (K) pi and theta of one segment, like ``tl.pi`` / ``tl.theta``."""
        if getattr(self, 'sharded', False):
            return self.dense_params_sharded()
        p = self._params_cache = getattr(self, '_params_cache', None) or self.params()
        a, n = int(p['seg_col0'][seg]), int(p['seg_ncol'][seg])
        pi = np.full(self.K, p['pi_rest'][seg])
        th = np.full(self.K, p['theta_rest'][seg])
        pi[p['gcol'][a:a + n]] = p['pi'][a:a + n]
        th[p['gcol'][a:a + n]] = p['theta'][a:a + n]
        return pi, th

    def emit_z(self):
        """Device tensor with z in the CSR order of the input matrix."""
        with torch.cuda.device(self.device):
            z = torch.empty(max(self.A, 1), dtype=torch.float64, device=self.device)
            self._check(self.lib.ss_em_emit_z(self.h, self._stream(), self._p(z)))
        return z[:self.A]

    # -- UMI de-duplication (the step before the path) ---------------------------
    def umi_dedup(self, row_group, n_groups):
        """ss_umi_dedup on the uploaded matrix.  ``row_group[i]`` = id of the barcode+UMI pair of row i.
        Returns numpy arrays: excluded bool[N], comp_root int32[N], group_size / group_ncomp int32[n_groups]."""
        dev, k = self.device, self._keep
        n_groups = int(n_groups)
        with torch.cuda.device(dev):
            rg = row_group if isinstance(row_group, torch.Tensor) else self._to_dev(row_group, np.int32)
            ex = torch.empty(max(self.N, 1), dtype=torch.uint8, device=dev)
            i32 = torch.empty(max(self.N, 1) + 2 * max(n_groups, 1), dtype=torch.int32, device=dev)
            root, gs, gc = i32[:max(self.N, 1)], i32[max(self.N, 1):max(self.N, 1) + max(n_groups, 1)], \
                i32[max(self.N, 1) + max(n_groups, 1):]
            rc = self.lib.ss_umi_dedup(self.h, self._stream(), self.N, self._p(k['row_ptr']), self._p(k['col_idx']),
                                       self._p(k['score']), self._p(rg), n_groups, self._p(ex), self._p(root),
                                       self._p(gs), self._p(gc))
            self._check(rc)
            h = to_host(i32)
            n1 = max(self.N, 1)
            return dict(excluded=to_host(ex[:self.N]).astype(bool), comp_root=h[:self.N],
                        group_size=h[n1:n1 + n_groups], group_ncomp=h[n1 + max(n_groups, 1):n1 + max(n_groups, 1) + n_groups])

    def scores_build(self, read, feat, alnscore, alnlen, min_as, n_reads, n_feats, keep=True):
        """ss_scores_build: integer-coded mapping records (host int32 arrays) -> canonical uint16 CSR.
        Returns (indptr int32[N+1], indices int32[nnz], data uint16[nnz], info int64[4]) as numpy arrays; with
        ``keep`` the device copy becomes the engine's matrix (``set_matrix_device``), ready for ``build_layout``."""
        n = len(read)
        n_reads, n_feats = int(n_reads), int(n_feats)
        dev = self.device
        with torch.cuda.device(dev):
            rec = np.empty((4, max(n, 1)), dtype=np.int32)
            rec[0, :n], rec[1, :n], rec[2, :n], rec[3, :n] = read, feat, alnscore, alnlen
            d = self._to_dev(rec, np.int32)
            self.h2d_bytes = rec.nbytes
            row_ptr = torch.empty(n_reads + 1, dtype=torch.int32, device=dev)
            col_idx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            score = torch.empty(max(n, 1), dtype=torch.uint16, device=dev)
            info = np.zeros(4, dtype=np.int64)
            rc = self.lib.ss_scores_build(self.h, self._stream(), n, self._p(d[0]), self._p(d[1]), self._p(d[2]),
                                          self._p(d[3]), int(min_as), n_reads, n_feats, self._p(row_ptr),
                                          self._p(col_idx), self._p(score), info.ctypes.data_as(_lib.c_i64p))
            if rc == _lib.SS_ERR_CAPACITY:
                raise OverflowError(self.lib.ss_last_error(self.h).decode())
            self._check(rc)
            nnz = int(info[0])
            col_idx, score = col_idx[:nnz], score[:nnz]
            if keep:
                self.set_matrix_device(row_ptr, col_idx, score, n_feats)
            return to_host(row_ptr), to_host(col_idx), to_host(score), info

    # -- reassign / count ------------------------------------------------------
    def reassign(self, mode, z=None, conf_thresh=0.9, tie_tol=0.0):
        """ss_reassign.  Returns (flag uint8[A] device, nbest int32[N] device, info numpy int64[68])."""
        if mode not in MODES:
            raise ValueError(f'Argument "method" should be one of {list(MODES)}')
        dev = self.device
        k = self._keep
        with torch.cuda.device(dev):
            flag = torch.empty(max(self.A, 1), dtype=torch.uint8, device=dev)
            nbest = torch.zeros(max(self.N, 1), dtype=torch.int32, device=dev)
            info = torch.zeros(68, dtype=torch.int64, device=dev)
            rc = self.lib.ss_reassign(self.h, self._stream(), MODES[mode],
                                      float(conf_thresh if conf_thresh is not None else 0.0), float(tie_tol),
                                      self.N, self._p(k['row_ptr']), self._p(k['score']), self._p(z),
                                      self._p(flag), self._p(nbest), self._p(info))
            self._check(rc)
            return flag[:self.A], nbest[:self.N], info.cpu().numpy()

    def tied_rows(self, nbest):
        """Rows with more than one tied-best alignment (ascending) and their tie-set sizes, as host arrays -- all
        the host needs to draw the picks of the random modes."""
        with torch.cuda.device(self.device):
            rows = torch.nonzero(nbest > 1).reshape(-1).to(torch.int32)
            return to_host(rows), to_host(nbest[rows.long()])

    def reassign_pick(self, flag, tied_rows, picks):
        """ss_reassign_pick: keep the drawn alignment in every tied row (``flag`` is updated in place)."""
        n = len(tied_rows)
        if n == 0:
            return
        with torch.cuda.device(self.device):
            tr = self._to_dev(tied_rows, np.int32)
            pk = self._to_dev(picks, np.int32)
            self._check(self.lib.ss_reassign_pick(self.h, self._stream(), n, self._p(tr), self._p(pk),
                                                  self._p(self._keep['row_ptr']), self._p(flag)))
            torch.cuda.current_stream(self.device).synchronize()      # tr / pk may go

    def reassign_csr(self, flag, nbest=None, fractions=False):
        """ss_reassign_csr: the flagged alignments as CSR.  Returns numpy (indptr int32 / int64 [N+1], indices int32[n],
        values float64[n] or None); ``fractions``: 1 / nbest per selected alignment (best_average)."""
        dev, k = self.device, self._keep
        with torch.cuda.device(dev):
            out_ptr = torch.empty(self.N + 1, dtype=torch.int64, device=dev)
            out_idx = torch.empty(max(self.A, 1), dtype=torch.int32, device=dev)
            out_val = torch.empty(max(self.A, 1), dtype=torch.float64, device=dev) if fractions else None
            n = C.c_int64(0)
            rc = self.lib.ss_reassign_csr(self.h, self._stream(), self.N, self._p(k['row_ptr']), self._p(k['col_idx']),
                                          self._p(flag), self._p(nbest if fractions else None), max(self.A, 1),
                                          self._p(out_ptr), self._p(out_idx), self._p(out_val), C.byref(n))
            self._check(rc)
            m = n.value
            if m < 2 ** 31:
                out_ptr = out_ptr.to(torch.int32)         # half the bytes to bring home; scipy's index dtype anyway
            return to_host(out_ptr), to_host(out_idx[:m]), (to_host(out_val[:m]) if fractions else None)

    def mtx_text(self, row, col, val, integer):
        """ss_mtx_measure + ss_mtx_write: MatrixMarket body ("row+1 col+1 value\\n" per entry, order kept) of the COO
        triples as a uint8 device tensor.  ``row`` / ``col`` / ``val``: host arrays or device tensors."""
        dev = self.device
        with torch.cuda.device(dev):
            def dev_arr(a, np_dtype, t_dtype):
                if isinstance(a, torch.Tensor):
                    return a.to(device=dev, dtype=t_dtype).contiguous()
                return self._to_dev(a, np_dtype)
            r = dev_arr(row, np.int32, torch.int32)
            c = dev_arr(col, np.int32, torch.int32)
            if isinstance(val, torch.Tensor) or np.asarray(val).dtype == np.float64:
                v = dev_arr(val, np.float64, torch.float64)
            else:
                v = self._to_dev(val, np.asarray(val).dtype.type).to(torch.float64)   # widen on the device
            n = int(r.numel())
            if int(c.numel()) != n or int(v.numel()) != n:
                raise ValueError('row, col and val must have the same length')
            off = torch.empty(n + 1, dtype=torch.int64, device=dev)
            nbytes = C.c_int64(0)
            rc = self.lib.ss_mtx_measure(self.h, self._stream(), n, self._p(r), self._p(c), self._p(v),
                                         1 if integer else 0, self._p(off), C.byref(nbytes))
            self._check(rc)
            out = torch.empty(max(int(nbytes.value), 1), dtype=torch.uint8, device=dev)
            rc = self.lib.ss_mtx_write(self.h, self._stream(), n, self._p(r), self._p(c), self._p(v),
                                       1 if integer else 0, self._p(off), self._p(out))
            self._check(rc)
            torch.cuda.current_stream(dev).synchronize()          # r / c / v / off may go
            return out[:int(nbytes.value)]

    def count_cells(self, flag, row_cell, n_cells, keep_device=False, nbest=None):
        """ss_count_cells: integer counts per (cell, locus) without a sort, as the CSC form of the (loci x cells)
        matrix: numpy ``(indptr int64[n_cells+1], loci int32[nnz], counts int32[nnz])``; with ``keep_device`` a
        fourth element, the three device tensors (for the MTX writer).  Returns None when the loci do not fit the
        kernel's shared-memory bins (the caller takes ``count``).  With ``nbest`` (device, from ``reassign``) the
        counts are fractional -- 1 / nbest of the row per selected alignment (best_average) -- and float64."""
        dev, k = self.device, self._keep
        n_cells = int(n_cells)
        with torch.cuda.device(dev):
            rc_t = row_cell if isinstance(row_cell, torch.Tensor) else self._to_dev(row_cell, np.int32)
            out_ptr = torch.empty(n_cells + 1, dtype=torch.int64, device=dev)
            # upper bound of the result: selected alignments (flag bytes are 0 / 1)
            cap = int(flag.sum(dtype=torch.int64)) if self.A else 0
            cap = max(1, min(cap, n_cells * self.K))
            oc = torch.empty(cap, dtype=torch.int32, device=dev)
            ov = torch.empty(cap, dtype=torch.float64 if nbest is not None else torch.int32, device=dev)
            n = C.c_int64(0)
            rc = self.lib.ss_count_cells(self.h, self._stream(), self.N, self.K, n_cells, self._p(k['row_ptr']),
                                         self._p(k['col_idx']), self._p(flag), self._p(nbest), self._p(rc_t), cap,
                                         self._p(out_ptr), self._p(oc), self._p(None if nbest is not None else ov),
                                         self._p(ov if nbest is not None else None), C.byref(n))
            if rc == _lib.SS_ERR_CAPACITY:
                return None
            self._check(rc)
            m = n.value
            host = (to_host(out_ptr), to_host(oc[:m]), to_host(ov[:m]))
            if not keep_device:
                return host
            # compact device copies for the MTX writer (the cap-sized buffers go); cudaMemGetInfo costs ~20 ms a call,
            # so the size is checked against a fixed bound instead (2^28 entries = 2 GB)
            kept = (out_ptr, oc[:m].clone(), ov[:m].clone()) if m <= (1 << 28) else None
            return host + (kept,)

    def count(self, flag, row_cell, weight=None, row_umi=None, cap=None, keep_device=False):
        """ss_count.  Returns COO (cell, col, val) numpy arrays sorted by (cell, col); with ``keep_device`` a
        fourth element: compact device copies of the three arrays (for the MTX writer), or None when they
        would take more than an eighth of the free device memory."""
        dev = self.device
        k = self._keep
        if self.A == 0:
            empty = (np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.float64))
            return empty + (None,) if keep_device else empty
        with torch.cuda.device(dev):
            rc_t = row_cell if isinstance(row_cell, torch.Tensor) else self._to_dev(row_cell, np.int32)
            ru_t = None
            if row_umi is not None:
                ru_t = row_umi if isinstance(row_umi, torch.Tensor) else self._to_dev(row_umi, np.int32)
            cap = int(cap if cap is not None else max(self.A, 1))
            oc = torch.empty(cap, dtype=torch.int32, device=dev)
            ol = torch.empty(cap, dtype=torch.int32, device=dev)
            ov = torch.empty(cap, dtype=torch.float64, device=dev)
            n = C.c_int64(0)
            rc = self.lib.ss_count(self.h, self._stream(), self.N, self.K, self._p(k['row_ptr']),
                                   self._p(k['col_idx']), self._p(flag), self._p(weight), self._p(rc_t),
                                   self._p(ru_t), cap, self._p(oc), self._p(ol), self._p(ov), C.byref(n))
            self._check(rc)
            m = n.value
            host = (to_host(oc[:m]), to_host(ol[:m]), to_host(ov[:m]))
            if not keep_device:
                return host
            kept = (oc[:m].clone(), ol[:m].clone(), ov[:m].clone()) if m <= (1 << 28) else None   # the cap-sized buffers go
            return host + (kept,)
