"""Device -> host transfer of large results (engine.to_host: pinned ring + threaded copy-out) is a plain copy."""
import numpy as np
import pytest


@pytest.mark.gpu
def test_to_host_equals_plain_copy_for_all_result_dtypes_and_sizes():
    import torch
    from stellarscope_b200 import engine
    g = torch.Generator(device='cuda').manual_seed(1)
    chunk = engine._HostStager.CHUNK
    for dtype, n in ((torch.uint8, 3 * chunk + 17), (torch.int32, chunk // 4 * 9 + 3), (torch.float64, chunk // 8 * 11 + 1),
                     (torch.uint16, chunk + 5), (torch.int32, 1000), (torch.float64, engine._HostStager.MIN_BYTES // 8)):
        if dtype.is_floating_point:
            t = torch.rand(n, generator=g, device='cuda', dtype=dtype)
        else:
            t = torch.randint(0, 200, (n,), generator=g, device='cuda', dtype=torch.int32).to(dtype)
        for rep in range(2):                                       # ring slots are reused across calls
            h = engine.to_host(t)
            ref = t.cpu().numpy()
            assert h.dtype == ref.dtype and h.shape == ref.shape and np.array_equal(h, ref)
    # a strided view and a 2-d result
    t2 = torch.arange(6 * (chunk // 8), device='cuda', dtype=torch.float64).reshape(6, -1)
    assert np.array_equal(engine.to_host(t2), t2.cpu().numpy())
    assert np.array_equal(engine.to_host(t2[:, ::2]), t2[:, ::2].cpu().numpy())
    assert engine.to_host(torch.zeros(0, device='cuda')).shape == (0,)


@pytest.mark.gpu
def test_push_through_the_ring_is_a_plain_upload(monkeypatch):
    import torch
    from stellarscope_b200 import engine
    rng = np.random.default_rng(2)
    chunk = engine._HostStager.CHUNK
    for dtype, n in ((np.int32, chunk // 4 * 10 + 7), (np.uint16, chunk // 2 * 3 + 1), (np.float64, chunk // 8 + 3), (np.uint8, 5)):
        a = rng.integers(0, 60000, n).astype(dtype)
        for rep in range(2):
            t = engine._stager.push(a, torch.device('cuda', 0))
            assert t.dtype == torch.from_numpy(a[:1]).dtype and np.array_equal(t.cpu().numpy(), a)
    # Engine._to_dev takes this route above the limit
    monkeypatch.setattr(engine, 'PINNED_UPLOAD_LIMIT', 1 << 20)
    eng = engine.Engine(0)
    a = rng.integers(0, 1000, 3_000_000).astype(np.int32)
    assert np.array_equal(eng._to_dev(a, np.int32).cpu().numpy(), a)
    eng.close()


@pytest.mark.gpu
def test_reassign_csr_is_the_compaction_of_the_flags():
    """ss_reassign_csr against numpy: row pointers, kept columns in row order, 1/nbest fractions; no selection at all."""
    import torch
    from stellarscope_b200.engine import Engine
    rng = np.random.default_rng(6)
    N, K = 20_000, 300
    lens = rng.integers(0, 7, N)                                  # empty rows included
    indptr = np.zeros(N + 1, dtype=np.int32)
    np.cumsum(lens, out=indptr[1:])
    A = int(indptr[-1])
    indices = np.concatenate([np.sort(rng.choice(K, size=l, replace=False)) for l in lens if l]).astype(np.int32)
    eng = Engine(0)
    eng.set_matrix(indptr, indices, rng.integers(1, 200, A).astype(np.uint16), K)
    flag = (rng.random(A) < 0.4).astype(np.uint8)
    nbest = rng.integers(0, 5, N).astype(np.int32)
    row = np.repeat(np.arange(N), lens)
    for fl in (flag, np.zeros(A, dtype=np.uint8), np.ones(A, dtype=np.uint8)):
        ptr, idx, val = eng.reassign_csr(torch.from_numpy(fl).cuda(), torch.from_numpy(nbest).cuda(), fractions=True)
        keep = fl.astype(bool)
        exp_ptr = np.zeros(N + 1, dtype=np.int64)
        np.cumsum(np.bincount(row[keep], minlength=N), out=exp_ptr[1:])
        assert np.array_equal(ptr, exp_ptr) and np.array_equal(idx, indices[keep])
        nb = nbest[row[keep]]
        assert np.array_equal(val, np.where(nb > 0, 1.0 / np.maximum(nb, 1), 1.0))
        ptr2, idx2, val2 = eng.reassign_csr(torch.from_numpy(fl).cuda())
        assert val2 is None and np.array_equal(ptr2, exp_ptr) and np.array_equal(idx2, indices[keep])
    eng.close()
