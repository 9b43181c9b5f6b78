"""Thread-safety of the boundary: the reference is re-entrant and has no mutable global
state, so concurrent calls from different Julia tasks are legal today and must stay legal.  ctypes
releases the GIL during a foreign call, so the threads below are truly concurrent inside
libfgq_cuda.so; every result must equal the one the same call gives when it runs alone."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _calls(fgq):
    return [
        ("legendre 100001", lambda: fgq.gausslegendre(100001)),
        ("legendre 42", lambda: fgq.gausslegendre(42)),
        ("legendre 3000001", lambda: fgq.gausslegendre(3000001)),
        ("jacobi 20000", lambda: fgq.gaussjacobi(20000, 0.3, -0.7)),
        ("jacobi 57", lambda: fgq.gaussjacobi(57, 1.5, 0.25)),
        ("radau 5000", lambda: fgq.gaussradau(5000)),
        ("lobatto 5000", lambda: fgq.gausslobatto(5000)),
        ("hermite 30000", lambda: fgq.gausshermite(30000)),
        ("hermite 100", lambda: fgq.gausshermite(100)),
        ("laguerre 20000", lambda: fgq.gausslaguerre(20000, 0.5)),
        ("laguerre 60", lambda: fgq.gausslaguerre(60)),
        ("chebyshev t 1000", lambda: fgq.gausschebyshevt(1000)),
    ]


def test_concurrent_host_calls_equal_serial_calls(fgq):
    calls = _calls(fgq)
    serial = [fn() for _, fn in calls]
    results = {}
    errors = []
    start = threading.Barrier(len(calls) * 2)

    def worker(slot, fn):
        try:
            start.wait()
            for rep in range(3):
                results[(slot, rep)] = fn()
        except Exception as e:  # noqa: BLE001
            errors.append((slot, repr(e)))

    threads = []
    for copy in range(2):
        for i, (_, fn) in enumerate(calls):
            threads.append(threading.Thread(target=worker, args=((i, copy), fn)))
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for ((i, _copy), _rep), (x, w) in results.items():
        xs, ws = serial[i]
        assert np.array_equal(x, xs) and np.array_equal(w, ws), f"{calls[i][0]} differs under concurrency"


def test_concurrent_device_calls_on_separate_streams(fgq):
    """Device entry points keep no global state: two rules generated at once on two streams, each into
    its own buffers and workspace, give the bits of the stand-alone calls."""
    import torch
    L = fgq.lib()
    chk = fgq._lib.check
    n1, n2 = 400_000, 300_001
    ref_j = fgq.gaussjacobi(n1, 0.3, -0.7)
    ref_h = fgq.gausshermite(n2)
    ref_l = fgq.gausslegendre(n1)
    s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    xj = torch.empty(n1, dtype=torch.float64, device="cuda")
    wj = torch.empty_like(xj)
    wsj = torch.empty(int(L.fgq_gaussjacobi_workspace_bytes(n1)) + 16, dtype=torch.uint8, device="cuda")
    xh = torch.empty(n2, dtype=torch.float64, device="cuda")
    wh = torch.empty_like(xh)
    wsh = torch.empty(int(L.fgq_gausshermite_workspace_bytes(n2)) + 16, dtype=torch.uint8, device="cuda")
    xl = torch.empty(n1, dtype=torch.float64, device="cuda")
    wl = torch.empty_like(xl)
    torch.cuda.synchronize()
    for _ in range(3):
        chk(L.fgq_gaussjacobi_device(n1, 0.3, -0.7, 0, n1, xj.data_ptr(), wj.data_ptr(), wsj.data_ptr(), s1.cuda_stream))
        chk(L.fgq_gausshermite_device(n2, 0, 0, n2, xh.data_ptr(), wh.data_ptr(), wsh.data_ptr(), s2.cuda_stream))
        chk(L.fgq_gausslegendre_device(n1, 0, n1, xl.data_ptr(), wl.data_ptr(), s3.cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(xj.cpu().numpy(), ref_j[0]) and np.array_equal(wj.cpu().numpy(), ref_j[1])
    assert np.array_equal(xh.cpu().numpy(), ref_h[0]) and np.array_equal(wh.cpu().numpy(), ref_h[1])
    assert np.array_equal(xl.cpu().numpy(), ref_l[0]) and np.array_equal(wl.cpu().numpy(), ref_l[1])
