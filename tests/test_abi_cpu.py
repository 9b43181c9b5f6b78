"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/fgq.h
declares, the Python mirror binds all of them, argument checks that need no GPU behave like the
reference, and there is no silent CPU fallback."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "fgq.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fgq_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def fgq_cpu():
    import fgq_b200
    if not os.path.exists(fgq_b200._lib.LIB_PATH):
        fgq_b200.build()
    return fgq_b200


def test_library_exports_every_declared_symbol(fgq_cpu):
    names = _declared()
    assert len(names) >= 10
    out = subprocess.run(["nm", "-D", "--defined-only", fgq_cpu._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (fgq_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in include/fgq.h but not exported: {missing}"


def test_python_mirror_binds_every_declared_symbol(fgq_cpu):
    L = fgq_cpu.lib()
    for n in _declared():
        assert n in fgq_cpu._lib.SIGNATURES, f"{n} has no ctypes signature"
        assert hasattr(L, n)
    assert L.fgq_version() >= 100


def test_only_sm100a_code_in_the_library(fgq_cpu):
    out = subprocess.run(["cuobjdump", "--list-elf", fgq_cpu._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_domain_errors_need_no_gpu(fgq_cpu):
    # gausslegendre.jl:26
    with pytest.raises(fgq_cpu.DomainError):
        fgq_cpu.gausslegendre(-1)
    L = fgq_cpu.lib()
    assert L.fgq_gausslegendre_host(-5, 0, 0) == 1
    assert b"non-negative" in L.fgq_last_error()
    x, w = fgq_cpu.gausslegendre(0)  # gausslegendre.jl:28
    assert len(x) == 0 and len(w) == 0
    with pytest.raises(TypeError):
        fgq_cpu.gausslegendre(2.5)


def test_no_cpu_fallback(fgq_cpu):
    """Without a device every compute entry point fails loudly."""
    if fgq_cpu.lib().fgq_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fgq_cpu.FGQError):
        fgq_cpu.gausslegendre(100)
    with pytest.raises(fgq_cpu.FGQError):
        fgq_cpu.gausslegendre_batch(np.array([5, 6, 7]))
    for call in (lambda: fgq_cpu.gaussjacobi(50, 0.3, -0.7), lambda: fgq_cpu.gausshermite(50),
                 lambda: fgq_cpu.gausslaguerre(50), lambda: fgq_cpu.gausschebyshevt(50),
                 lambda: fgq_cpu.approx_besselroots(0.3, 50), lambda: fgq_cpu.gausshermite_reduced(500),
                 lambda: fgq_cpu.gaussjacobi_batch([5, 6], 0.3, 0.2), lambda: fgq_cpu.gausshermite_batch([5, 6]),
                 lambda: fgq_cpu.gausslaguerre_batch([5, 6], 0.0)):
        with pytest.raises(fgq_cpu.FGQError):
            call()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fastgaussquadrature.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_partitions_cover_the_rule(fgq_cpu):
    for n in (61, 100, 101, 1000003, 10 ** 9, 10 ** 9 + 1):
        for world in (1, 2, 3, 4, 8):
            parts = fgq_cpu.index_partition(n, world)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert all(lo % 2 == 0 for lo, _ in parts)  # 16-byte aligned shard starts
            hp = fgq_cpu.half_index_partition(n, world)
            assert hp[0][0] == 1 and hp[-1][1] == (n + 1) // 2 + 1
            assert all(hp[i][1] == hp[i + 1][0] for i in range(world - 1))
            # the two segments of all ranks tile [0, n) exactly once
            cover = np.zeros(min(n, 2000), dtype=int) if n <= 2000 else None
            total = 0
            for k_lo, k_hi in hp:
                cnt = k_hi - k_lo
                m = (n + 1) // 2
                skip = 1 if (n % 2 and k_hi - 1 == m and cnt > 0) else 0
                total += cnt + (cnt - skip)
                if cover is not None:
                    cover[k_lo - 1:k_hi - 1] += 1
                    cover[n - k_hi + 1 + skip:n - k_lo + 1] += 1
            assert total == n
            if cover is not None:
                assert np.all(cover == 1)


def test_batch_descriptions_are_validated_before_any_device_work(fgq_cpu):
    """Host logic of the small-rule batches: argument errors map onto the reference's exceptions (DomainError) or
    ValueError without needing a device."""
    i64 = np.int64
    with pytest.raises(fgq_cpu.DomainError):
        fgq_cpu.gaussjacobi_batch([5, -1], 0.3, 0.2, offsets=np.array([0, 5, 5], dtype=i64))
    with pytest.raises(fgq_cpu.DomainError):
        fgq_cpu.gaussjacobi_batch([5, 6], [0.3, -1.0], 0.2)     # non-integrable weight (gaussjacobi.jl:44-45)
    with pytest.raises(ValueError):
        fgq_cpu.gaussjacobi_batch([5, 101], 0.3, 0.2)           # beyond the batched branch
    with pytest.raises(ValueError):
        fgq_cpu.gausshermite_batch([5, 201])
    with pytest.raises(fgq_cpu.DomainError):
        fgq_cpu.gausshermite_batch([5, -2], offsets=np.array([0, 5, 5], dtype=i64))
    with pytest.raises(ValueError):
        fgq_cpu.gausslaguerre_batch([5, 128], 0.0)
    with pytest.raises(fgq_cpu.DomainError):
        fgq_cpu.gausslaguerre_batch([5, 6], [-1.0, 0.0])        # gausslaguerre.jl:61
    for off, x, w in (fgq_cpu.gaussjacobi_batch([], 0.1, 0.2), fgq_cpu.gausshermite_batch([]),
                      fgq_cpu.gausslaguerre_batch([]), fgq_cpu.gausshermite_batch([0, 0])):
        assert len(x) == 0 and len(w) == 0
    # truncated Hermite rule: the a-priori size bound needs no device
    L = fgq_cpu.lib()
    assert L.fgq_gausshermite_reduced_bound(0) == 0
    assert L.fgq_gausshermite_reduced_bound(150) == 150          # small rules are never truncated
    b = L.fgq_gausshermite_reduced_bound(1_000_000)
    assert 24_000 < b < 25_000 and b % 2 == 0
    assert L.fgq_gausshermite_reduced_bound(1_000_001) % 2 == 1  # odd rules keep their centre node


def test_c_partition_equals_python_partition(fgq_cpu):
    """The half-index partition of the single-process multi-GPU calls (multi.cu) is the one the
    one-process-per-GPU mirror uses (api.half_index_partition): same shards either way."""
    import ctypes
    L = fgq_cpu.lib()
    for n in (61, 100, 101, 26383, 1000003, 10 ** 9, 10 ** 9 + 1, 2 ** 40 + 7):
        for world in (1, 2, 3, 4, 7, 8, 16):
            hp = fgq_cpu.half_index_partition(n, world)
            for r in range(world):
                lo, hi = ctypes.c_int64(0), ctypes.c_int64(0)
                assert L.fgq_debug_half_index_range(n, world, r, ctypes.byref(lo), ctypes.byref(hi)) == 0
                assert (lo.value, hi.value) == hp[r], (n, world, r)
    assert L.fgq_debug_half_index_range(100, 4, 4, None, None) == 3


def test_multi_gpu_calls_fail_loudly_without_a_device(fgq_cpu):
    if fgq_cpu.lib().fgq_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fgq_cpu.FGQError):
        fgq_cpu.multi.init()
    with pytest.raises(fgq_cpu.FGQError):
        fgq_cpu.multi.gausslegendre_sharded(1000)
    with pytest.raises(fgq_cpu.FGQError):
        fgq_cpu.multi.gausslegendre_host_multi(1000)
    with pytest.raises(fgq_cpu.DomainError):
        fgq_cpu.multi.gausslegendre_sharded(-1)
    assert fgq_cpu.multi.num_devices() == 0


def test_out_buffers_and_offsets_are_validated(fgq_cpu):
    """a strided `out` view or an inconsistent CSR description must be refused before the C side
    writes through the base pointer."""
    big = np.zeros(20)
    with pytest.raises(ValueError):
        fgq_cpu.gausslegendre(10, out=(big[::2], np.zeros(10)))
    ro = np.zeros(10)
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        fgq_cpu.gausslegendre(10, out=(np.zeros(10), ro))
    i64 = np.int64
    with pytest.raises(ValueError):
        fgq_cpu.gausslegendre_batch([5, 6], offsets=np.array([0, 5], dtype=i64))          # len(n_i) entries only
    with pytest.raises(ValueError):
        fgq_cpu.gausslegendre_batch([5, 6], offsets=np.array([0, 5, 8], dtype=i64))       # last rule overruns the total
    with pytest.raises(ValueError):
        fgq_cpu.gaussjacobi_batch([5, 6], 0.1, 0.2, offsets=np.array([0, 9, 5], dtype=i64))  # decreasing
    with pytest.raises(ValueError):
        fgq_cpu.gausshermite_batch([5, 6], offsets=np.array([3, 4, 11], dtype=i64))       # overlap


def test_user_integrand_expressions_are_checked_without_a_device(fgq_cpu):
    """The integrand string is validated before any compilation or device work; without a device a well-formed
    request fails loudly like every other compute entry point."""
    for bad in ("", "exp(x", "x; x", "x }", "x # y", 'printf("x")'):
        with pytest.raises(ValueError):
            fgq_cpu.gausslegendre_integrate(1000, bad)
    L = fgq_cpu.lib()
    assert L.fgq_rule_integrate_device(None, None, -1, b"x", None, None) == 3
    assert L.fgq_rule_integrate_device(None, None, 0, b"exp(x) * cos(3 * x)", (ctypes.c_double * 2)(), None) == 0
    if L.fgq_device_count() == 0:
        with pytest.raises(fgq_cpu.FGQError):
            fgq_cpu.gausslegendre_integrate(1000, "exp(x)")


def test_laguerre_batch_buckets_cover_every_rule_once(fgq_cpu):
    """Host planning of the bucketed Laguerre batch launches: terminates for any size list (empty rules included),
    every rule with n > 0 lands in exactly one bucket, a bucket spans at most 16 sizes."""
    L = fgq_cpu.lib()
    rng = np.random.default_rng(3)
    for n_i in (np.arange(0, 128), np.zeros(5, dtype=int), np.array([], dtype=int), rng.integers(0, 128, 10000),
                np.array([127, 127, 1, 0, 0, 16, 17]), np.array([0]), np.array([15])):
        n_i = np.ascontiguousarray(n_i, dtype=np.int32)
        first = np.zeros(16, dtype=np.int64)
        count = np.zeros(16, dtype=np.int64)
        top = np.zeros(16, dtype=np.int32)
        nb = L.fgq_debug_laguerre_batch_buckets(len(n_i), n_i.ctypes.data, first.ctypes.data, count.ctypes.data,
                                                top.ctypes.data, 16)
        srt = np.sort(n_i)[::-1]
        assert 0 <= nb <= 8
        assert count[:nb].sum() == np.count_nonzero(n_i > 0)
        pos = 0
        for b in range(nb):
            assert first[b] == pos and count[b] > 0 and top[b] == srt[pos]
            seg = srt[pos:pos + count[b]]
            assert seg.min() > (0 if top[b] <= 16 else ((top[b] - 1) // 16) * 16) and seg.max() == top[b]
            pos += count[b]


def test_ordered_planner_delivers_prefixes(fgq_cpu):
    """The background planner of the batch calls (OrderedPlanner): once wait(hi) returns every position below hi has
    been processed, and in the end every position exactly once — for sizes around its block size and wait steps that
    do not divide them."""
    L = fgq_cpu.lib()
    for n, step in ((1, 1), (1023, 100), (1024, 1024), (1025, 7), (100_000, 12_500), (250_001, 33_333)):
        assert L.fgq_debug_ordered_planner(n, step) == 0, (n, step)
    assert L.fgq_debug_ordered_planner(0, 1) == -1


def test_host_store_peak_needs_no_device(fgq_cpu):
    """The host-memory roofline of the host-returning calls is measured by host threads only."""
    L = fgq_cpu.lib()
    buf = np.zeros(1 << 20)
    g = ctypes.c_double(0.0)
    assert L.fgq_measure_host_store_peak(buf.ctypes.data, len(buf), 2, ctypes.byref(g)) == 0
    assert g.value > 0.1 and np.all(buf == buf[0]) and buf[0] >= 1.0
    assert L.fgq_measure_host_store_peak(None, len(buf), 2, ctypes.byref(g)) == 3
