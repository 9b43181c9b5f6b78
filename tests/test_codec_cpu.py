"""Transport coding of large host-returning calls (csrc/delta_codec.h): the reference encoder + the product's
decoder on the HOST must round-trip ANY 64-bit content bit for bit, and smooth rule data must compress to about
one byte per value.  (The device packer is checked on the GPU by the host-path parity tests, which compare whole
rules against the oracle.)"""
import ctypes

import numpy as np
import pytest


def _roundtrip(lib, a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.full_like(a, np.nan)
    nbytes = ctypes.c_int64(0)
    rc = lib.fgq_debug_codec_roundtrip_host(a.ctypes.data, len(a), out.ctypes.data, ctypes.byref(nbytes))
    assert rc == 0
    assert np.array_equal(a.view(np.int64), out.view(np.int64))   # bit patterns (NaN payloads, signed zeros too)
    return nbytes.value


@pytest.fixture(scope="module")
def lib():
    import fgq_b200
    return fgq_b200.lib()


@pytest.mark.parametrize("cnt", [0, 1, 2, 3, 1023, 1024, 1025, 5000, 1 << 16])
def test_roundtrip_arbitrary_content(lib, cnt):
    rng = np.random.default_rng(cnt)
    bits = rng.integers(-2 ** 63, 2 ** 63 - 1, cnt, dtype=np.int64)   # random patterns incl. NaN / inf / denormals
    nbytes = _roundtrip(lib, bits.view(np.float64))
    assert nbytes <= cnt * 8 + 64 * ((cnt + 1023) // 1024)           # never worse than raw + headers
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 1.0, -1.0] * 300)[:cnt]
    _roundtrip(lib, special)


@pytest.mark.parametrize("n,limit", [(4_000_000, 4.2), (40_000_000, 2.1)])
def test_rule_data_compresses(lib, oracle_mod, n, limit):
    """Left half of gausslegendre(n): the (stride-4) second difference of the bit patterns is 16 x the curvature
    of the rule in ulps, 16 (pi/n)^2 / 2^-53 = 88000 at n = 4e6 (32-bit class), 880 at n = 4e7 (16-bit class) and
    1.4 at n = 1e9 (8-bit class, 1 B/value), plus a few ulp of rounding noise; the first blocks of the weights
    and the blocks with a binade crossing take a wider class."""
    x, w = oracle_mod.gausslegendre(n)
    m = n // 2
    bx = _roundtrip(lib, x[:m])
    bw = _roundtrip(lib, w[:m])
    assert bx < limit * m and bw < limit * m, (bx / m, bw / m)
    if n <= 4_000_000:  # the whole rule at once (crosses the centre: sign change of x) still round-trips
        _roundtrip(lib, x)
        _roundtrip(lib, w)


def test_headline_size_compresses_to_one_byte(lib):
    """A window of the n = 1e9 rule (nodes -cos(theta_k), weights ~ (pi/n) sin(theta_k), correctly rounded here):
    one byte per value."""
    n = 10 ** 9
    k = np.arange(2 * 10 ** 8, 2 * 10 ** 8 + (1 << 20), dtype=np.float64)
    th = (k + 0.75) * np.pi / (n + 0.5)
    for v in (-np.cos(th), np.pi / (n + 0.5) * np.sin(th)):
        assert _roundtrip(lib, v) < 1.08 * len(v)


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("negate", [0, 1])
def test_unpack_into_slice_and_mirror(lib, fused, negate):
    """The routine the pipeline's host threads run (decode a block into the left slice and its mirror image), for
    every destination alignment, lengths that are not multiples of the block / vector size, a rule with and without
    a centre node, smooth (8-bit class) and random (64-bit class) content; with the bounce buffer and with the
    register-to-destination variant."""
    rng = np.random.default_rng(7 + fused + 2 * negate)
    for cnt in (1, 7, 8, 1023, 1024, 1028, 4096, 5 * 1024 + 3, 40_000):
        k = np.arange(3 * 10 ** 8, 3 * 10 ** 8 + cnt, dtype=np.float64)
        smooth = -np.cos((k + 0.75) * np.pi / (10 ** 9 + 0.5))
        noisy = rng.integers(-2 ** 62, 2 ** 62, cnt, dtype=np.int64).view(np.float64)
        for src in (smooth, noisy):
            for mirror_cnt in (cnt, cnt - 1):
                for off_l in range(4):
                    for off_r in (0, 1, 3):
                        left = np.full(cnt + 8, 7.25)
                        right = np.full(cnt + 8, 7.25)
                        lv, rv = left[off_l:off_l + cnt], right[off_r:off_r + cnt]
                        rc = lib.fgq_debug_codec_unpack_host(src.ctypes.data, cnt, lv.ctypes.data, rv.ctypes.data,
                                                             mirror_cnt, negate, fused)
                        assert rc == 0
                        assert np.array_equal(lv.view(np.int64), src.view(np.int64))
                        want = (-src if negate else src)[:mirror_cnt][::-1]
                        assert np.array_equal(rv[cnt - mirror_cnt:].view(np.int64), want.view(np.int64))
                        # nothing outside the destinations was touched (guard bands, and the centre slot)
                        assert np.all(left[:off_l] == 7.25) and np.all(left[off_l + cnt:] == 7.25)
                        assert np.all(right[:off_r + cnt - mirror_cnt] == 7.25) and np.all(right[off_r + cnt:] == 7.25)
