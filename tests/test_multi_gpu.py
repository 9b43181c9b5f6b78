"""The multi-GPU face of the C ABI (fgq_init ... fgq_allgather, fgq_gausslegendre_host_multi) through ctypes, on
however many devices the box has (the logic is identical for 1 and 8; the 1 -> 8 numbers are in profiles/).
Everything is compared bit for bit with the single-device calls — the shards are the same kernels on the same
half-index ranges."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def multi(fgq):
    m = fgq.multi
    nd = m.init()
    assert nd == fgq.lib().fgq_device_count() and m.num_devices() == nd
    yield m
    m.synchronize()


@pytest.mark.parametrize("n", [0, 1, 5, 40, 60, 61, 1013, 100000, 100001, 26383, 1000003, 1 << 25])
def test_sharded_rule_equals_the_single_call(fgq, multi, n):
    rule = multi.gausslegendre_sharded(n)
    x1, w1 = fgq.gausslegendre(n)
    x, w = rule.to_host()
    assert np.array_equal(x, x1) and np.array_equal(w, w1)
    if n:
        assert abs(rule.sums[0] - 2.0) < 1e-12
    if n >= 2:   # a one-point rule integrates x^2 to 0
        assert abs(rule.sums[1] - 2.0 / 3.0) < 1e-12
    # shards tile [0, n) exactly once, two per device for the asymptotic path
    segs = rule.segments()
    assert len(segs) == (0 if n == 0 else 1 if n <= 60 else 2 * multi.num_devices())
    cover = np.zeros(n, dtype=int)
    for _, lo, hi, _, _ in segs:
        cover[lo:hi] += 1
    assert np.all(cover == 1)
    # steady state: the same descriptors are refilled without allocation, asynchronously
    ptrs = [(s.x, s.w) for s in rule.shards[:rule.count]]
    multi.gausslegendre_sharded(n, out=rule, sums=False)
    multi.synchronize()
    assert ptrs == [(s.x, s.w) for s in rule.shards[:rule.count]]
    x2, w2 = rule.to_host()
    assert np.array_equal(x2, x1) and np.array_equal(w2, w1)
    rule.free()


def test_domain_error_and_bad_descriptors(fgq, multi):
    with pytest.raises(fgq.DomainError):
        multi.gausslegendre_sharded(-1)
    rule = multi.gausslegendre_sharded(100000)
    rule.n = 100002  # descriptors of another rule must be refused, not overrun
    with pytest.raises(ValueError):
        rule.fill()
    rule.n = 100000
    rule.free()


@pytest.mark.parametrize("transport", ["push", "ce", "nccl"])
@pytest.mark.parametrize("n", [61, 100001, 3000000])
def test_allgather_replicates_the_rule(fgq, multi, n, transport, monkeypatch):
    monkeypatch.setenv("FGQ_ALLGATHER", transport)
    x1, w1 = fgq.gausslegendre(n)
    rule = multi.gausslegendre_sharded(n, sums=False)
    rep = multi.allgather(rule)
    multi.synchronize()
    for d in range(multi.num_devices()):
        x, w = rep.tensors(d)
        assert np.array_equal(x.cpu().numpy(), x1) and np.array_equal(w.cpu().numpy(), w1), (transport, d)
    rep.free()
    rule.free()


@pytest.mark.parametrize("n", [40, 100001, 3000000])
def test_replication_by_generation(fgq, multi, n):
    x1, w1 = fgq.gausslegendre(n)
    rep = multi.gausslegendre_replicated(n)
    multi.synchronize()
    for d in range(multi.num_devices()):
        x, w = rep.tensors(d)
        assert np.array_equal(x.cpu().numpy(), x1) and np.array_equal(w.cpu().numpy(), w1)
    rep.free()


@pytest.mark.parametrize("n", [1000, (1 << 23) + 1, 40_000_001])
def test_host_multi_equals_host(fgq, multi, n):
    x1, w1 = fgq.gausslegendre(n)
    x, w = multi.gausslegendre_host_multi(n)
    assert np.array_equal(x, x1) and np.array_equal(w, w1)


def test_timer(fgq, multi):
    rule = multi.gausslegendre_sharded(1 << 26, sums=False)
    multi.synchronize()
    multi.timer_start()
    for _ in range(5):
        multi.gausslegendre_sharded(1 << 26, out=rule, sums=False)
    ms, per = multi.timer_stop()
    assert 0 < ms < 1000 and len(per) == multi.num_devices() and max(per) == ms
    rule.free()


def test_explicit_device_list_and_reinit(fgq, multi):
    """fgq_init with an explicit list: idempotent for the same list, refused for a different one until fgq_finalize,
    out-of-range and duplicate ordinals are argument errors; after a re-init the sharded rule is the same rule."""
    nd = multi.num_devices()
    assert multi.init(list(range(nd))) == nd                      # same list: nothing to do
    with pytest.raises(ValueError):
        multi.init([nd + 3])
    with pytest.raises(ValueError):
        multi.init([0, 0])
    if nd > 1:
        with pytest.raises(ValueError):
            multi.init([0])                                       # different list while initialised
    multi.synchronize()
    multi.finalize()
    assert multi.num_devices() == 0
    assert multi.init([0]) == 1                                   # one device only
    rule = multi.gausslegendre_sharded(100001)
    x1, w1 = fgq.gausslegendre(100001)
    x, w = rule.to_host()
    assert np.array_equal(x, x1) and np.array_equal(w, w1) and len(rule.segments()) == 2
    rule.free()
    multi.finalize()
    assert multi.init() == nd                                     # back to all devices for the fixture's teardown
