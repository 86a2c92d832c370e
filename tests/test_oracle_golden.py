"""The C oracle against the fixtures generated from the unmodified reference (CPU only)."""
import numpy as np
import pytest

from . import cases as K
from . import golden_io, runners

# fp32 tolerance for the oracle.  north_star's bar for the product is 1e-5; the oracle follows
# the reference's op order and lands ~5e-7 away, so it is held to a tighter 3e-6.
ORACLE_TOL = 3e-6


@pytest.mark.parametrize("c", K.GOLDEN_CASES, ids=lambda c: c.name)
def test_oracle_f32_matches_reference_fixture(c):
    params, io, ref = golden_io.load(c)
    # the fixture must have been generated from the case definition in tests/cases.py
    want_params = K.make_params(c)
    for k, v in want_params.items():
        np.testing.assert_array_equal(params[k], v, err_msg=f"fixture weights drifted: {k}")
    got = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
    errs = runners.compare(got, ref)
    assert set(errs) == set(ref), (sorted(errs), sorted(ref))
    bad = {k: e for k, e in errs.items() if not e <= ORACLE_TOL}
    assert not bad, bad


@pytest.mark.parametrize("c", K.GOLDEN_CASES, ids=lambda c: c.name)
def test_oracle_f64_is_closer_to_reference_than_tolerance(c):
    """fp64 oracle vs the fp32 reference fixture: bounded by the reference's own fp32 noise."""
    params, io, ref = golden_io.load(c)
    got = runners.run_oracle(c, params=params, io=io, dtype=np.float64)
    errs = runners.compare(got, ref)
    bad = {k: e for k, e in errs.items() if not e <= 1e-5}
    assert not bad, bad


def test_oracle_edge_cases():
    """Batch of one, zero steps (identity), no grad_input requested."""
    c = K.case("edge_b1", "mnist", B=1, size=8, num_steps=2, dt=0.05)
    full = runners.run_oracle(c, dtype=np.float64)
    nog = runners.run_oracle(c, dtype=np.float64, need_gin=False)
    assert nog["gin"] is None
    np.testing.assert_array_equal(full["g_alpha_base"], nog["g_alpha_base"])
    c0 = K.case("edge_steps0", "cifar10", B=2, size=8, channels=3, num_steps=0)
    u, g = K.make_io(c0)
    r = runners.run_oracle(c0, dtype=np.float32)
    np.testing.assert_array_equal(r["y"], u)
    np.testing.assert_array_equal(r["gin"], g)
    assert not np.any(r["g_alpha_base"])
    # empty batch
    ce = K.case("edge_empty", "tiny", B=0, size=8, channels=2)
    r = runners.run_oracle(ce, dtype=np.float32)
    assert r["y"].shape == (0, 2, 8, 8) and not np.any(r["g_alpha_base"])


def test_oracle_threads_do_not_change_forward():
    c = K.case("thr", "svhn", B=5, size=8, channels=3, num_steps=2)
    a = runners.run_oracle(c, dtype=np.float32, nthreads=1)
    b = runners.run_oracle(c, dtype=np.float32, nthreads=4)
    np.testing.assert_array_equal(a["y"], b["y"])
    np.testing.assert_array_equal(a["gin"], b["gin"])
    np.testing.assert_allclose(a["g_beta_base"], b["g_beta_base"], rtol=1e-12, atol=1e-300)
