"""Pin the oracle against the LIVE reference at the BASELINE.json config shapes.

Runs only where /root/reference exists (the build container).  The slowest case (SVHN,
batch 256) takes the reference ~20 s on 8 vCPUs.
"""
import numpy as np
import pytest

from . import cases as K
from . import refload, runners

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not refload.available(), reason="reference not present")]

# big-batch cases cost the reference minutes in fp64; fp32 only here.
FAST = [c for c in K.CONFIG_CASES if c.name not in ("C5_tiny_b512",)]


@pytest.mark.parametrize("c", FAST, ids=lambda c: c.name)
def test_oracle_f32_vs_live_reference_at_config_shape(c):
    params, io = K.make_params(c), K.make_io(c)
    ref = runners.run_reference(c, params=params, io=io)
    got = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
    errs = runners.compare(got, ref)
    assert set(errs) == set(ref)
    bad = {k: e for k, e in errs.items() if not e <= 5e-6}
    assert not bad, bad


@pytest.mark.parametrize("c", [K.GOLDEN_CASES[1], K.GOLDEN_CASES[6], K.GOLDEN_CASES[8], K.GOLDEN_CASES[10],
                               K.GOLDEN_CASES[13], K.GOLDEN_CASES[15]], ids=lambda c: c.name)
def test_oracle_f64_vs_reference_fp64_twin(c):
    params, io = K.make_params(c), K.make_io(c)
    ref = runners.run_reference(c, params=params, io=io, double=True)
    got = runners.run_oracle(c, params=params, io=io, dtype=np.float64)
    errs = runners.compare(got, ref)
    bad = {k: e for k, e in errs.items() if not e <= 1e-12}
    assert not bad, bad


# plane edges beyond the scripts' own (served by csrc/adi_generic.cu): the reference classes take any `size`,
# so the oracle is pinned there too -- odd edges, a 2 x 2 plane, 48 and 64 with every channel op
ODD_SIZES = [
    K.case("ref_size2_mnist", "mnist", B=3, size=2, num_steps=3, dt=0.05, dx=0.7, dy=1.3),
    K.case("ref_size7_svhn", "svhn", B=2, size=7, channels=3, num_steps=2),
    K.case("ref_size30_cifar10_c4", "cifar10", B=2, size=30, channels=4, dt=0.01, num_steps=2, dx=1.0, dy=1.5),
    K.case("ref_size36_cifar2_c2", "cifar2", B=3, size=36, channels=2, dt=0.02, num_steps=3),
    K.case("ref_size48_mnist", "mnist", B=2, size=48, num_steps=2, dt=0.05),
    K.case("ref_size64_svhn", "svhn", B=1, size=64, channels=3, num_steps=1),
]


@pytest.mark.parametrize("c", ODD_SIZES, ids=lambda c: c.name)
def test_oracle_vs_live_reference_at_other_plane_sizes(c):
    params, io = K.make_params(c), K.make_io(c)
    ref = runners.run_reference(c, params=params, io=io)
    got = runners.run_oracle(c, params=params, io=io, dtype=np.float32)
    errs = runners.compare(got, ref)
    assert set(errs) == set(ref)
    bad = {k: e for k, e in errs.items() if not e <= 5e-6}
    assert not bad, bad
    ref64 = runners.run_reference(c, params=params, io=io, double=True)
    got64 = runners.run_oracle(c, params=params, io=io, dtype=np.float64)
    bad = {k: e for k, e in runners.compare(got64, ref64).items() if not e <= 1e-12}
    assert not bad, bad
