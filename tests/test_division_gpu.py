"""The mover's shared-reciprocal quotients must equal IEEE fp64 division bit for bit (cell-assignment parity depends on it)."""
import numpy as np
import pytest

from tests import parity_util as pu
from amps_b200 import api

pytestmark = pytest.mark.gpu


def test_shared_reciprocal_division_is_ieee():
    m, cfg, parts, fields = pu.make_case(n_cells=(8, 8, 8), ppc=1, seed=1)
    g = api.Context(cfg, m)
    rng = np.random.default_rng(12345)
    n = 1 << 22
    total_bad = 0
    # (1) generic operands over many binades, both signs
    a = rng.standard_normal(n) * np.exp(rng.uniform(-40, 40, n))
    b = rng.standard_normal(n) * np.exp(rng.uniform(-40, 40, n))
    total_bad += g.selftest_division(a, b)
    # (2) stencil weights divided by norms within a few ulp of 1 (incl. the all-ones significand 1-2^-53)
    a = rng.random(n) * rng.random(n) * rng.random(n)
    k = rng.integers(-6, 10, n)
    b = np.where(k < 0, 1.0 + k * 2.0 ** -53, 1.0 + k * 2.0 ** -52)
    total_bad += g.selftest_division(a, b)
    # (3) block constants: cell sizes / spans / lattice steps, numerators = coordinate differences
    a = rng.uniform(-100.0, 100.0, n)
    b = rng.choice(np.array([1.0, 0.5, 0.25, 8.0, 80.0 / 40960.0, 1.0 / 3.0, 0.1, 16.0 / 7.0, 2.0 - 2.0 ** -52, 1.0 - 2.0 ** -53]), n)
    total_bad += g.selftest_division(a, b)
    # (4) tiny / huge / zero numerators
    a = np.concatenate([np.zeros(1024), 10.0 ** rng.uniform(-300, 300, n - 1024)])
    b = rng.uniform(0.5, 2.0, n)
    total_bad += g.selftest_division(a, b)
    g.close()
    assert total_bad == 0
