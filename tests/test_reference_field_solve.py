"""Row f1 of SURVEY 8f -- the field half of the ECSIM step -- against the REFERENCE'S OWN CODE run here (oracle/_ref/libref_pic.so,
see tests/test_reference_ecsim.py): ECSIM::TimeStep = UpdateRhs, UpdateMatrixElement, GMRES, ProcessFinalSolution, UpdateB,
UpdateE on the fast-wave box with the J / mass matrix its own UpdateJMassMatrix left, a non-trivial E^n and B^n.

  * the constant operator tables == the reference's LaplacianStencil / GradDivStencil                      (exact)
  * right-hand side and operator (one product with a random vector), row by row                           (<= 1e-13 of the maximum)
  * E^{n+theta}, E^{n+1}, B^{n+1} after a solve to 1e-12                                                   (<= 1e-9 of the maximum)
The reference's GMRES lives in the un-vendored SWMF share library; it is replaced by oracle/ref_pic/gmres_single.cpp (the same
published algorithm), so the solve itself is compared at the tolerance both sides iterate to, not bit for bit."""
import numpy as np
import pytest

from oracle import ecsim_field
from oracle.ref_pic import ref_pic

needs_ref = pytest.mark.skipif(not ref_pic.available(), reason="oracle/_ref/libref_pic.so not built (needs /root/reference at build time)")


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.fixture(scope="module")
def live():
    from tests import ref_ecsim_case as rc

    return rc.case()


def solver_of(c):
    r = c["refpic"]
    return ecsim_field.EcsimField(c["mesh"], (1.0, 1.0, 1.0), r.light_speed, r.dt, theta=c["ref"]["field"]["theta"])


@needs_ref
def test_operator_tables_are_the_references(live):
    f = live["ref"]["field"]
    for p in range(3):
        assert np.array_equal(ecsim_field.graddiv_table(p, p), f["laplacian"][p])
        for q in range(3):
            assert np.array_equal(ecsim_field.graddiv_table(p, q), f["graddiv"][p][q])


@needs_ref
def test_rhs_and_operator_match_the_reference_row_by_row(live):
    f, ref = live["ref"]["field"], live["ref"]
    s = solver_of(live)
    J, M = ref["J1"][0], ref["M1"][0]
    assert rel(s.rhs(f["E"], f["B"], J, M), f["rhs"]) <= 1e-13
    assert rel(s.matvec(f["matvec_in"], M), f["matvec_out"]) <= 1e-13


@needs_ref
def test_field_step_matches_the_reference(live):
    f, ref = live["ref"]["field"], live["ref"]
    assert f["rel_residual"] <= 1e-12 and f["iterations"] > 10
    Eh, En, Bn, its = solver_of(live).step(f["E"], f["B"], ref["J1"][0], ref["M1"][0], tol=1e-12, max_iter=400)
    assert rel(Eh, f["E_half"]) <= 1e-9 and rel(En, f["E_new"]) <= 1e-9 and rel(Bn, f["B_new"]) <= 1e-9
    # given the same E^{n+theta}, UpdateB and UpdateE are node-local arithmetic: to rounding
    s = solver_of(live)
    assert rel(s.update_B(f["B"], f["E_half"]), f["B_new"]) <= 1e-14
    assert rel(s.update_E(f["E"], f["E_half"]), f["E_new"]) <= 1e-14
