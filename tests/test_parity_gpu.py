"""GPU parity tests proper: the CUDA path through the C ABI against the CPU oracle on the same seeded inputs."""
import pytest

from tests import parity_util as pu

pytestmark = pytest.mark.gpu

CASES = {
    "cube16_ppc8": dict(n_cells=(16, 16, 16), ppc=8, seed=3),
    "slab_32x16x8_blocks_16x8x4": dict(n_cells=(32, 16, 8), ppc=6, seed=5, block_cells=(16, 8, 4)),  # fast-wave geometry
    "cube24_fast": dict(n_cells=(24, 24, 24), ppc=4, seed=7, vscale=8.0),  # many cell/block crossers and periodic wraps
    "ghost2": dict(n_cells=(16, 16, 16), ppc=4, seed=9, ghost_cells=(2, 2, 2)),
    "single_block_dim": dict(n_cells=(8, 16, 8), ppc=8, seed=11),
    "dense_cells_320": dict(n_cells=(8, 8, 8), ppc=160, seed=17),  # cells larger than one deposit chunk / ring batch
    "open_box_user_function": dict(n_cells=(16, 16, 16), ppc=6, seed=15, periodic=False, vscale=6.0, boundary_mode=2),  # exit records
    "open_box": dict(n_cells=(16, 16, 16), ppc=6, seed=13, periodic=False, vscale=6.0),  # DELETE boundary
    # 16^3-cell blocks: the E + B tiles (300 KB) do not fit in shared memory, the movers read them through L2 (kSmemTiles = false)
    "big_blocks_16": dict(n_cells=(32, 32, 32), ppc=2, seed=27, block_cells=(16, 16, 16), vscale=4.0),
    "four_species_neutral": dict(n_cells=(16, 16, 16), ppc=8, seed=25, four_species=True, vscale=3.0),  # > 2 species: diag_kernel path
    # _PIC_FIELD_SOLVER_B_CORNER_BASED_: B on the corner nodes, same stencil as E (pic_mover_boris.cpp:952-965)
    "corner_B_periodic": dict(n_cells=(16, 16, 16), ppc=8, seed=19, b_mode=1),
    "corner_B_open": dict(n_cells=(16, 16, 16), ppc=6, seed=21, periodic=False, vscale=6.0, b_mode=1),
    # BASELINE config 4 scaled down: 3-level sphere-refined open box, mixed ppc, drifting Maxwellian (particles cross levels)
    "amr_3_levels_corner_B": dict(n_cells=(32, 32, 32), ppc=8, seed=23, amr_radii=(9.0, 4.0), ppc_by_level=(8, 4, 2), vscale=4.0, b_mode=1,
                                  E_amp=0.02),
}


@pytest.mark.parametrize("exact", [0, 1])
@pytest.mark.parametrize("name", list(CASES))
def test_one_step_parity(name, exact):
    res = pu.run_parity_case(exact_arithmetic=exact, **CASES[name])
    print(name, "exact" if exact else "fast", res)
    if exact:
        assert res["bit_mismatch_xv"] == 0, res       # every operation rounded like the CPU build
    else:
        # the fast kernel really did the work (open boxes: leavers and the truncated boundary stencils of centre-based B
        # go to the exact kernel, and every block of these small boxes touches the boundary)
        assert res["n_redo"] < (0.05 if CASES[name].get("periodic", True) and "amr_radii" not in CASES[name] else 0.5) * res["n"], res
    assert res["oracle_lists"] == 0
    assert res["cell_mismatch"] == 0, res           # bit-exact block/cell assignment
    assert res["stats_equal"], res                   # bit-exact crossing counts
    assert res["max_rel_x"] <= pu.REL_TOL and res["max_rel_v"] <= pu.REL_TOL, res
    assert res["sorted_ok"] and res["table_ok"] and res["perm_ok"], res
    assert res["max_rel_J"] <= pu.REL_TOL and res["max_rel_M"] <= pu.REL_TOL, res
    assert res["rel_energy"] <= pu.REL_TOL and res["rel_cfl"] <= pu.REL_TOL, res
    assert res["gpu_launches"] > 0
    assert res.get("records_equal", True), res


@pytest.mark.parametrize("name", ["cube16_ppc8", "open_box", "corner_B_periodic", "amr_3_levels_corner_B"])
def test_fused_step_equals_the_separate_phases(name):
    """amps_gpu_step (move + permutation-only sort + gathering deposit that writes the sorted copy) against
    move / sort / deposit called one by one: same particles per cell, same J, M, diagnostics."""
    import numpy as np
    from amps_b200 import api

    m, cfg, parts, fields = pu.make_case(**CASES[name])
    ref = pu.run_gpu(m, cfg, parts, fields)
    x, v, w, sp, cells = parts
    E, B, Bcur = fields
    g = api.Context(cfg, m)
    g.fields_upload(E, B, Bcur)
    g.particles_upload(x, v, w, sp, cells)
    g.step()
    srt = g.particles_download()
    table = g.cell_table()
    J, M = g.JM_download()
    en, cfl = g.diagnostics()
    g.close()
    a, b = ref["sorted"], srt
    assert len(a["ptrs"]) == len(b["ptrs"]) and (table == ref["table"]).all()
    assert (np.diff(b["cells"].astype(np.int64)) >= 0).all()
    oa, ob = np.argsort(a["ptrs"]), np.argsort(b["ptrs"])
    assert (a["ptrs"][oa] == b["ptrs"][ob]).all() and (a["cells"][oa] == b["cells"][ob]).all()
    assert (a["x"][:, oa] == b["x"][:, ob]).all() and (a["v"][:, oa] == b["v"][:, ob]).all()
    assert (a["w"][oa] == b["w"][ob]).all() and (a["species"][oa] == b["species"][ob]).all()
    assert pu.rel_scaled(J, ref["J"]) <= 1e-12 and pu.rel_scaled(M, ref["M"]) <= 1e-12
    assert abs(en - ref["energy"]) <= 1e-12 * abs(ref["energy"])
    assert max(abs(p - q) for p, q in zip(cfl, ref["cfl"])) <= 1e-12 * max(ref["cfl"])


@pytest.mark.parametrize("name", ["cube24_fast", "open_box", "amr_3_levels_corner_B"])
def test_pipelined_step_download_equals_step_then_download(name):
    """amps_gpu_step_JM (deposit in block ranges, J/M of finished corners copied out meanwhile) == step + JM_download"""
    import numpy as np
    from amps_b200 import api

    kw = dict(CASES[name])
    if name == "cube24_fast":
        kw["n_cells"] = (32, 32, 32)   # 64 blocks: the schedule really has 8 ranges
    m, cfg, parts, fields = pu.make_case(**kw)
    x, v, w, sp, cells = parts
    E, B, Bcur = fields
    out = []
    for pipelined in (False, True):
        g = api.Context(cfg, m)
        g.fields_upload(E, B, Bcur)
        g.particles_upload(x, v, w, sp, cells)
        if pipelined:
            J, M = np.full((m.n_corners, 3), np.nan), np.full((m.n_corners, 243), np.nan)
            g.step_JM(J, M)
        else:
            g.step()
            J, M = g.JM_download()
        out.append((J, M, g.diagnostics(), g.particles_download()))
        g.close()
    (J0, M0, d0, p0), (J1, M1, d1, p1) = out
    assert not np.isnan(J1).any() and not np.isnan(M1).any()          # every corner was copied
    assert pu.rel_scaled(J1, J0) <= 1e-12 and pu.rel_scaled(M1, M0) <= 1e-12
    assert abs(d0[0] - d1[0]) <= 1e-12 * abs(d0[0])
    assert (np.sort(p0["ptrs"]) == np.sort(p1["ptrs"])).all() and (p0["cells"] == p1["cells"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cube24_fast", "open_box", "ghost2"])
def test_packed_rows_rebuild_the_full_mass_matrix(name):
    """amps_gpu_step_JM_packed / amps_gpu_JM_download_packed ship J and 14 of the 27 neighbour blocks per corner; the other 13 are
    the mirror images (ProcessCell adds the same 3x3 block to both corners of a pair, :2411-2420).  Rebuilt on the host, the rows
    equal the full download: the shipped half exactly, the mirrored half to the summation order of the atomics."""
    import numpy as np
    from amps_b200 import api

    kw = dict(CASES[name])
    if name == "cube24_fast":
        kw["n_cells"] = (32, 32, 32)
    m, cfg, parts, fields = pu.make_case(**kw)
    x, v, w, sp, cells = parts
    E, B, Bcur = fields
    g = api.Context(cfg, m)
    g.fields_upload(E, B, Bcur)
    g.particles_upload(x, v, w, sp, cells)
    packed = np.full((m.n_corners, 129), np.nan)
    g.step_JM_packed(packed)
    J0, M0 = g.JM_download()
    again = g.JM_download_packed()
    J1, M1 = g.expand_packed(packed)
    slots = g.packed_slots()
    g.close()
    assert not np.isnan(packed).any() and (again == packed).all()      # pipelined ranges == one pack of everything
    assert (J1 == J0).all()
    M0r, M1r = M0.reshape(-1, 27, 9), M1.reshape(-1, 27, 9)
    assert (M1r[:, slots, :] == M0r[:, slots, :]).all()
    assert pu.rel_scaled(M1, M0) <= 1e-12
    assert np.abs(M0r[:, [s for s in range(27) if s not in slots], :]).max() > 0
