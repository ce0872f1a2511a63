"""The C++ host layer (amps_b200/host/amps_gpu_host.hpp) on an AMPS-layout particle buffer: compiled with g++, driven by
tests/cpp/host_roundtrip.cpp, compared with the CPU oracle.  Without a GPU the same binary must fail loudly (no CPU fallback)."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from amps_b200 import _capi
from tests import parity_util as pu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MESH_ARRAYS = ["node_parent", "node_child", "node_level", "node_imin", "node_isize", "node_xmin", "node_xmax", "node_leaf", "node_flags",
               "node_thread", "root_node", "leaf_node", "leaf_real", "leaf_face_boundary", "leaf_corner_uid", "leaf_center_uid"]
# packed basic record of the reference (picParticleDataMacro.h:55-81): next, prev, species byte, v, x, weight correction
OFF_NEXT, OFF_PREV, OFF_SPEC, OFF_V, OFF_X, OFF_W, STRIDE = 0, 8, 16, 17, 41, 65, 73


def build_driver(tmp_path):
    exe = str(tmp_path / "host_roundtrip")
    libdir = os.path.join(ROOT, "amps_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "amps_b200", "host"),
                           os.path.join(ROOT, "tests", "cpp", "host_roundtrip.cpp"), "-o", exe, "-L", libdir, "-lamps_gpu",
                           "-Wl,-rpath," + libdir])
    return exe


def blob(f, b):
    b = bytes(b)
    f.write(struct.pack("<q", len(b)))
    f.write(b)


def write_case(path, m, cfg, parts, fields):
    x, v, w, sp, cells = parts
    n = x.shape[1]
    cap = n + 5
    buf = np.zeros((cap, STRIDE), dtype=np.uint8)
    rec = buf[:n]
    rec[:, OFF_V:OFF_V + 24] = np.ascontiguousarray(v.T).view(np.uint8).reshape(n, 24)
    rec[:, OFF_X:OFF_X + 24] = np.ascontiguousarray(x.T).view(np.uint8).reshape(n, 24)
    rec[:, OFF_W:OFF_W + 8] = np.ascontiguousarray(w).view(np.uint8).reshape(n, 8)
    rec[:, OFF_SPEC] = sp | 0x80                      # allocated flag (bit 7)
    # per-cell lists as InitiateParticle builds them: every new particle becomes the head of its cell's list
    order = np.argsort(cells, kind="stable")
    cs = cells[order]
    nxt = -np.ones(n, dtype=np.int64)
    prv = -np.ones(n, dtype=np.int64)
    same = cs[1:] == cs[:-1]
    nxt[order[1:][same]] = order[:-1][same]           # the particle added before me (same cell) follows me
    prv[order[:-1][same]] = order[1:][same]
    first = -np.ones(m.n_cells, dtype=np.int64)
    last_of_cell = np.r_[~same, True]
    first[cs[last_of_cell]] = order[last_of_cell]
    rec[:, OFF_NEXT:OFF_NEXT + 8] = nxt.view(np.uint8).reshape(n, 8)
    rec[:, OFF_PREV:OFF_PREV + 8] = prv.view(np.uint8).reshape(n, 8)
    lay = _capi.AosLayout()
    lay.stride, lay.off_species, lay.off_v, lay.off_x, lay.off_w, lay.off_mu, lay.off_next, lay.off_prev = STRIDE, OFF_SPEC, OFF_V, OFF_X, OFF_W, -1, OFF_NEXT, OFF_PREV
    lay.off_vpar = -1
    with open(path, "wb") as f:
        blob(f, bytes(cfg))
        c = m.c
        scal = struct.pack("<8i13d", c.n_root[0], c.n_root[1], c.n_root[2], c.max_refinement_level, c.n_nodes, c.n_leaves, c.n_corners, c.n_centers,
                           *[c.x_global_min[d] for d in range(3)], *[c.x_global_max[d] for d in range(3)],
                           *[c.dx_max_refinement[d] for d in range(3)], *[c.dx_root_block[d] for d in range(3)], c.eps)
        blob(f, scal)
        for name in MESH_ARRAYS:
            blob(f, np.ascontiguousarray(m.arrays[name]).tobytes())
        blob(f, bytes(lay))
        blob(f, buf.tobytes())
        blob(f, first.tobytes())
        for a in fields:
            blob(f, np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return cap


def read_blobs(path):
    out = []
    with open(path, "rb") as f:
        while True:
            h = f.read(8)
            if len(h) < 8:
                break
            (k,) = struct.unpack("<q", h)
            out.append(f.read(k))
    return out


def test_cpp_host_compiles_and_fails_loudly_without_a_gpu(tmp_path):
    import torch

    exe = build_driver(tmp_path)
    if torch.cuda.is_available():
        return
    m, cfg, parts, fields = pu.make_case(n_cells=(8, 8, 8), ppc=1)
    write_case(str(tmp_path / "case.bin"), m, cfg, parts, fields)
    r = subprocess.run([exe, str(tmp_path / "case.bin"), str(tmp_path / "out.bin")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "HOST_ROUNDTRIP_OK" not in r.stdout      # std::runtime_error from amps_gpu_init: no CPU fallback
    assert not os.path.exists(tmp_path / "out.bin")


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("periodic", dict(n_cells=(16, 16, 16), ppc=6, seed=31)),
                                     ("open_corner_B", dict(n_cells=(16, 16, 16), ppc=5, seed=33, periodic=False, vscale=5.0, b_mode=1))])
def test_cpp_host_matches_oracle(tmp_path, name, kw):
    exe = build_driver(tmp_path)
    m, cfg, parts, fields = pu.make_case(**kw)
    cap = write_case(str(tmp_path / "case.bin"), m, cfg, parts, fields)
    r = subprocess.run([exe, str(tmp_path / "case.bin"), str(tmp_path / "out.bin")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "HOST_ROUNDTRIP_OK" in r.stdout, r.stderr
    st_b, n_b, buf_b, first_b, J_b, M_b, e_b, cfl_b, rho_b, smp_b, cnt_b = read_blobs(str(tmp_path / "out.bin"))
    ora = pu.run_oracle(m, cfg, parts, fields)
    # ComputeNetCharge and Sampling of the host layer after the move: against the oracle in the same state
    from oracle.oracle_py import Oracle
    o = Oracle(cfg, m)
    o.set_fields(*fields)
    o.add_particles(*parts)
    o.move(0, 1)
    rho_ref = o.net_charge(0.7)
    smp_ref, cnt_ref = o.sample_cells()
    o.close()
    assert pu.rel_scaled(np.frombuffer(rho_b), rho_ref) <= pu.REL_TOL
    smp = np.frombuffer(smp_b).reshape(smp_ref.shape)
    assert (smp[:, :, 1] == smp_ref[:, :, 1]).all() and pu.rel_scaled(smp, smp_ref) <= 1e-12
    assert (np.frombuffer(cnt_b, dtype=np.int64)[: cfg.n_species] == cnt_ref).all()
    n = parts[0].shape[1]
    stats = struct.unpack("<8q", st_b[:64])
    keys = ["n_moved", "n_cross_cell", "n_cross_block", "n_left_domain", "n_not_in_use", "n_periodic_wrap", "n_error", "n_sub_steps"]
    assert dict(zip(keys, stats)) == ora["stats"]
    buf = np.frombuffer(buf_b, dtype=np.uint8).reshape(cap, STRIDE)
    first = np.frombuffer(first_b, dtype=np.int64)
    nxt = buf[:, OFF_NEXT:OFF_NEXT + 8].copy().view(np.int64).ravel()
    prv = buf[:, OFF_PREV:OFF_PREV + 8].copy().view(np.int64).ravel()
    # walk the rebuilt lists: every surviving particle hangs on exactly the cell the oracle put it in
    cell_of = -np.ones(n, dtype=np.int64)
    for c in np.nonzero(first >= 0)[0]:
        p, prev = int(first[c]), -1
        while p != -1:
            assert cell_of[p] == -1 and prv[p] == prev
            cell_of[p] = c
            prev, p = p, int(nxt[p])
    oc = ora["final_cell"].astype(np.int64)
    assert (cell_of == oc).all()
    alive = oc >= 0
    assert struct.unpack("<q", n_b)[0] == int(alive.sum())
    gx = buf[:n, OFF_X:OFF_X + 24].copy().view(np.float64).reshape(n, 3).T
    gv = buf[:n, OFF_V:OFF_V + 24].copy().view(np.float64).reshape(n, 3).T
    assert pu.rel_elementwise(gx[:, alive], ora["particles"]["x"][:, alive]) <= pu.REL_TOL
    assert pu.rel_elementwise(gv[:, alive], ora["particles"]["v"][:, alive]) <= pu.REL_TOL
    J = np.frombuffer(J_b).reshape(-1, 3)
    M = np.frombuffer(M_b).reshape(-1, 243)
    assert pu.rel_scaled(J, ora["J"]) <= pu.REL_TOL and pu.rel_scaled(M, ora["M"]) <= pu.REL_TOL
    (e,) = struct.unpack("<d", e_b)
    assert abs(e - ora["energy"]) <= pu.REL_TOL * abs(ora["energy"])
