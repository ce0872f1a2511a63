"""ctypes wrapper of the CPU oracle (TEST INFRASTRUCTURE -- see oracle/amps_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def load(kind="parity"):
    """kind: 'parity' (-O2 -ffp-contract=off) or 'fast' (-O3 -march=x86-64-v3 -fopenmp)."""
    if kind in _libs:
        return _libs[kind]
    path = os.path.join(HERE, "_build", f"liboracle_{kind}.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", HERE])
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.oracle_create.restype = vp
    lib.oracle_create.argtypes = [vp, vp]
    lib.oracle_destroy.argtypes = [vp]
    lib.oracle_last_error.restype = C.c_char_p
    lib.oracle_last_error.argtypes = [vp]
    lib.oracle_particle_data_length.restype = C.c_int64
    lib.oracle_particle_data_length.argtypes = [vp]
    lib.oracle_set_fields.argtypes = [vp, vp, vp, vp]
    lib.oracle_set_background.argtypes = [vp, vp, vp]
    lib.oracle_exit_records.restype = C.c_int64
    lib.oracle_exit_records.argtypes = [vp, vp, C.c_int64]
    lib.oracle_set_background_gca.argtypes = [vp, vp]
    lib.oracle_magnetic_moment_init.argtypes = [vp, C.c_int, vp, C.c_int64]
    lib.oracle_set_background_gradB.argtypes = [vp, vp]
    lib.oracle_get_magnetic_moment.argtypes = [vp, vp, vp, C.c_int64]
    lib.oracle_set_reduced_state.argtypes = [vp, vp, vp, C.c_int64]
    lib.oracle_set_v_normal.argtypes = [vp, vp, C.c_int64]
    lib.oracle_set_E_current.argtypes = [vp, vp]
    lib.oracle_set_global_stencil_length.argtypes = [vp, C.c_int]
    lib.oracle_ecsim_fields.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp]
    lib.oracle_get_v_parallel.argtypes = [vp, vp, C.c_int64]
    lib.oracle_add_particles.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int64]
    lib.oracle_particle_count.restype = C.c_int64
    lib.oracle_particle_count.argtypes = [vp]
    lib.oracle_move.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    lib.oracle_get_particles.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_int64]
    lib.oracle_deposit_JM.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    lib.oracle_find_tree_node.argtypes = [vp, vp, C.c_int]
    lib.oracle_find_cell_index.argtypes = [vp, vp, C.c_int, vp]
    lib.oracle_corner_stencil.argtypes = [vp, vp, C.c_int, vp, vp, vp]
    lib.oracle_center_stencil.argtypes = [vp, vp, C.c_int, vp, vp]
    lib.oracle_check_particle_lists.argtypes = [vp]
    lib.oracle_net_charge.argtypes = [vp, C.c_double, vp]
    lib.oracle_species_moments.argtypes = [vp, vp]
    lib.oracle_sample_cells.argtypes = [vp, vp, vp]
    lib.oracle_set_phi.argtypes = [vp, vp]
    lib.oracle_correct_particle_location.argtypes = [vp, C.c_double, C.c_double, vp, vp, vp]
    lib.oracle_coupler_stencil.argtypes = [vp, vp, C.c_int, vp, vp]
    lib.oracle_neib_levels.argtypes = [vp, C.c_int, vp]
    lib.oracle_neib.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    _libs[kind] = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, cfg, mesh, kind="parity"):
        self.lib = load(kind)
        self.cfg, self.mesh = cfg, mesh
        self.h = self.lib.oracle_create(C.byref(cfg), C.byref(mesh.c))
        self.capacity = int(cfg.capacity)
        self.n_added = 0

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_fields(self, E_half=None, B_prev=None, B_cur=None):
        a = [None if t is None else np.ascontiguousarray(t, dtype=np.float64) for t in (E_half, B_prev, B_cur)]
        self.lib.oracle_set_fields(self.h, _p(a[0]), _p(a[1]), _p(a[2]))

    def set_background(self, E_center=None, B_center=None):
        a = [None if t is None else np.ascontiguousarray(t, dtype=np.float64) for t in (E_center, B_center)]
        self.lib.oracle_set_background(self.h, _p(a[0]), _p(a[1]))

    def set_background_gca(self, var15):
        a = np.ascontiguousarray(var15, dtype=np.float64)
        self.lib.oracle_set_background_gca(self.h, _p(a))

    def set_background_gradB(self, gradB):
        a = np.ascontiguousarray(gradB, dtype=np.float64)
        self.lib.oracle_set_background_gradB(self.h, _p(a))

    def magnetic_moment_init(self, mover=5):
        mu = np.zeros(self.n_added)
        rc = self.lib.oracle_magnetic_moment_init(self.h, mover, _p(mu), self.n_added)
        assert rc == 0
        return mu

    def magnetic_moment(self):
        mu = np.zeros(self.n_added)
        flag = np.zeros(self.n_added, dtype=np.uint8)
        self.lib.oracle_get_magnetic_moment(self.h, _p(mu), _p(flag), self.n_added)
        return mu, flag

    def set_E_current(self, E):
        a = np.ascontiguousarray(E, dtype=np.float64)
        assert a.shape == (self.mesh.n_corners, 3)
        self.lib.oracle_set_E_current(self.h, _p(a))

    def ecsim_fields(self, x, leaf):
        """ECSIM::GetElectricField / GetMagneticField / GetMagneticFieldGradient at the points x[n][3] (each inside block leaf[n])"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        leaf = np.ascontiguousarray(leaf, dtype=np.int32)
        n = x.shape[0]
        E, B, G = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 9))
        bad = self.lib.oracle_ecsim_fields(self.h, n, _p(x), _p(leaf), _p(E), _p(B), _p(G))
        return E, B, G, bad

    def set_global_stencil_length(self, length):
        self.lib.oracle_set_global_stencil_length(self.h, int(length))

    def set_v_normal(self, vnormal):
        a = np.ascontiguousarray(vnormal, dtype=np.float64)
        assert a.shape == (self.n_added,)
        self.lib.oracle_set_v_normal(self.h, _p(a), self.n_added)

    def set_reduced_state(self, mu, vpar):
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        vpar = np.ascontiguousarray(vpar, dtype=np.float64)
        assert mu.shape == (self.n_added,) and vpar.shape == (self.n_added,)
        self.lib.oracle_set_reduced_state(self.h, _p(mu), _p(vpar), self.n_added)

    def v_parallel(self):
        a = np.empty(self.n_added)
        self.lib.oracle_get_v_parallel(self.h, _p(a), self.n_added)
        return a

    def exit_records(self, max_records=1 << 20):
        from amps_b200._capi import ExitRecord

        buf = (ExitRecord * max_records)()
        n = int(self.lib.oracle_exit_records(self.h, C.cast(buf, C.c_void_p), max_records))
        return n, [(r.ptr, r.species, r.face, r.leaf, tuple(r.x), tuple(r.v)) for r in buf[:min(n, max_records)]]

    def add_particles(self, x, v, w, species, cells):
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        n = x.shape[1]
        w = None if w is None else np.ascontiguousarray(w, dtype=np.float64)
        species = np.ascontiguousarray(species, dtype=np.uint8)
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        rc = self.lib.oracle_add_particles(self.h, _p(x), _p(v), _p(w), _p(species), _p(cells), n)
        assert rc == 0, self.lib.oracle_last_error(self.h)
        self.n_added += n

    def move(self, mover=0, n_threads=1, want_stats=True):
        from amps_b200._capi import MoveStats

        st = MoveStats()
        ret = np.zeros(self.capacity, dtype=np.int32)
        fc = np.zeros(self.capacity, dtype=np.int32)
        rc = self.lib.oracle_move(self.h, mover, n_threads, C.cast(C.byref(st), C.c_void_p) if want_stats else None, _p(ret), _p(fc))
        return rc, st.as_dict(), ret[: self.n_added], fc[: self.n_added]

    def move_fast(self, mover=0, n_threads=1):
        """timing path: no per-particle outputs, no statistics pass"""
        return self.lib.oracle_move(self.h, mover, n_threads, None, None, None)

    def particles(self):
        n = self.n_added
        x, v, w = np.empty((3, n)), np.empty((3, n)), np.empty(n)
        sp = np.empty(n, dtype=np.uint8)
        cells = np.empty(n, dtype=np.int32)
        alive = np.empty(n, dtype=np.uint8)
        self.lib.oracle_get_particles(self.h, _p(x), _p(v), _p(w), _p(sp), _p(cells), _p(alive), n)
        return {"x": x, "v": v, "w": w, "species": sp, "cells": cells, "alive": alive}

    def deposit(self, n_threads=1, want_arrays=True):
        nc = self.mesh.n_corners
        J = np.empty((nc, 3)) if want_arrays else None
        M = np.empty((nc, 243)) if want_arrays else None
        e = C.c_double()
        cfl = (C.c_double * 8)()
        self.lib.oracle_deposit_JM(self.h, n_threads, _p(J), _p(M), C.cast(C.byref(e), C.c_void_p), C.cast(cfl, C.c_void_p))
        return J, M, float(e.value), [float(cfl[s]) for s in range(self.cfg.n_species)]

    def find_tree_node(self, x, start_leaf=-1):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return self.lib.oracle_find_tree_node(self.h, _p(x), start_leaf)

    def find_cell_index(self, x, leaf):
        x = np.ascontiguousarray(x, dtype=np.float64)
        ijk = np.zeros(3, dtype=np.int32)
        r = self.lib.oracle_find_cell_index(self.h, _p(x), leaf, _p(ijk))
        return r, ijk

    def corner_stencil(self, x, leaf):
        x = np.array(x, dtype=np.float64)
        W = np.zeros(8)
        ids = np.zeros(8, dtype=np.int32)
        wn = np.zeros(8)
        n = self.lib.oracle_corner_stencil(self.h, _p(x), leaf, _p(W), _p(ids), _p(wn))
        return n, x, W, ids, wn

    def center_stencil(self, x, leaf):
        x = np.ascontiguousarray(x, dtype=np.float64)
        ids = np.zeros(64, dtype=np.int32)
        w = np.zeros(64)
        n = self.lib.oracle_center_stencil(self.h, _p(x), leaf, _p(ids), _p(w))
        return n, ids[:n], w[:n]

    def coupler_stencil(self, x, leaf):
        x = np.ascontiguousarray(x, dtype=np.float64)
        ids = np.zeros(64, dtype=np.int32)
        w = np.zeros(64)
        n = self.lib.oracle_coupler_stencil(self.h, _p(x), leaf, _p(ids), _p(w))
        return n, ids[:max(n, 0)], w[:max(n, 0)]

    def neib(self, leaf, kind, idx):
        lo, hi = np.zeros(3), np.zeros(3)
        lev = C.c_int()
        r = self.lib.oracle_neib(self.h, leaf, kind, idx, _p(lo), _p(hi), C.cast(C.byref(lev), C.c_void_p))
        return None if r < 0 else (lo, hi, int(lev.value))

    def neib_levels(self, leaf):
        mm = np.zeros(2, dtype=np.int32)
        self.lib.oracle_neib_levels(self.h, leaf, _p(mm))
        return int(mm[0]), int(mm[1])

    def net_charge(self, charge_conv=1.0):
        rho = np.zeros(self.mesh.n_centers)
        assert self.lib.oracle_net_charge(self.h, charge_conv, _p(rho)) == 0
        return rho

    def sample_cells(self):
        """PIC::Sampling: one more sample in the collecting buffer -> (buffer [n_cells, n_species, 13], particles sampled so far per species)"""
        n_cells = self.mesh.c.n_leaves * int(np.prod(self.mesh.block_cells))
        out = np.zeros((n_cells, self.cfg.n_species, 13))
        cnt = np.zeros(self.cfg.n_species, dtype=np.int64)
        assert self.lib.oracle_sample_cells(self.h, _p(out), _p(cnt)) == 0
        return out, cnt

    def species_moments(self):
        """corner species moments of UpdateJMassMatrix (_PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_): [n_corners, n_species, 10]"""
        out = np.zeros((self.mesh.n_corners, self.cfg.n_species, 10))
        assert self.lib.oracle_species_moments(self.h, _p(out)) == 0
        return out

    def set_phi(self, phi_center):
        phi = np.ascontiguousarray(phi_center, dtype=np.float64)
        assert phi.shape == (self.mesh.n_centers,)
        self.lib.oracle_set_phi(self.h, _p(phi))

    def correct_particle_location(self, charge_conv=1.0, mass_conv=1.0):
        """ECSIM::CorrectParticleLocation -> (rc, n_displaced, n_deleted, final_cell[n_added])"""
        fc = np.zeros(self.capacity, dtype=np.int32)
        nd, nx = C.c_int64(), C.c_int64()
        rc = self.lib.oracle_correct_particle_location(self.h, charge_conv, mass_conv, _p(fc), C.cast(C.byref(nd), C.c_void_p),
                                                       C.cast(C.byref(nx), C.c_void_p))
        return rc, int(nd.value), int(nx.value), fc[: self.n_added]

    def check_lists(self):
        return self.lib.oracle_check_particle_lists(self.h)
