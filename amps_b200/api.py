"""Thin Python handle over the C ABI (include/amps_gpu.h) used by tests and bench.py.

Method names follow the reference entry points they stand for:
  MoveParticles()        <- PIC::Mover::MoveParticles            (src/pic/pic_mover.cpp:580)
  UpdateJMassMatrix()    <- ECSIM::UpdateJMassMatrix             (src/pic/pic_field_solver_ecsim.cpp:3244)
  ParticleBuffer upload/download <- PIC::ParticleBuffer          (src/pic/pic_pbuffer.cpp)
Everything here is plumbing: numpy arrays in, ctypes calls, numpy arrays out.  The compute is in
libamps_gpu.so; if it is missing or there is no GPU these calls raise.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import Config, MoveStats


class AmpsGpuError(RuntimeError):
    pass


def make_config(block_cells=(8, 8, 8), ghost_cells=(1, 1, 1), charge=(-1.0, 1.0), mass=(1.0, 1836.0), species_weight=None, dt=1.0,
                periodic=True, capacity=1 << 20, device=0, boundary_mode=_capi.BOUNDARY_DELETE, B_conv=1.0, length_conv=1.0,
                light_speed=1.0):
    """Normalised-unit ECSIM configuration (_PIC_FIELD_SOLVER_INPUT_UNIT_NORM_, c = 1)."""
    cfg = Config()
    ns = len(charge)
    for d in range(3):
        cfg.block_cells[d] = block_cells[d]
        cfg.ghost_cells[d] = ghost_cells[d]
    cfg.n_species = ns
    cfg.b_mode = _capi.B_CENTER_BASED
    cfg.periodic = 1 if periodic else 0
    cfg.boundary_mode = boundary_mode
    cfg.time_step_mode = _capi.DT_SINGLE_GLOBAL
    cfg.device = device
    cfg.capacity = int(capacity)
    for s in range(ns):
        cfg.charge[s] = charge[s]
        cfg.mass[s] = mass[s]
        cfg.species_weight[s] = 1.0 if species_weight is None else species_weight[s]
        cfg.time_step[s] = dt
    cfg.ecsim_dt_total = dt
    cfg.ecsim_B_conv = B_conv
    cfg.ecsim_length_conv = length_conv
    cfg.ecsim_light_speed = light_speed
    cfg.coupler_interpolation = _capi.CPLR_LINEAR
    cfg.backward_time_integration = 0
    cfg.speed_of_light = 299792458.0
    cfg.internal_sphere_radius = 0.0
    cfg.exit_record_capacity = 0
    cfg.gravity_gm = 0.0
    cfg.carry_magnetic_moment = 0
    cfg.carry_v_parallel = 0
    cfg.ideal_mhd = 1
    cfg.gc_species_mask = 0
    cfg.gc_fields_ecsim = 0
    cfg.exact_arithmetic = 0
    return cfg


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One device context == one AMPS rank's particle store, mesh copy and J/M arrays."""

    def __init__(self, cfg, mesh):
        self.lib = _capi.load_library()
        self.cfg = cfg
        self.mesh = mesh
        self._h = C.c_void_p()
        rc = self.lib.amps_gpu_init(C.byref(cfg), C.byref(self._h))
        if rc != _capi.OK:
            msg = self.lib.amps_gpu_last_error(self._h).decode() if self._h else ""
            raise AmpsGpuError(f"amps_gpu_init failed rc={rc} {msg} (no CPU fallback)")
        self._ck(self.lib.amps_gpu_mesh_upload(self._h, C.byref(mesh.c)))

    def mesh_upload(self, mesh):
        """a new mesh epoch (UpdateBlockTable after nMeshModificationCounter changed): the resident particles and fields are dropped
        with the old mesh and must be uploaded again"""
        self._ck(self.lib.amps_gpu_mesh_upload(self._h, C.byref(mesh.c)))
        self.mesh = mesh

    def _ck(self, rc):
        if rc != _capi.OK:
            raise AmpsGpuError(f"rc={rc}: {self.lib.amps_gpu_last_error(self._h).decode()}")

    def close(self):
        if self._h:
            self.lib.amps_gpu_finalize(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- fields ------------------------------------------------------------------------
    def fields_upload(self, E_half=None, B_prev=None, B_cur=None):
        arrs = []
        nb = self.mesh.n_corners if self.cfg.b_mode == _capi.B_CORNER_BASED else self.mesh.n_centers
        for a, n in ((E_half, self.mesh.n_corners), (B_prev, nb), (B_cur, nb)):
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float64)
                assert a.shape == (n, 3), (a.shape, n)
            arrs.append(a)
        self._ck(self.lib.amps_gpu_fields_upload(self._h, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2])))

    def background_upload(self, E_center=None, B_center=None):
        """coupler table (E, B on unique centre nodes) of the test-particle movers"""
        arrs = []
        for a in (E_center, B_center):
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float64)
                assert a.shape == (self.mesh.n_centers, 3)
            arrs.append(a)
        self._ck(self.lib.amps_gpu_background_upload(self._h, _ptr(arrs[0]), _ptr(arrs[1])))

    def background_upload_gca(self, var15_center):
        """15 drift variables of Relativistic::GuidingCenter on the unique centre nodes (pic.h:8643-8680)"""
        a = np.ascontiguousarray(var15_center, dtype=np.float64)
        assert a.shape == (self.mesh.n_centers, 15)
        self._ck(self.lib.amps_gpu_background_upload_gca(self._h, _ptr(a)))

    def background_upload_gradB(self, gradB_center):
        """grad B on the unique centre nodes, [n_centers][9] (pic.h:8434-8470)"""
        a = np.ascontiguousarray(gradB_center, dtype=np.float64)
        assert a.shape == (self.mesh.n_centers, 9)
        self._ck(self.lib.amps_gpu_background_upload_gradB(self._h, _ptr(a)))

    def InitiateMagneticMoment(self, mover=_capi.MOVER_RELATIVISTIC_GCA):
        """(Relativistic::)GuidingCenter::InitiateMagneticMoment for every resident particle"""
        self._ck(self.lib.amps_gpu_magnetic_moment_init(self._h, mover))

    def magnetic_moment_upload(self, mu_by_ptr):
        a = np.ascontiguousarray(mu_by_ptr, dtype=np.float64)
        self._ck(self.lib.amps_gpu_magnetic_moment_upload(self._h, _ptr(a), a.shape[0]))

    def magnetic_moment_download(self):
        """mu in the current device order (pair with particles_download()['ptrs'])"""
        n = self.particle_count()
        mu = np.empty(max(n, 1))
        k = C.c_int64()
        self._ck(self.lib.amps_gpu_magnetic_moment_download(self._h, _ptr(mu), mu.shape[0], C.byref(k)))
        return mu[: int(k.value)]

    def global_stencil_set(self, full):
        """the reference's global StencilTable holds an 8-cell stencil (ComputeNetCharge ran): full B stencils stay un-normalised"""
        self._ck(self.lib.amps_gpu_global_stencil_set(self._h, 1 if full else 0))

    def v_normal_upload(self, vnormal_by_ptr):
        """PB::GetVNormal of the guiding-centre species by ParticleBuffer slot (read by the deposit's diagnostics)"""
        a = np.ascontiguousarray(vnormal_by_ptr, dtype=np.float64)
        self._ck(self.lib.amps_gpu_v_normal_upload(self._h, _ptr(a), a.shape[0]))

    def v_parallel_upload(self, vpar_by_ptr):
        a = np.ascontiguousarray(vpar_by_ptr, dtype=np.float64)
        self._ck(self.lib.amps_gpu_v_parallel_upload(self._h, _ptr(a), a.shape[0]))

    def v_parallel_download(self):
        """v_parallel in the current device order (pair with particles_download()['ptrs'])"""
        n = self.particle_count()
        a = np.empty(max(n, 1))
        k = C.c_int64()
        self._ck(self.lib.amps_gpu_v_parallel_download(self._h, _ptr(a), a.shape[0], C.byref(k)))
        return a[: int(k.value)]

    def exit_records(self, max_records=1 << 20):
        buf = (_capi.ExitRecord * max_records)()
        n = C.c_int64()
        self._ck(self.lib.amps_gpu_exit_records(self._h, C.cast(buf, C.c_void_p), max_records, C.byref(n)))
        k = min(int(n.value), max_records, int(self.cfg.exit_record_capacity))
        return int(n.value), [(r.ptr, r.species, r.face, r.leaf, tuple(r.x), tuple(r.v)) for r in buf[:k]]

    # ---- PIC::ParticleBuffer ---------------------------------------------------------------
    def particles_upload(self, x, v, w, species, cells, ptrs=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        n = x.shape[1]
        assert x.shape == (3, n) and v.shape == (3, n)
        w = None if w is None else np.ascontiguousarray(w, dtype=np.float64)
        species = np.ascontiguousarray(species, dtype=np.uint8)
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        ptrs = None if ptrs is None else np.ascontiguousarray(ptrs, dtype=np.int32)
        self._ck(self.lib.amps_gpu_particles_upload_soa(self._h, _ptr(x), _ptr(v), _ptr(w), _ptr(species), _ptr(cells), _ptr(ptrs), n))

    def particles_append(self, x, v, w, species, cells, ptrs=None):
        """the same behind the resident particles (injection; large populations in pieces)"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        n = x.shape[1]
        w = None if w is None else np.ascontiguousarray(w, dtype=np.float64)
        species = np.ascontiguousarray(species, dtype=np.uint8)
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        ptrs = None if ptrs is None else np.ascontiguousarray(ptrs, dtype=np.int32)
        self._ck(self.lib.amps_gpu_particles_append_soa(self._h, _ptr(x), _ptr(v), _ptr(w), _ptr(species), _ptr(cells), _ptr(ptrs), n))

    def particle_count(self):
        n = C.c_int64()
        self._ck(self.lib.amps_gpu_particle_count(self._h, C.byref(n)))
        return int(n.value)

    def particles_download(self):
        n = self.particle_count()
        x = np.empty((3, n)), np.empty((3, n))
        w = np.empty(n)
        sp = np.empty(n, dtype=np.uint8)
        cells = np.empty(n, dtype=np.int32)
        ptrs = np.empty(n, dtype=np.int32)
        nn = C.c_int64()
        self._ck(self.lib.amps_gpu_particles_download_soa(self._h, _ptr(x[0]), _ptr(x[1]), _ptr(w), _ptr(sp), _ptr(cells), _ptr(ptrs), n,
                                                          C.byref(nn)))
        return {"x": x[0], "v": x[1], "w": w, "species": sp, "cells": cells, "ptrs": ptrs}

    def slot_delta(self):
        """(n_new, released): records without a ParticleBuffer slot (arrivals) and the slots whose particle is gone."""
        n_new, n_rel = C.c_int64(), C.c_int64()
        self._ck(self.lib.amps_gpu_particles_slot_delta(self._h, C.byref(n_new), None, 0, C.byref(n_rel)))
        rel = np.empty(max(1, int(n_rel.value)), dtype=np.int64)
        self._ck(self.lib.amps_gpu_particles_slot_delta(self._h, C.byref(n_new), _ptr(rel), rel.size, C.byref(n_rel)))
        return int(n_new.value), rel[: int(n_rel.value)].copy()

    def assign_slots(self, slots):
        slots = np.ascontiguousarray(slots, dtype=np.int64)
        self._ck(self.lib.amps_gpu_particles_assign_slots(self._h, _ptr(slots), slots.size))

    # ---- PIC::Restart (particle file of the reference's format) -------------------------------
    def restart_save(self, fname, header, leaf_node_ids, lay):
        ids = np.ascontiguousarray(leaf_node_ids, dtype=np.uint8)
        assert ids.shape[0] == self.mesh.n_leaves
        n = C.c_int64()
        hb = bytes(header)
        self._ck(self.lib.amps_gpu_restart_save(self._h, fname.encode(), C.c_char_p(hb), len(hb), _ptr(ids), ids.shape[1], C.byref(lay), C.byref(n)))
        return int(n.value)

    def restart_read(self, fname, header_bytes, leaf_node_ids, lay):
        ids = np.ascontiguousarray(leaf_node_ids, dtype=np.uint8)
        n = C.c_int64()
        self._ck(self.lib.amps_gpu_restart_read(self._h, fname.encode(), int(header_bytes), _ptr(ids), ids.shape[1], C.byref(lay), C.byref(n)))
        return int(n.value)

    def cell_table(self):
        t = np.empty(self.mesh.n_cells + 1, dtype=np.int64)
        self._ck(self.lib.amps_gpu_cell_table_download(self._h, _ptr(t), t.size))
        return t

    def sort(self):
        self._ck(self.lib.amps_gpu_sort(self._h))

    # ---- PIC::Mover::MoveParticles -----------------------------------------------------------
    def MoveParticles(self, mover=_capi.MOVER_LAPENTA2017, stats=True, raise_on_particle_error=True):
        if stats:
            st = MoveStats()
            rc = self.lib.amps_gpu_move(self._h, mover, C.byref(st))
            if rc == _capi.ERR_PARTICLE and not raise_on_particle_error:
                return st.as_dict()  # n_error particles hit a place where the reference exit()s; they were dropped
            self._ck(rc)
            return st.as_dict()
        self._ck(self.lib.amps_gpu_move(self._h, mover, None))
        return None

    # ---- ECSIM::UpdateJMassMatrix ---------------------------------------------------------
    def UpdateJMassMatrix(self, diagnostics=True):
        if diagnostics:
            e = C.c_double()
            cfl = (C.c_double * _capi.MAX_SPECIES)()
            self._ck(self.lib.amps_gpu_deposit_JM(self._h, C.cast(C.byref(e), C.c_void_p), C.cast(cfl, C.c_void_p)))
            return float(e.value), [float(cfl[s]) for s in range(self.cfg.n_species)]
        self._ck(self.lib.amps_gpu_deposit_JM(self._h, None, None))
        return None

    def JM_download(self, want_J=True, want_M=True, out_J=None, out_M=None):
        nc = self.mesh.n_corners
        J = (out_J if out_J is not None else np.empty((nc, 3))) if want_J else None
        M = (out_M if out_M is not None else np.empty((nc, 243))) if want_M else None
        self._ck(self.lib.amps_gpu_JM_download(self._h, _ptr(J), _ptr(M)))
        return J, M

    def step_JM(self, out_J, out_M, mover=_capi.MOVER_LAPENTA2017):
        """step() + JM_download() with the download pipelined behind the deposit (pin out_J / out_M for real overlap)"""
        assert out_J.shape == (self.mesh.n_corners, 3) and out_M.shape == (self.mesh.n_corners, 243)
        self._ck(self.lib.amps_gpu_step_JM(self._h, mover, _ptr(out_J), _ptr(out_M)))
        return out_J, out_M

    def packed_slots(self):
        p = self.lib.amps_gpu_JM_packed_slots()
        return [int(p[i]) for i in range(14)]

    def JM_download_packed(self):
        out = np.empty((self.mesh.n_corners, 129))
        self._ck(self.lib.amps_gpu_JM_download_packed(self._h, _ptr(out)))
        return out

    def step_JM_packed(self, out, mover=_capi.MOVER_LAPENTA2017):
        """step() + the packed J/M rows (J[3] + 14 of the 27 neighbour blocks), download pipelined behind the deposit"""
        assert out.shape == (self.mesh.n_corners, 129)
        self._ck(self.lib.amps_gpu_step_JM_packed(self._h, mover, _ptr(out)))
        return out

    def expand_packed(self, packed):
        """host side of the packed format (what the AMPS shim does while scattering into the corner buffers): J [n,3], M [n,243]"""
        nb = self.mesh.corner_neighbours()  # [n_corners, 27] unique corner at offset slot, -1 outside
        slots = self.packed_slots()
        n = packed.shape[0]
        J = packed[:, :3].copy()
        M = np.zeros((n, 27, 9))
        blocks = packed[:, 3:].reshape(n, 14, 9)
        for b, s in enumerate(slots):
            M[:, s, :] = blocks[:, b, :]
        code = {0: 0, -1: 1, 1: 2}
        inv = {v: k for k, v in code.items()}
        for s in range(27):
            if s in slots:
                continue
            d = (inv[s % 3], inv[(s // 3) % 3], inv[s // 9])
            so = code[-d[0]] + 3 * code[-d[1]] + 9 * code[-d[2]]  # the opposite slot is one of the 14
            b = slots.index(so)
            partner = nb[:, s]
            ok = partner >= 0
            M[ok, s, :] = blocks[partner[ok], b, :]
        return J, M.reshape(n, 243)

    def ComputeNetCharge(self, charge_conv=1.0):
        """ECSIM::ComputeNetCharge: rho_new on the unique centre nodes"""
        rho = np.empty(self.mesh.n_centers)
        self._ck(self.lib.amps_gpu_net_charge(self._h, charge_conv, _ptr(rho)))
        return rho

    def SampleCells(self):
        """PIC::Sampling: add one sample of the resident particles to the collecting buffer on the device"""
        self._ck(self.lib.amps_gpu_sample_cells(self._h))

    def sample_download(self, clear=False):
        """-> (buffer [n_cells, n_species, 13], particles sampled per species)"""
        n_cells = self.mesh.c.n_leaves * int(np.prod(self.mesh.block_cells))
        out = np.empty((n_cells, self.cfg.n_species, 13))
        cnt = np.zeros(_capi.MAX_SPECIES, dtype=np.int64)
        self._ck(self.lib.amps_gpu_sample_download(self._h, _ptr(out), _ptr(cnt), 1 if clear else 0))
        return out, cnt[: self.cfg.n_species]

    def ComputeSpeciesMoments(self, download=True):
        """corner species moments of UpdateJMassMatrix (_PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_): [n_corners, n_species, 10]"""
        out = np.empty((self.mesh.n_corners, self.cfg.n_species, 10)) if download else None
        self._ck(self.lib.amps_gpu_species_moments(self._h, _ptr(out) if download else None))
        return out

    def SetPhi(self, phi_center):
        phi = np.ascontiguousarray(phi_center, dtype=np.float64)
        assert phi.shape == (self.mesh.n_centers,)
        self._ck(self.lib.amps_gpu_phi_upload(self._h, _ptr(phi)))

    def CorrectParticleLocation(self, charge_conv=1.0, mass_conv=1.0):
        """ECSIM::CorrectParticleLocation -> (n_displaced, n_deleted); call sort() afterwards"""
        nd, nx = C.c_int64(), C.c_int64()
        self._ck(self.lib.amps_gpu_correct_particle_location(self._h, charge_conv, mass_conv, C.cast(C.byref(nd), C.c_void_p),
                                                            C.cast(C.byref(nx), C.c_void_p)))
        return int(nd.value), int(nx.value)

    def diagnostics(self):
        e = C.c_double()
        cfl = (C.c_double * _capi.MAX_SPECIES)()
        self._ck(self.lib.amps_gpu_diagnostics(self._h, C.cast(C.byref(e), C.c_void_p), C.cast(cfl, C.c_void_p)))
        return float(e.value), [float(cfl[s]) for s in range(self.cfg.n_species)]

    def step(self, mover=_capi.MOVER_LAPENTA2017):
        self._ck(self.lib.amps_gpu_step(self._h, mover))

    # ---- ECSIM::TimeStep, the field half (row f1) -------------------------------------------------
    def comm_uses_peer_memory(self):
        return bool(self.lib.amps_gpu_comm_uses_peer_memory(self._h))

    def field_solver_init(self, dist=None):
        """dist (torch.distributed, several ranks): gathers the global node keys that define the field halo lists"""
        nb, cc, zc = self.mesh.field_solver_tables()
        self._ck(self.lib.amps_gpu_field_solver_init(self._h, _ptr(np.ascontiguousarray(nb)), _ptr(np.ascontiguousarray(cc)), _ptr(np.ascontiguousarray(zc))))
        if self.mesh.n_ranks > 1:
            from . import mesh as meshmod

            m = self.mesh
            gathered = [None] * m.n_ranks
            dist.all_gather_object(gathered, (m.corner_gkey, m.corner_target_gkeys, m.center_gkey, m.center_own_gkeys))
            mask, lists = meshmod.field_halo_lists(m, gathered)
            self._ck(self.lib.amps_gpu_field_primary_set(self._h, _ptr(np.ascontiguousarray(mask, dtype=np.uint8))))
            for peer, ls in lists.items():
                a = [np.ascontiguousarray(v, dtype=np.int32) for v in ls]
                self._ck(self.lib.amps_gpu_field_halo_set(self._h, peer, _ptr(a[0]), a[0].size, _ptr(a[1]), a[1].size, _ptr(a[2]), a[2].size,
                                                          _ptr(a[3]), a[3].size))
            self.field_primary = mask

    def E_upload(self, E):
        E = np.ascontiguousarray(E, dtype=np.float64)
        assert E.shape == (self.mesh.n_corners, 3)
        self._ck(self.lib.amps_gpu_E_upload(self._h, _ptr(E)))

    def field_step(self, theta=0.5, tol=1e-6, max_iter=200, restart=30, warm_start=False):
        it, rel = C.c_int(), C.c_double()
        self._ck(self.lib.amps_gpu_field_step(self._h, theta, tol, max_iter, restart, 1 if warm_start else 0, C.byref(it), C.byref(rel)))
        return int(it.value), float(rel.value)

    def fields_download(self, E=True, E_half=True, B=True, out=None):
        """-> dict of the requested fields; out = preallocated (pinned) arrays to fill instead"""
        res = out or {}
        if E and "E" not in res:
            res["E"] = np.empty((self.mesh.n_corners, 3))
        if E_half and "E_half" not in res:
            res["E_half"] = np.empty((self.mesh.n_corners, 3))
        if B and "B" not in res:
            res["B"] = np.empty((self.mesh.n_centers, 3))
        self._ck(self.lib.amps_gpu_fields_download(self._h, _ptr(res.get("E")) if E else None, _ptr(res.get("E_half")) if E_half else None,
                                                   _ptr(res.get("B")) if B else None))
        return res

    # ---- multi-GPU (one rank per GPU) -------------------------------------------------------
    def comm_init(self, dist):
        """Join the library's NCCL communicator; `dist` = torch.distributed (initialised) used only to broadcast the id
        and to gather the corner keys that define the J/M exchange lists."""
        import torch

        from . import mesh as meshmod

        rank, world = dist.get_rank(), dist.get_world_size()
        assert rank == self.mesh.rank and world == self.mesh.n_ranks
        idbuf = (C.c_ubyte * 128)()
        if rank == 0:
            rc = self.lib.amps_gpu_comm_unique_id(C.cast(idbuf, C.c_void_p))
            if rc != _capi.OK:
                raise AmpsGpuError("amps_gpu_comm_unique_id failed: libnccl.so.2 not loadable")
        obj = [bytes(idbuf) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        idbuf = (C.c_ubyte * 128).from_buffer_copy(obj[0])
        self._ck(self.lib.amps_gpu_comm_init(self._h, C.cast(idbuf, C.c_void_p), rank, world))
        gathered = [None] * world
        dist.all_gather_object(gathered, self.mesh.corner_target_gkeys)
        lists = meshmod.shared_corner_lists(self.mesh, gathered)
        for peer, uids in lists.items():
            uids = np.ascontiguousarray(uids, dtype=np.int32)
            self._ck(self.lib.amps_gpu_set_shared_corners(self._h, peer, _ptr(uids), uids.size))
        self.shared_lists = lists
        return lists

    def migrate(self):
        ns, nr = C.c_int64(), C.c_int64()
        self._ck(self.lib.amps_gpu_migrate(self._h, C.byref(ns), C.byref(nr)))
        return int(ns.value), int(nr.value)

    def exchange_JM(self):
        self._ck(self.lib.amps_gpu_exchange_JM(self._h))

    PHASES = ("move", "sort", "deposit", "exchange")

    def profile(self, enable=True):
        """returns {phase: (ms, count)} accumulated since the last call; (re)arms the event recording"""
        ms = (C.c_double * 4)()
        cnt = (C.c_int64 * 4)()
        self._ck(self.lib.amps_gpu_profile(self._h, 1 if enable else 0, C.cast(ms, C.c_void_p), C.cast(cnt, C.c_void_p)))
        return {p: (float(ms[i]), int(cnt[i])) for i, p in enumerate(self.PHASES)}

    def selftest_division(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        assert a.size == b.size
        bad = C.c_int64()
        self._ck(self.lib.amps_gpu_selftest_division(self._h, _ptr(a), _ptr(b), a.size, C.byref(bad)))
        return int(bad.value)

    def synchronize(self):
        self._ck(self.lib.amps_gpu_synchronize(self._h))

    def launch_count(self):
        return int(self.lib.amps_gpu_launch_count(self._h))

    def last_move_redo(self):
        n = C.c_int64()
        self._ck(self.lib.amps_gpu_last_move_redo(self._h, C.byref(n)))
        return int(n.value)

    def stream(self):
        return self.lib.amps_gpu_stream(self._h)
