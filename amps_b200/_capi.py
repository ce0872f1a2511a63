"""ctypes mirror of include/amps_gpu.h (plain C structs + function prototypes).

The shared library is built in-tree by ``__graft_entry__.build()`` as
``amps_b200/libamps_gpu.so``.  There is no CPU fallback: loading fails loudly
when the library is missing, and ``amps_gpu_init`` fails loudly without a GPU.
"""
import ctypes as C
import os

MAX_SPECIES = 8

# status codes
OK, ERR_CUDA, ERR_ARG, ERR_CAPACITY, ERR_STATE, ERR_PARTICLE, ERR_NO_DEVICE = range(7)

# movers
MOVER_LAPENTA2017 = 0
MOVER_BORIS = 1
MOVER_RELATIVISTIC_BORIS = 2
MOVER_GC_FIRST_ORDER = 3
MOVER_GC_SECOND_ORDER = 4
MOVER_RELATIVISTIC_GCA = 5
MOVER_MARKIDIS2010 = 6
MOVER_GYROKINETIC_FIRST_ORDER = 7
MOVER_GYROKINETIC_SECOND_ORDER = 8

PARTICLE_LEFT_THE_DOMAIN = 2
PARTICLE_MOTION_FINISHED = 3
PARTICLE_IN_NOT_IN_USE_NODE = 4

BOUNDARY_DELETE, BOUNDARY_SPECULAR_REFLECTION, BOUNDARY_USER_FUNCTION = 0, 1, 2
B_CENTER_BASED, B_CORNER_BASED = 0, 1
DT_SINGLE_GLOBAL, DT_SPECIES_GLOBAL = 0, 1
NODE_USED, NODE_PERIODIC_GHOST = 1, 2

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


class Config(C.Structure):
    _fields_ = [
        ("block_cells", C.c_int32 * 3),
        ("ghost_cells", C.c_int32 * 3),
        ("n_species", C.c_int32),
        ("b_mode", C.c_int32),
        ("periodic", C.c_int32),
        ("boundary_mode", C.c_int32),
        ("time_step_mode", C.c_int32),
        ("device", C.c_int32),
        ("capacity", C.c_int64),
        ("charge", C.c_double * MAX_SPECIES),
        ("mass", C.c_double * MAX_SPECIES),
        ("species_weight", C.c_double * MAX_SPECIES),
        ("time_step", C.c_double * MAX_SPECIES),
        ("ecsim_dt_total", C.c_double),
        ("ecsim_B_conv", C.c_double),
        ("ecsim_length_conv", C.c_double),
        ("ecsim_light_speed", C.c_double),
        ("coupler_interpolation", C.c_int32),
        ("backward_time_integration", C.c_int32),
        ("speed_of_light", C.c_double),
        ("internal_sphere_radius", C.c_double),
        ("exit_record_capacity", C.c_int64),
        ("gravity_gm", C.c_double),
        ("carry_magnetic_moment", C.c_int32),
        ("exact_arithmetic", C.c_int32),
        ("carry_v_parallel", C.c_int32),
        ("ideal_mhd", C.c_int32),
        ("gc_species_mask", C.c_int32),
        ("gc_fields_ecsim", C.c_int32),
    ]


class Mesh(C.Structure):
    _fields_ = [
        ("n_root", C.c_int32 * 3),
        ("max_refinement_level", C.c_int32),
        ("x_global_min", C.c_double * 3),
        ("x_global_max", C.c_double * 3),
        ("dx_max_refinement", C.c_double * 3),
        ("dx_root_block", C.c_double * 3),
        ("eps", C.c_double),
        ("n_nodes", C.c_int32),
        ("node_parent", _i32p),
        ("node_child", _i32p),
        ("node_level", _i32p),
        ("node_imin", _i32p),
        ("node_isize", _i32p),
        ("node_xmin", _f64p),
        ("node_xmax", _f64p),
        ("node_leaf", _i32p),
        ("node_flags", _i32p),
        ("node_thread", _i32p),
        ("root_node", _i32p),
        ("n_leaves", C.c_int32),
        ("leaf_node", _i32p),
        ("leaf_real", _i32p),
        ("leaf_face_boundary", _i32p),
        ("n_corners", C.c_int32),
        ("n_centers", C.c_int32),
        ("leaf_corner_uid", _i32p),
        ("leaf_center_uid", _i32p),
        ("this_rank", C.c_int32),
        ("n_ranks", C.c_int32),
        ("n_global_leaves", C.c_int32),
        ("leaf_owner", _i32p),
        ("leaf_global_id", _i32p),
        ("global_leaf_to_local", _i32p),
    ]


class AosLayout(C.Structure):
    _fields_ = [
        ("stride", C.c_int64),
        ("off_species", C.c_int32),
        ("off_v", C.c_int32),
        ("off_x", C.c_int32),
        ("off_w", C.c_int32),
        ("off_mu", C.c_int32),
        ("off_next", C.c_int32),
        ("off_prev", C.c_int32),
        ("off_vpar", C.c_int32),
    ]


class ExitRecord(C.Structure):
    _fields_ = [("ptr", C.c_int32), ("species", C.c_int32), ("face", C.c_int32), ("leaf", C.c_int32), ("x", C.c_double * 3), ("v", C.c_double * 3)]


CPLR_CONSTANT, CPLR_LINEAR = 0, 1
EXIT_SPHERE = 6


class MoveStats(C.Structure):
    _fields_ = [
        ("n_moved", C.c_int64),
        ("n_cross_cell", C.c_int64),
        ("n_cross_block", C.c_int64),
        ("n_left_domain", C.c_int64),
        ("n_not_in_use", C.c_int64),
        ("n_periodic_wrap", C.c_int64),
        ("n_error", C.c_int64),
        ("n_sub_steps", C.c_int64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


# every symbol include/amps_gpu.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
PROTOTYPES = {
    "amps_gpu_init": (C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    "amps_gpu_finalize": (C.c_int, [_vp]),
    "amps_gpu_last_error": (C.c_char_p, [_vp]),
    "amps_gpu_launch_count": (C.c_int64, [_vp]),
    "amps_gpu_last_move_redo": (C.c_int, [_vp, _i64p]),
    "amps_gpu_stream": (_vp, [_vp]),
    "amps_gpu_mesh_upload": (C.c_int, [_vp, C.POINTER(Mesh)]),
    "amps_gpu_fields_upload": (C.c_int, [_vp, _vp, _vp, _vp]),
    "amps_gpu_background_upload": (C.c_int, [_vp, _vp, _vp]),
    "amps_gpu_exit_records": (C.c_int, [_vp, _vp, C.c_int64, _i64p]),
    "amps_gpu_background_upload_gca": (C.c_int, [_vp, _vp]),
    "amps_gpu_magnetic_moment_init": (C.c_int, [_vp, C.c_int]),
    "amps_gpu_background_upload_gradB": (C.c_int, [_vp, _vp]),
    "amps_gpu_magnetic_moment_upload": (C.c_int, [_vp, _vp, C.c_int64]),
    "amps_gpu_magnetic_moment_download": (C.c_int, [_vp, _vp, C.c_int64, _i64p]),
    "amps_gpu_particles_upload_aos": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, C.POINTER(AosLayout)]),
    "amps_gpu_particles_upload_soa": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64]),
    "amps_gpu_particles_append_soa": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64]),
    "amps_gpu_particle_count": (C.c_int, [_vp, _i64p]),
    "amps_gpu_particles_download_soa": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, _i64p]),
    "amps_gpu_particles_download_aos": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.POINTER(AosLayout), _i64p]),
    "amps_gpu_field_solver_init": (C.c_int, [_vp, _vp, _vp, _vp]),
    "amps_gpu_E_upload": (C.c_int, [_vp, _vp]),
    "amps_gpu_field_halo_set": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int64]),
    "amps_gpu_field_primary_set": (C.c_int, [_vp, _vp]),
    "amps_gpu_field_step": (C.c_int, [_vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "amps_gpu_fields_download": (C.c_int, [_vp, _vp, _vp, _vp]),
    "amps_gpu_comm_uses_peer_memory": (C.c_int, [_vp]),
    "amps_gpu_restart_save": (C.c_int, [_vp, C.c_char_p, _vp, C.c_int64, _vp, C.c_int32, C.POINTER(AosLayout), _i64p]),
    "amps_gpu_restart_read": (C.c_int, [_vp, C.c_char_p, C.c_int64, _vp, C.c_int32, C.POINTER(AosLayout), _i64p]),
    "amps_gpu_particles_slot_delta": (C.c_int, [_vp, _i64p, _vp, C.c_int64, _i64p]),
    "amps_gpu_particles_assign_slots": (C.c_int, [_vp, _vp, C.c_int64]),
    "amps_gpu_cell_table_download": (C.c_int, [_vp, _vp, C.c_int64]),
    "amps_gpu_sort": (C.c_int, [_vp]),
    "amps_gpu_move": (C.c_int, [_vp, C.c_int, C.POINTER(MoveStats)]),
    "amps_gpu_deposit_JM": (C.c_int, [_vp, _vp, _vp]),
    "amps_gpu_diagnostics": (C.c_int, [_vp, _vp, _vp]),
    "amps_gpu_v_parallel_upload": (C.c_int, [_vp, _vp, C.c_int64]),
    "amps_gpu_v_normal_upload": (C.c_int, [_vp, _vp, C.c_int64]),
    "amps_gpu_global_stencil_set": (C.c_int, [_vp, C.c_int32]),
    "amps_gpu_v_parallel_download": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "amps_gpu_net_charge": (C.c_int, [_vp, C.c_double, _vp]),
    "amps_gpu_JM_packed_slots": (C.POINTER(C.c_int32), []),
    "amps_gpu_JM_download_packed": (C.c_int, [_vp, _vp]),
    "amps_gpu_step_JM_packed": (C.c_int, [_vp, C.c_int, _vp]),
    "amps_gpu_sample_cells": (C.c_int, [_vp]),
    "amps_gpu_sample_download": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "amps_gpu_species_moments": (C.c_int, [_vp, _vp]),
    "amps_gpu_phi_upload": (C.c_int, [_vp, _vp]),
    "amps_gpu_correct_particle_location": (C.c_int, [_vp, C.c_double, C.c_double, _vp, _vp]),
    "amps_gpu_step_JM": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "amps_gpu_JM_download": (C.c_int, [_vp, _vp, _vp]),
    "amps_gpu_JM_device": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "amps_gpu_comm_unique_id": (C.c_int, [_vp]),
    "amps_gpu_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "amps_gpu_set_shared_corners": (C.c_int, [_vp, C.c_int, _vp, C.c_int64]),
    "amps_gpu_migrate": (C.c_int, [_vp, _i64p, _i64p]),
    "amps_gpu_exchange_JM": (C.c_int, [_vp]),
    "amps_gpu_step": (C.c_int, [_vp, C.c_int]),
    "amps_gpu_profile": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "amps_gpu_selftest_division": (C.c_int, [_vp, _vp, _vp, C.c_int64, _i64p]),
    "amps_gpu_synchronize": (C.c_int, [_vp]),
}

# AMPS_GPU_LIB selects another build of the same library (A/B runs of kernel variants); there is still no fallback
LIB_PATH = os.environ.get("AMPS_GPU_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libamps_gpu.so")
_lib = None


def load_library(path=None):
    """dlopen the C-ABI library and bind every prototype. Raises if missing (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "amps_b200 has no CPU fallback.")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
