"""ctypes driver of oracle/_ref/libref_pic.so: the reference's own PIC core (fast-wave configuration, one rank) as a checker.
Test infrastructure (see oracle/ref_pic/ref_pic_shim.cpp); only tests/ may use it."""
import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.environ.get("AMPS_REF_PIC_LIB") or os.path.join(os.path.dirname(HERE), "_ref", "libref_pic.so")


def available():
    return os.path.exists(LIB)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class quiet:
    """the reference prints its progress to stdout: send fd 1 to /dev/null while it runs"""

    def __enter__(self):
        import sys

        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        C.CDLL(None).fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


class RefPic:
    """One process can initialise the reference once (its state is global)."""

    def __init__(self):
        self.lib = C.CDLL(LIB)
        self.lib.ref_pic_corner_rw.restype = C.c_long
        self.lib.ref_pic_center_rw.restype = C.c_long
        self.lib.ref_pic_particles.restype = C.c_long
        cwd = os.getcwd()
        self.tmp = tempfile.mkdtemp(prefix="ref_pic_")  # the reference writes its volume table into the working directory
        os.chdir(self.tmp)
        try:
            with quiet():
                self.n_blocks = self.lib.ref_pic_init()
        finally:
            os.chdir(cwd)
        d = np.zeros(10, dtype=np.int64)
        self.lib.ref_pic_dims(_p(d))
        self.N, self.g = tuple(int(v) for v in d[0:3]), tuple(int(v) for v in d[3:6])
        self.n_species, self.n_particles, self.record_len = int(d[7]), int(d[8]), int(d[9])
        c = np.zeros(3 * self.n_species + 16)
        self.lib.ref_pic_constants(_p(c))
        ns = self.n_species
        self.charge_si, self.mass_si, self.weight = c[0:ns].copy(), c[ns:2 * ns].copy(), c[2 * ns:3 * ns].copy()
        k = 3 * ns
        self.dt, self.light_speed, self.B_conv, self.E_conv, self.length_conv = c[k], c[k + 1], c[k + 2], c[k + 3], c[k + 4]
        self.charge_conv, self.mass_conv = c[k + 5], c[k + 6]
        q, mm = np.zeros(ns), np.zeros(ns)
        self.lib.ref_pic_species(_p(q), _p(mm))
        self.charge, self.mass = q, mm
        self.bxmin, self.bxmax = np.zeros((self.n_blocks, 3)), np.zeros((self.n_blocks, 3))
        self.ghost = np.zeros(self.n_blocks, dtype=np.int32)
        self.lib.ref_pic_blocks(_p(self.bxmin), _p(self.bxmax), _p(self.ghost))

    # node arrays [block][k][j][i][len] incl. the ghost layers
    def corner_shape(self, length):
        N, g = self.N, self.g
        return (self.n_blocks, N[2] + 2 * g[2] + 1, N[1] + 2 * g[1] + 1, N[0] + 2 * g[0] + 1, length)

    def center_shape(self):
        N, g = self.N, self.g
        return (self.n_blocks, N[2] + 2 * g[2], N[1] + 2 * g[1], N[0] + 2 * g[0], 3)

    def corner(self, what):
        a = np.empty(self.corner_shape({0: 3, 1: 3, 2: 3, 3: 243}[what]))
        self.lib.ref_pic_corner_rw(what, _p(a), 0)
        return a

    def set_corner(self, what, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self.corner_shape({0: 3, 1: 3, 2: 3, 3: 243}[what])
        self.lib.ref_pic_corner_rw(what, _p(a), 1)

    def center(self, what):
        a = np.empty(self.center_shape())
        self.lib.ref_pic_center_rw(what, _p(a), 0)
        return a

    def set_center(self, what, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self.center_shape()
        self.lib.ref_pic_center_rw(what, _p(a), 1)

    def corner_positions(self):
        """x[block][k][j][i][3] of every corner node incl. ghost layers"""
        N, g = self.N, self.g
        dx = (self.bxmax - self.bxmin) / np.array(N)
        out = np.empty(self.corner_shape(3))
        for d, (n, gg) in enumerate(zip(N, g)):
            idx = np.arange(-gg, n + gg + 1, dtype=np.float64)
            shp = [1, 1, 1, 1]
            shp[3 - d] = idx.size
            out[..., d] = self.bxmin[:, d].reshape(-1, 1, 1, 1) + idx.reshape(shp) * dx[:, d].reshape(-1, 1, 1, 1)
        return out

    def center_positions(self):
        N, g = self.N, self.g
        dx = (self.bxmax - self.bxmin) / np.array(N)
        out = np.empty(self.center_shape())
        for d, (n, gg) in enumerate(zip(N, g)):
            idx = np.arange(-gg, n + gg, dtype=np.float64) + 0.5
            shp = [1, 1, 1, 1]
            shp[3 - d] = idx.size
            out[..., d] = self.bxmin[:, d].reshape(-1, 1, 1, 1) + idx.reshape(shp) * dx[:, d].reshape(-1, 1, 1, 1)
        return out

    def particles(self):
        n = self.n_particles + 1024
        ptr = np.zeros(n, dtype=np.int64)
        x, v, w = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n)
        spec, block, cell = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        k = self.lib.ref_pic_particles(n, _p(ptr), _p(x), _p(v), _p(w), _p(spec), _p(block), _p(cell))
        assert k <= n
        return {"ptr": ptr[:k], "x": x[:k].T.copy(), "v": v[:k].T.copy(), "w": w[:k], "species": spec[:k], "block": block[:k], "cell": cell[:k]}

    def thin(self, keep_every):
        self.lib.ref_pic_thin.restype = C.c_long
        self.n_particles = int(self.lib.ref_pic_thin(int(keep_every)))
        return self.n_particles

    def set_weight_correction(self, ptr, w):
        ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        self.lib.ref_pic_set_weight_correction(C.c_long(ptr.size), _p(ptr), _p(w))

    # ---- the gyrokinetic variant (oracle/_ref/libref_pic_gk.so: REF_PIC_VARIANT=gk of build_ref_pic.sh) ----
    def gyrokinetic(self):
        return hasattr(self.lib, "ref_pic_gyrokinetic") and self.lib.ref_pic_gyrokinetic() == 1

    def set_gc_species(self, spec, on=True):
        self.lib.ref_pic_set_gc_species(int(spec), 1 if on else 0)

    def set_mover_mode(self, mode):
        """0: PIC::GYROKINETIC::Mover (first-order guiding centre / Lapenta2017), 1: second-order guiding centre / Lapenta2017"""
        self.lib.ref_pic_set_mover_mode(int(mode))

    def set_reduced(self, ptr, mu=None, vnormal=None, init_flag=None):
        ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        mu = None if mu is None else np.ascontiguousarray(mu, dtype=np.float64)
        vn = None if vnormal is None else np.ascontiguousarray(vnormal, dtype=np.float64)
        fl = None if init_flag is None else np.ascontiguousarray(init_flag, dtype=np.int32)
        self.lib.ref_pic_set_reduced(C.c_long(ptr.size), _p(ptr), None if mu is None else _p(mu), None if vn is None else _p(vn),
                                     None if fl is None else _p(fl))

    def get_reduced(self, ptr):
        ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        mu, vn, fl = np.zeros(ptr.size), np.zeros(ptr.size), np.zeros(ptr.size, dtype=np.int32)
        self.lib.ref_pic_get_reduced(C.c_long(ptr.size), _p(ptr), _p(mu), _p(vn), _p(fl))
        return mu, vn, fl

    def sample_cells(self):
        """PIC::Sampling::SamplingManager once more -> (collecting buffer [block][cell][species][10], particles sampled per species)"""
        out = np.zeros((self.n_blocks, self.N[0] * self.N[1] * self.N[2], self.n_species, 10))
        cnt = np.zeros(self.n_species, dtype=np.int64)
        with quiet():
            self.lib.ref_pic_sample_cells(_p(out), _p(cnt))
        return out, cnt

    # ---- the particle passes of ECSIM::divECorrection ----
    def center_scalar_shape(self):
        return self.center_shape()[:-1]

    def center_scalar(self, index):
        """one double of the ECSIM centre data, [block][k][j][i] incl. ghost cells: 6 netChargeOld, 7 netChargeNew, 8 divE, 9 phi"""
        a = np.empty(self.center_scalar_shape())
        self.lib.ref_pic_center_scalar_rw(int(index), _p(a), 0)
        return a

    def set_center_scalar(self, index, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self.center_scalar_shape()
        self.lib.ref_pic_center_scalar_rw(int(index), _p(a), 1)

    def compute_net_charge(self, update_old=False):
        with quiet():
            self.lib.ref_pic_compute_net_charge(1 if update_old else 0)
        return self.center_scalar(7)

    def samples_species_on_corners(self):
        return hasattr(self.lib, "ref_pic_samples_species_on_corners") and self.lib.ref_pic_samples_species_on_corners() == 1

    def species_moments(self):
        """[block][k][j][i][species][10] on every corner node incl. ghost layers (after update_JM)"""
        a = np.empty(self.corner_shape(10 * self.n_species))
        self.lib.ref_pic_species_moments(_p(a))
        return a.reshape(a.shape[:-1] + (self.n_species, 10))

    def correct_particle_location(self):
        with quiet():
            self.lib.ref_pic_correct_particle_location()

    def move(self):
        with quiet():
            self.lib.ref_pic_move()

    def update_JM(self):
        e = C.c_double()
        with quiet():
            self.lib.ref_pic_update_JM(C.byref(e))
        return float(e.value)
