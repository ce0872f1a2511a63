"""CPU restatement (numpy) of the field half of the reference's ECSIM step -- TEST INFRASTRUCTURE, never on the product path.

  ECSIM::TimeStep                src/pic/pic_field_solver_ecsim.cpp:6004-6157 (:6449-6560 in the generated tree)
  GetStencil (row of the system) src/pic/ecsim/get_stencil.cpp         identity + (theta c dt)^2 (grad div - laplace) + 4 pi theta dt M
  InitDiscritizationStencil      src/pic/pic_field_solver_ecsim.cpp:7451  the compact 27-node tables
  UpdateRhs / SampleRhsScalar    src/pic/ecsim/update_rhs.cpp
  UpdateMatrixElement            :6271 (parameter + 4 pi dt theta * mass-matrix entry)
  ProcessFinalSolution           :6581   E^{n+theta} = x / E_conv + E^n
  UpdateB                        :5388   B^{n+1} = B^n - c dt curl E^{n+theta}, corner E averaged over the 4 edges of a face pair
  UpdateE                        :6170   E^{n+1} = (E^{n+theta} - (1 - theta) E^n) / theta

All fields live on the unique nodes of amps_b200.mesh (single-level mesh); normalised units (E_conv = B_conv = 1,
_PIC_FIELD_SOLVER_INPUT_UNIT_NORM_), the 81-element compact stencil (_PIC_STENCIL_NUMBER_ 81, corrCoeff = 0: the defaults).
Pinned against the reference's own code run here (oracle/_ref/libref_pic.so): tests/test_reference_field_solve.py compares the
tables below with the reference's LaplacianStencil / GradDivStencil, the right-hand side and the operator with UpdateRhs / matvec
row by row, and E^{n+theta}, E^{n+1}, B^{n+1} with ECSIM::TimeStep."""
import numpy as np

D2 = np.array([1.0, -2.0, 1.0])    # second difference
AV = np.array([0.25, 0.5, 0.25])   # the transverse average of the compact stencils
D1 = np.array([-0.5, 0.0, 0.5])    # central first difference


def graddiv_table(p, q):
    """GradDivStencil[p][q] (= LaplacianStencil[p] for p == q) as a [3,3,3] array indexed by offset + 1"""
    v = [AV, AV, AV]
    if p == q:
        v[p] = D2
    else:
        v[p], v[q] = D1, D1
    return np.einsum("i,j,k->ijk", *v)


def slot(d):
    """neighbour slot code of one dimension: 0 -> 0, -1 -> 1, +1 -> 2 (indexAddition = {0,-1,1})"""
    return (3 * d * d + d) >> 1


def operator_constants(dx, c_light, dt, theta):
    """K[p][q][slot]: the parameter part of the matrix row p (GetStencil): identity, minus the Laplacian, plus grad-div"""
    coeff = c_light * dt / np.asarray(dx, dtype=np.float64) * theta
    K = np.zeros((3, 3, 27))
    for p in range(3):
        for q in range(3):
            G = graddiv_table(p, q)
            for a in (-1, 0, 1):
                for b in (-1, 0, 1):
                    for c in (-1, 0, 1):
                        s = slot(a) + 3 * slot(b) + 9 * slot(c)
                        K[p, q, s] += coeff[p] * coeff[q] * G[a + 1, b + 1, c + 1]
                        if p == q:
                            for e in range(3):
                                K[p, p, s] -= coeff[e] ** 2 * graddiv_table(e, e)[a + 1, b + 1, c + 1]
        K[p, p, 0] += 1.0
    return K, coeff


class EcsimField:
    def __init__(self, mesh, dx, c_light, dt, theta=0.5):
        self.nb, self.cc, self.zc = (t.astype(np.int64) for t in mesh.field_solver_tables())
        assert (self.nb >= 0).all() and (self.cc >= 0).all() and (self.zc >= 0).all(), "periodic single-level meshes only"
        self.K, self.coeff = operator_constants(dx, c_light, dt, theta)
        self.dx, self.c, self.dt, self.theta = np.asarray(dx, dtype=np.float64), c_light, dt, theta
        self.f = 4.0 * np.pi * dt * theta

    def matvec(self, x, M):
        """y = A x, x[n_corners,3]; M[n_corners,243] with entry 9 slot + 3 p + q (MassMatrixOffsetTable, :678-699)"""
        y = np.zeros_like(x)
        for s in range(27):
            xn = x[self.nb[:, s]]
            for p in range(3):
                for q in range(3):
                    y[:, p] += (self.K[p, q, s] + self.f * M[:, 9 * s + 3 * p + q]) * xn[:, q]
        return y

    def rhs(self, E, B, J, M):
        """UpdateRhs: -(operator - identity) E^n - 4 pi dt theta (J + M E^n) + theta c dt curl B^n (2x2 face averages of centre B)"""
        R = np.zeros_like(E)
        for s in range(27):
            En = E[self.nb[:, s]]
            for p in range(3):
                for q in range(3):
                    k = self.K[p, q, s] - (1.0 if (p == q and s == 0) else 0.0)
                    R[:, p] -= (k + self.f * M[:, 9 * s + 3 * p + q]) * En[:, q]
        R -= self.f * J
        c4 = 0.25 * self.coeff
        cell = lambda a, b, c: self.cc[:, (a + 1) + 2 * (b + 1) + 4 * (c + 1)]  # noqa: E731
        for a in (-1, 0):
            for b in (-1, 0):
                R[:, 0] += c4[1] * (B[cell(a, 0, b), 2] - B[cell(a, -1, b), 2]) - c4[2] * (B[cell(a, b, 0), 1] - B[cell(a, b, -1), 1])
                R[:, 1] += c4[2] * (B[cell(a, b, 0), 0] - B[cell(a, b, -1), 0]) - c4[0] * (B[cell(0, b, a), 2] - B[cell(-1, b, a), 2])
                R[:, 2] += c4[0] * (B[cell(0, b, a), 1] - B[cell(-1, b, a), 1]) - c4[1] * (B[cell(b, 0, a), 0] - B[cell(b, -1, a), 0])
        return R

    def solve(self, R, M, tol=1e-10, max_iter=300, restart=100):
        """restarted GMRES from x0 = 0 on the relative residual (what linear_solver_wrapper("GMRES", ...) is asked for, :3282)"""
        import scipy.sparse.linalg as sla

        n = R.size
        A = sla.LinearOperator((n, n), matvec=lambda v: self.matvec(v.reshape(-1, 3), M).reshape(-1))
        its = [0]
        x, info = sla.gmres(A, R.reshape(-1), rtol=tol, atol=0.0, restart=restart, maxiter=max(1, max_iter // restart + 1),
                            callback=lambda r: its.__setitem__(0, its[0] + 1), callback_type="pr_norm")
        return x.reshape(-1, 3), its[0], info

    def update_B(self, B, Eh):
        cdt4 = 0.25 * self.c * self.dt / self.dx
        Ec = {(a, b, c): Eh[self.zc[:, a + 2 * b + 4 * c]] for a in (0, 1) for b in (0, 1) for c in (0, 1)}
        t = np.zeros_like(B)
        for a in (0, 1):
            for b in (0, 1):
                t[:, 0] += -cdt4[1] * (Ec[(a, 1, b)][:, 2] - Ec[(a, 0, b)][:, 2]) + cdt4[2] * (Ec[(a, b, 1)][:, 1] - Ec[(a, b, 0)][:, 1])
                t[:, 1] += -cdt4[2] * (Ec[(a, b, 1)][:, 0] - Ec[(a, b, 0)][:, 0]) + cdt4[0] * (Ec[(1, a, b)][:, 2] - Ec[(0, a, b)][:, 2])
                t[:, 2] += -cdt4[0] * (Ec[(1, a, b)][:, 1] - Ec[(0, a, b)][:, 1]) + cdt4[1] * (Ec[(a, 1, b)][:, 0] - Ec[(a, 0, b)][:, 0])
        return B + t

    def update_E(self, E, Eh):
        return (Eh - (1.0 - self.theta) * E) / self.theta

    def step(self, E, B, J, M, tol=1e-10, max_iter=300):
        """-> E^{n+theta}, E^{n+1}, B^{n+1}, iterations"""
        R = self.rhs(E, B, J, M)
        x, its, _ = self.solve(R, M, tol, max_iter)
        Eh = E + x
        return Eh, self.update_E(E, Eh), self.update_B(B, Eh), its
