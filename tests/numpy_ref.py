"""Independent vectorised numpy statement of the ECSIM push + deposit on a UNIFORM PERIODIC box.

Written from the equations (Lapenta 2017, JCP 334, App. D; doc/ecsim.tex of the reference) and the node
layout conventions only -- it shares no code or data structures with oracle/amps_oracle.cpp, so agreement
between the two is a second opinion on the restatement (see DESIGN.md "oracle pinning").
Fields live on global periodic lattices: E on corners [nx,ny,nz,3], B on centres [nx,ny,nz,3].
"""
import numpy as np

# cell-corner order used by ProcessCell: (0,0,0)(1,0,0)(1,1,0)(0,1,0)(0,0,1)(1,0,1)(1,1,1)(0,1,1)
CORNER = np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)])


def _nb_slot(d):
    # neighbour slot of the 27-stencil: offset 0,-1,+1 -> 0,1,2 (pic_field_solver_ecsim.cpp:625-637)
    return np.where(d == 0, 0, np.where(d < 0, 1, 2))


def _cross_term(B, beta):
    n = B.shape[0]
    P = -beta[:, None] * B
    T = np.zeros((n, 3, 3))
    T[:, 0, 1], T[:, 0, 2] = -P[:, 2], P[:, 1]
    T[:, 1, 0], T[:, 1, 2] = P[:, 2], -P[:, 0]
    T[:, 2, 0], T[:, 2, 1] = -P[:, 1], P[:, 0]
    return T


def trilinear_corner(x, n_cells, dx=1.0):
    """cell index and the 8 un-normalised corner weights (cell-corner order)"""
    xi = x / dx
    i0 = np.floor(xi).astype(np.int64)
    f = xi - i0
    W = np.empty((x.shape[0], 8))
    for c, (a, b, d) in enumerate(CORNER):
        W[:, c] = (f[:, 0] if a else 1 - f[:, 0]) * (f[:, 1] if b else 1 - f[:, 1]) * (f[:, 2] if d else 1 - f[:, 2])
    return i0, W


def gather_corner(F, i0, W, n_cells):
    out = np.zeros((i0.shape[0], 3))
    for c, off in enumerate(CORNER):
        idx = (i0 + off[None, :]) % np.asarray(n_cells)[None, :]
        out += W[:, c:c + 1] * F[idx[:, 0], idx[:, 1], idx[:, 2]]
    return out


def gather_center(F, x, n_cells, dx=1.0):
    xi = x / dx - 0.5
    i0 = np.floor(xi).astype(np.int64)
    f = xi - i0
    out = np.zeros((x.shape[0], 3))
    for a in (0, 1):
        for b in (0, 1):
            for d in (0, 1):
                w = (f[:, 0] if a else 1 - f[:, 0]) * (f[:, 1] if b else 1 - f[:, 1]) * (f[:, 2] if d else 1 - f[:, 2])
                idx = (i0 + np.array([a, b, d])[None, :]) % np.asarray(n_cells)[None, :]
                out += w[:, None] * F[idx[:, 0], idx[:, 1], idx[:, 2]]
    return out


def push(x, v, q_over_m, dt, E_corner, B_center, n_cells, dx=1.0):
    """x, v: [n,3]; returns x', v' (x' wrapped into the box)"""
    i0, W = trilinear_corner(x, n_cells, dx)
    E = gather_corner(E_corner, i0, W, n_cells)
    B = gather_center(B_center, x, n_cells, dx)
    beta = 0.5 * q_over_m * dt
    al = _alpha(B, beta)
    vt = v + beta[:, None] * E
    vp = np.einsum("nij,nj->ni", al, vt)
    vn = 2.0 * vp - v
    xn = x + dt * vn
    L = np.asarray(n_cells) * dx
    return np.mod(xn, L[None, :]), vn


def _alpha(B, beta):
    n = B.shape[0]
    c0 = 1.0 / (1.0 + beta ** 2 * (B ** 2).sum(1))
    BB = (beta ** 2)[:, None, None] * B[:, :, None] * B[:, None, :]
    return c0[:, None, None] * (np.eye(3)[None] + _cross_term(B, beta) + BB)


def deposit(x, v, qw, mw, dt, B_center, n_cells, dx=1.0):
    """J[nx,ny,nz,3], M[nx,ny,nz,27,9] on the periodic corner lattice; qw, mw = charge, mass times statistical weight"""
    n_cells = np.asarray(n_cells)
    V = dx ** 3
    i0, W = trilinear_corner(x, n_cells, dx)
    B = gather_center(B_center, x, n_cells, dx)
    beta = 0.5 * qw * dt / mw
    al = _alpha(B, beta)
    J = np.zeros(tuple(n_cells) + (3,))
    M = np.zeros(tuple(n_cells) + (27, 9))
    vrot = np.einsum("nij,nj->ni", al, v)
    k = qw * beta / V
    for c, offc in enumerate(CORNER):
        ic = (i0 + offc[None, :]) % n_cells[None, :]
        np.add.at(J, (ic[:, 0], ic[:, 1], ic[:, 2]), (qw * W[:, c] / V)[:, None] * vrot)
        for d, offd in enumerate(CORNER):
            delta = offd - offc
            slot = int(_nb_slot(delta[0]) + 3 * _nb_slot(delta[1]) + 9 * _nb_slot(delta[2]))
            val = (k * W[:, c] * W[:, d])[:, None] * al.reshape(-1, 9)
            np.add.at(M, (ic[:, 0], ic[:, 1], ic[:, 2], slot), val)
    return J, M
