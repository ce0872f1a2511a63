// mover_common.cuh -- device helpers shared by the mover translation units (both compiled --fmad=false):
// exact shared-reciprocal fp64 division, the flattened-tree search (a14) and the periodic wrap (a16).
#pragma once
#include "amps_dev.cuh"

namespace amps {

// ------------------------------------------------------------------------------------------------
// Correctly rounded fp64 quotients with a shared reciprocal.  An IEEE fp64 division costs ~25
// instructions; the mover needs ~30 of them per particle to round exactly like the CPU reference,
// and most divide by a per-block constant (dx, xmax-xmin, dx_max_refinement) or by one stencil norm.
// Given y = RN(1/b), two Newton corrections give a faithful quotient and Markstein's final step
//     r = fma(-q,b,a);  q' = fma(r,y,q)
// returns RN(a/b) exactly (Markstein 1990; Muller et al., Handbook of FP Arithmetic, thm 4.7-4.8).
// Excluded and routed to the plain division: b with an all-ones significand, tiny |a| (the exact
// residual could underflow).  tests/test_division_gpu.py checks bit equality against '/'.
// ------------------------------------------------------------------------------------------------
struct Recip {
  double y;
  bool ok;
};
__device__ __forceinline__ bool allones_significand(double b) {
  return (((unsigned long long)__double_as_longlong(b)) & 0x000FFFFFFFFFFFFFull) == 0x000FFFFFFFFFFFFFull;
}
// the IEEE division, kept out of line: it is the rarely taken fallback of ~30 call sites and inlining it
// made the mover's loop body overflow the instruction cache
static __device__ __noinline__ double div_slow(double a, double b) { return a / b; }
__device__ __forceinline__ Recip make_recip(double b) {
  Recip r;
  r.y = div_slow(1.0, b);
  r.ok = !allones_significand(b) && fabs(b) > 1e-150 && fabs(b) < 1e150;
  return r;
}
// exponent field outside [128,1792): zero, denormal, |a| < 2^-895, |a| >= 2^769, inf, nan
__device__ __forceinline__ bool exp_unsafe(double a) {
  const unsigned e = (((unsigned)__double2hiint(a)) >> 20) & 0x7ffu;
  return (e - 128u) >= 1664u;
}
// unguarded Markstein sequence: exact for y = RN(1/b), b without an all-ones significand, a "exp safe"
__device__ __forceinline__ double div_fast(double a, double b, double y) {
  double q = a * y;
  double r = fma(-q, b, a);
  q = fma(r, y, q);
  r = fma(-q, b, a);
  return fma(r, y, q);
}
__device__ __forceinline__ double div_rn(double a, double b, const Recip &rc) {
  if (!rc.ok || exp_unsafe(a)) return div_slow(a, b);
  return div_fast(a, b, rc.y);
}
// three quotients by per-block constants with one combined guard
__device__ __forceinline__ void div_rn3(const double (&a)[3], const double (&b)[3], const Recip (&rc)[3], bool allOk, double (&q)[3]) {
  if (allOk && !(exp_unsafe(a[0]) | exp_unsafe(a[1]) | exp_unsafe(a[2]))) {
#pragma unroll
    for (int d = 0; d < 3; d++) q[d] = div_fast(a[d], b[d], rc[d].y);
  } else {
#pragma unroll
    for (int d = 0; d < 3; d++) q[d] = div_slow(a[d], b[d]);
  }
}
// w[0..7] /= norm for a trilinear stencil whose weights sum to norm ~ 1 (a3/a4 Normalize()).
// norm == 1: identity.  norm == 1-2^-53 (all-ones significand): a/norm = a(1+2^-53+...) rounds to the
// next double above a.  1-4*2^-53 <= norm <= 1+8*2^-52: RN(1/norm) == 2-norm, then Markstein.  Anything
// else (stencils that lost cells at a domain boundary, zero weights) takes the IEEE division.
__device__ __forceinline__ void normalize8(double (&w)[8], double norm) {
  if (norm == 1.0 || !(norm > 0.0)) return;
  int hmin = __double2hiint(w[0]);
#pragma unroll
  for (int s = 1; s < 8; s++) hmin = min(hmin, __double2hiint(w[s]));  // weights are >= 0: hi words order like the values
  const bool safe = (norm >= 0x1.ffffffffffffcp-1) && (norm <= 0x1.0000000000008p+0) && (hmin >= 0x0C000000);
  if (safe) {
    if (norm == 0x1.fffffffffffffp-1) {
#pragma unroll
      for (int s = 0; s < 8; s++) w[s] = __longlong_as_double(__double_as_longlong(w[s]) + 1);
    } else {
      const double y = 2.0 - norm;
#pragma unroll
      for (int s = 0; s < 8; s++) w[s] = div_fast(w[s], norm, y);
    }
  } else {
#pragma unroll 1
    for (int s = 0; s < 8; s++) w[s] = div_slow(w[s], norm);
  }
}

// ------------------------------------------------------------------------------------------------
// a14: tree search.  findTreeNode(int*) walks up from the start node and down again
// (meshAMRgeneric.h:2793-2848); the result is the unique leaf containing the lattice point, so the
// device descends from the root grid directly.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_node_ix(const DevMesh &m, int ix0, int ix1, int ix2) {
  if (ix0 < 0 || ix1 < 0 || ix2 < 0) return -1;
  const int r0 = ix0 >> m.L, r1 = ix1 >> m.L, r2 = ix2 >> m.L;
  if (r0 >= m.nRoot[0] || r1 >= m.nRoot[1] || r2 >= m.nRoot[2]) return -1;
  int n = m.rootNode[r0 + m.nRoot[0] * (r1 + m.nRoot[1] * r2)];
  while (true) {
    const int h = m.isize[n] / 2;
    const int i = (ix0 - m.imin[3 * n] < h) ? 0 : 1;
    const int j = (ix1 - m.imin[3 * n + 1] < h) ? 0 : 1;
    const int k = (ix2 - m.imin[3 * n + 2] < h) ? 0 : 1;
    const int t = m.child[8 * n + i + 2 * (j + 2 * k)];
    if (t < 0) return n;
    n = t;
  }
}

// findTreeNode(double*), meshAMRgeneric.h:2851-2882.  Returns node id or -1.
__device__ __forceinline__ int find_tree_node(const DevMesh &m, const double x[3], const LeafGeo &start, const Recip (&rRef)[3], bool allOk) {
  int ix[3];
  {
    const double a[3] = {x[0] - m.xGlobalMin[0], x[1] - m.xGlobalMin[1], x[2] - m.xGlobalMin[2]};
    const double b[3] = {m.dxMaxRef[0], m.dxMaxRef[1], m.dxMaxRef[2]};
    double q[3];
    div_rn3(a, b, rRef, allOk, q);
#pragma unroll
    for (int d = 0; d < 3; d++) ix[d] = (int)floor(q[d]);
  }
  int node;
  const bool in = ix[0] >= start.imin[0] && ix[0] < start.imin[0] + start.isize && ix[1] >= start.imin[1] && ix[1] < start.imin[1] + start.isize &&
                  ix[2] >= start.imin[2] && ix[2] < start.imin[2] + start.isize;
  node = in ? start.node : find_node_ix(m, ix[0], ix[1], ix[2]);
  if (node >= 0) {
    bool flag = false;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const double lo = in ? start.xmin[d] : m.nxmin[3 * node + d];
      const double hi = in ? start.xmax[d] : m.nxmax[3 * node + d];
      if (x[d] < lo) ix[d]--, flag = true;
      if (x[d] >= hi) ix[d]++, flag = true;
    }
    if (flag) node = find_node_ix(m, ix[0], ix[1], ix[2]);
  }
  return node;
}

struct FaceGeo {  // PIC::Mover::cExternalBoundaryFace after Init (pic_mover.cpp:24-28, 48-75)
  double norm[6][3], e0[6][3], e1[6][3], x0[6][3], lE0[6], lE1[6];
};
__device__ __forceinline__ void init_faces(const DevMesh &m, FaceGeo &f) {
  const double nrm[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
  const int nX0[6][3] = {{0, 0, 0}, {1, 0, 0}, {0, 0, 0}, {0, 1, 0}, {0, 0, 0}, {0, 0, 1}};
  const double e0[6][3] = {{0, 1, 0}, {0, 1, 0}, {1, 0, 0}, {1, 0, 0}, {1, 0, 0}, {1, 0, 0}};
  const double e1[6][3] = {{0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 1, 0}, {0, 1, 0}};
  for (int n = 0; n < 6; n++) {
    double cE0 = 0.0, cE1 = 0.0;
    for (int d = 0; d < 3; d++) {
      f.norm[n][d] = nrm[n][d], f.e0[n][d] = e0[n][d], f.e1[n][d] = e1[n][d];
      f.x0[n][d] = (nX0[n][d] == 0) ? m.xGlobalMin[d] : m.xGlobalMax[d];
      const double a0 = ((e0[n][d] + nX0[n][d] < 0.5) ? m.xGlobalMin[d] : m.xGlobalMax[d]) - f.x0[n][d];
      const double a1 = ((e1[n][d] + nX0[n][d] < 0.5) ? m.xGlobalMin[d] : m.xGlobalMax[d]) - f.x0[n][d];
      cE0 += a0 * a0, cE1 += a1 * a1;  // pow(.,2)
    }
    f.lE0[n] = sqrt(cE0), f.lE1[n] = sqrt(cE1);
  }
}

__device__ __forceinline__ void add_exit_record(amps_gpu_exit_record *buf, unsigned long long *count, long long cap, int ptr, int spec, int face,
                                                int leaf, const double x[3], const double v[3]) {
  const unsigned long long i = atomicAdd(count, 1ull);
  if ((long long)i < cap) {
    amps_gpu_exit_record r;
    r.ptr = ptr, r.species = spec, r.face = face, r.leaf = leaf;
    for (int d = 0; d < 3; d++) r.x[d] = x[d], r.v[d] = v[d];
    buf[i] = r;
  }
}


// warp-reduce the per-thread counters of a mover and add them to the device statistics
__device__ __forceinline__ void flush_move_counters(DevMoveStats *stats, unsigned nMoved, unsigned nXCell, unsigned nXBlock, unsigned nLeft, unsigned nNotUsed,
                                                    unsigned nWrap, unsigned nErr) {
  unsigned int c[7] = {nMoved, nXCell, nXBlock, nLeft, nNotUsed, nWrap, nErr};
#pragma unroll
  for (int q = 0; q < 7; q++) {
    unsigned int v = c[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    c[q] = v;
  }
  if ((threadIdx.x & 31) == 0) {
    unsigned long long *s = reinterpret_cast<unsigned long long *>(stats);
#pragma unroll
    for (int q = 0; q < 7; q++)
      if (c[q]) atomicAdd(&s[q], (unsigned long long)c[q]);
  }
}

// tree search with plain IEEE divisions (exit handling and the test-particle movers)
__device__ __forceinline__ int find_tree_node_slow(const DevMesh &m, const double x[3]) {
  int ix[3];
#pragma unroll
  for (int d = 0; d < 3; d++) ix[d] = (int)floor(div_slow(x[d] - m.xGlobalMin[d], m.dxMaxRef[d]));
  int node = find_node_ix(m, ix[0], ix[1], ix[2]);
  if (node >= 0) {
    bool flag = false;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      if (x[d] < m.nxmin[3 * node + d]) ix[d]--, flag = true;
      if (x[d] >= m.nxmax[3 * node + d]) ix[d]++, flag = true;
    }
    if (flag) node = find_node_ix(m, ix[0], ix[1], ix[2]);
  }
  return node;
}

// a15: domain exit of the single-step movers (Boris :362-462, Lapenta2017 :1159-1265): mid-velocity ray against the six
// faces, EPS inset, clamp into the box.  Updates xInit/vInit (and copies them to xFinal/vFinal) like the reference.
// returns 0 = _PARTICLE_DELETED_ON_THE_FACE_ (user function: the caller records the exit), 1 = _PARTICLE_REJECTED_ON_THE_FACE_
// (specular: the reference then exit()s "not implemented"), -1 = the reference would exit()
static __device__ __noinline__ int domain_exit_vmiddle(const DevMesh &m, const FaceGeo &f, int boundaryMode, double dtTotal, double xInit[3], double vInit[3],
                                                double xFinal[3], double vFinal[3], int startNode, int *faceOut, int *nodeOut) {
  (void)startNode;
  int nIntersectionFace = -1;
  const double vMiddle[3] = {0.5 * (vInit[0] + vFinal[0]), 0.5 * (vInit[1] + vFinal[1]), 0.5 * (vInit[2] + vFinal[2])};
  double dtIntersection = -1.0;
  for (int nface = 0; nface < 6; nface++) {
    double cx = 0.0, cv = 0.0, r0[3];
    for (int d = 0; d < 3; d++) {
      r0[d] = xInit[d] - f.x0[nface][d];
      cx += r0[d] * f.norm[nface][d];
      cv += vMiddle[d] * f.norm[nface][d];
    }
    if (cv > 0.0) {
      const double dt = -cx / cv;
      if ((dtIntersection < 0.0) || ((dt < dtIntersection) && (dt > 0.0))) {
        double cE0 = 0.0, cE1 = 0.0;
        for (int d = 0; d < 3; d++) {
          const double c = r0[d] + dt * vMiddle[d];
          cE0 += c * f.e0[nface][d], cE1 += c * f.e1[nface][d];
        }
        if ((cE0 < -m.eps) || (cE0 > f.lE0[nface] + m.eps) || (cE1 < -m.eps) || (cE1 > f.lE1[nface] + m.eps)) continue;
        nIntersectionFace = nface, dtIntersection = dt;
      }
    }
  }
  if (nIntersectionFace == -1) return -1;
  const double tVelocityIncrement = ((dtIntersection / dtTotal < 1) ? dtIntersection / dtTotal : 1);
  for (int d = 0; d < 3; d++) {
    xInit[d] += dtIntersection * vMiddle[d] - f.norm[nIntersectionFace][d] * m.eps;
    vInit[d] += tVelocityIncrement * (vFinal[d] - vInit[d]);
  }
  int newNode = find_tree_node_slow(m, xInit);
  if (newNode < 0) {
    for (int d = 0; d < 3; d++) {
      if (m.xGlobalMin[d] >= xInit[d]) xInit[d] = m.xGlobalMin[d] + m.eps;
      if (m.xGlobalMax[d] <= xInit[d]) xInit[d] = m.xGlobalMax[d] - m.eps;
    }
    newNode = find_tree_node_slow(m, xInit);
    if (newNode < 0) return -1;
  }
  int code;
  if (boundaryMode == AMPS_BOUNDARY_USER_FUNCTION) code = 0;
  else if (boundaryMode == AMPS_BOUNDARY_SPECULAR_REFLECTION) {
    double cc = 0.0;
    for (int d = 0; d < 3; d++) cc += f.norm[nIntersectionFace][d] * vInit[d];
    for (int d = 0; d < 3; d++) vInit[d] -= 2.0 * cc * f.norm[nIntersectionFace][d];
    code = 1;
  } else
    return -1;
  for (int d = 0; d < 3; d++) vFinal[d] = vInit[d], xFinal[d] = xInit[d];
  *faceOut = nIntersectionFace, *nodeOut = newNode;
  return code;
}

}  // namespace amps
