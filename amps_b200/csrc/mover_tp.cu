// mover_tp.cu -- test-particle movers on the coupler's centre-node table (sm_100a).  COMPILED WITH --fmad=false
// so that positions, velocities and the (block,cell) assignment round exactly like the reference CPU build.
//
//   a7  PIC::Mover::Relativistic::Boris   src/pic/pic_mover_relativistic_boris.cpp:16-583
//       gamma-aware Boris rotation sub-cycled with the local gyro period (dt = min(dtLeft, 1/f_g), :115-125),
//       backward-time mode (:128-132, :169-173), internal sphere (:270-302), domain exit search (:305-455)
//   a15 domain exit: DELETE, or the face-intersection search whose result is handed to the host callbacks as an
//       exit record (USER_FUNCTION = Earth::CutoffRigidity::ProcessOutsideDomainParticles); SPECULAR is an error
//       in this mover exactly as in the reference (:452-453 "not implemented")
//   fields: PIC::CPLR::InitInterpolationStencil (pic_swmf.cpp:76-90, cell-centred constant | linear) +
//       GetBackgroundElectricField/MagneticField (pic.h:8338-8425) on per-leaf centre tiles [nCenterLocal][6]
//
// Mapping: thread <-> particle.  The trip count of the sub-cycle loop varies per particle and a particle may
// cross several blocks inside one call, so the field tiles are read from global memory (L2 resident) instead of
// being staged per block.
#include "amps_dev.cuh"
#include "mover_common.cuh"

namespace amps {

__global__ void stage_background_kernel(DevMesh m, const double *__restrict__ E, const double *__restrict__ B, double *__restrict__ tile) {
  const int leaf = blockIdx.x;
  const int *cuid = m.centerUid + (size_t)leaf * m.nCenterLocal;
  double *dst = tile + (size_t)leaf * m.nCenterLocal * 6;
  for (int i = threadIdx.x; i < m.nCenterLocal; i += blockDim.x) {
    const int u = cuid[i];
    for (int d = 0; d < 3; d++) {
      if (E) dst[6 * i + d] = (u >= 0) ? E[3 * (size_t)u + d] : 0.0;
      if (B) dst[6 * i + 3 + d] = (u >= 0) ? B[3 * (size_t)u + d] : 0.0;
    }
  }
}
void launch_stage_background(const DevMesh &m, const double *E, const double *B, double *tile, cudaStream_t s) {
  stage_background_kernel<<<m.nLeaves, 256, 0, s>>>(m, E, B, tile);
}

// plain-division variant of the tree search (this mover is not bound by the division count)
__device__ __forceinline__ int find_tree_node_plain(const DevMesh &m, const double x[3], int startNode) {
  int ix[3];
#pragma unroll
  for (int d = 0; d < 3; d++) ix[d] = (int)floor((x[d] - m.xGlobalMin[d]) / m.dxMaxRef[d]);
  int node = find_node_ix(m, ix[0], ix[1], ix[2]);
  (void)startNode;
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

// FindCellIndex (meshAMRgeneric.h:2256-2323); returns false where the reference returns -1
__device__ __forceinline__ bool find_cell_index(const DevMesh &m, const double x[3], int node, int ijk[3]) {
  const int lev = m.nodeLevel[node];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double lo = m.nxmin[3 * node + d], hi = m.nxmax[3 * node + d];
    if ((x[d] < lo) || (hi < x[d])) return false;
    const double dx = m.dxRoot[d] / (1 << lev) / double(m.N[d]);
    int c = (int)((x[d] - lo) / dx);
    if (c == m.N[d]) c = m.N[d] - 1;
    ijk[d] = c;
  }
  return true;
}

}  // namespace amps
#include "cplr_stencil.cuh"
namespace amps {

// unique-node tables of the coupler (read by the AMR stencil, whose cells belong to several blocks)
struct CplrUnique {
  const double *E, *B;  // [nCenters][3]
};
template <int K>
__device__ __forceinline__ void cplr_gather_unique(const BgStencil &st, const double *__restrict__ U, double out[K]) {
  for (int q = 0; q < K; q++) out[q] = 0.0;
  for (int s = 0; s < st.n; s++) {
    const double *t = U + (size_t)K * st.nd[s];
    for (int q = 0; q < K; q++) out[q] += st.w[s] * t[q];
  }
}
__device__ __noinline__ bool background_fields_amr(const DevMesh &m, const CplrUnique &U, const double x[3], int leaf, double E[3], double B[3]) {
  BgStencil st;
  if (!cplr_linear_stencil_amr(m, x, leaf, st) || st.n == 0 || st.overflow) return false;
  cplr_gather_unique<3>(st, U.E, E);
  cplr_gather_unique<3>(st, U.B, B);
  return true;
}

// E, B at x inside leaf `leaf` through the coupler stencil; false = the reference would exit()
__device__ __forceinline__ bool background_fields(const DevMesh &m, int interp, const double *__restrict__ tile, const CplrUnique &U,
                                                  const double x[3], int leaf, double E[3], double B[3]) {
  const LeafGeo &lg = m.leaf[leaf];
  if (interp == AMPS_CPLR_CELL_CENTERED_LINEAR && !leaf_single_level(lg)) return background_fields_amr(m, U, x, leaf, E, B);
  const double *T = tile + (size_t)leaf * m.nCenterLocal * 6;
  E[0] = E[1] = E[2] = 0.0, B[0] = B[1] = B[2] = 0.0;
  if (interp == AMPS_CPLR_CELL_CENTERED_LINEAR) {
    // GetTriliniarInterpolationStencil (:820-907), Normalize only when Length != 8
    const double iLoc = (x[0] - lg.xmin[0]) / (lg.xmax[0] - lg.xmin[0]) * m.N[0];
    const double jLoc = (x[1] - lg.xmin[1]) / (lg.xmax[1] - lg.xmin[1]) * m.N[1];
    const double kLoc = (x[2] - lg.xmin[2]) / (lg.xmax[2] - lg.xmin[2]) * m.N[2];
    const int i0 = (iLoc < 0.5) ? -1 : (int)(iLoc - 0.50);
    const int j0 = (jLoc < 0.5) ? -1 : (int)(jLoc - 0.50);
    const int k0 = (kLoc < 0.5) ? -1 : (int)(kLoc - 0.50);
    const double w0 = iLoc - (i0 + 0.5), w1 = jLoc - (j0 + 0.5), w2 = kLoc - (k0 + 0.5);
    double w[8];
    w[0] = (1.0 - w0) * (1.0 - w1) * (1.0 - w2);
    w[1] = (1.0 - w0) * (1.0 - w1) * w2;
    w[2] = (1.0 - w0) * w1 * (1.0 - w2);
    w[3] = (1.0 - w0) * w1 * w2;
    w[4] = w0 * (1.0 - w1) * (1.0 - w2);
    w[5] = w0 * (1.0 - w1) * w2;
    w[6] = w0 * w1 * (1.0 - w2);
    w[7] = w0 * w1 * w2;
    unsigned valid = 0xffu;
    if (!m.periodic && lg.face) {
      if ((lg.face & 1) && i0 < 0) valid &= 0xf0u;
      if ((lg.face & 2) && i0 + 1 >= m.N[0]) valid &= 0x0fu;
      if ((lg.face & 4) && j0 < 0) valid &= 0xccu;
      if ((lg.face & 8) && j0 + 1 >= m.N[1]) valid &= 0x33u;
      if ((lg.face & 16) && k0 < 0) valid &= 0xaau;
      if ((lg.face & 32) && k0 + 1 >= m.N[2]) valid &= 0x55u;
    }
    if (valid != 0xffu) {
      double norm = 0.0;
#pragma unroll
      for (int s = 0; s < 8; s++)
        if (valid & (1u << s)) norm += w[s];
      if (norm > 0.0) {
#pragma unroll
        for (int s = 0; s < 8; s++) w[s] /= norm;
      }
    }
    const int nd0 = centerLocalNumber(m, i0, j0, k0);
    const int BS0 = m.TN[0], BS1 = m.TN[0] * m.TN[1];
    // E over the whole stencil first, then B (two loops in the reference)
#pragma unroll
    for (int s = 0; s < 8; s++)
      if (valid & (1u << s)) {
        const double *t = T + 6 * (nd0 + ((s >> 2) & 1) + ((s >> 1) & 1) * BS0 + (s & 1) * BS1);
        E[0] += w[s] * t[0], E[1] += w[s] * t[1], E[2] += w[s] * t[2];
      }
#pragma unroll
    for (int s = 0; s < 8; s++)
      if (valid & (1u << s)) {
        const double *t = T + 6 * (nd0 + ((s >> 2) & 1) + ((s >> 1) & 1) * BS0 + (s & 1) * BS1) + 3;
        B[0] += w[s] * t[0], B[1] += w[s] * t[1], B[2] += w[s] * t[2];
      }
    return true;
  }
  // Constant::InitStencil (:168-220): the cell that contains x, weight 1
  int ijk[3];
  if (!find_cell_index(m, x, lg.node, ijk)) return false;
  const double *t = T + 6 * centerLocalNumber(m, ijk[0], ijk[1], ijk[2]);
#pragma unroll
  for (int d = 0; d < 3; d++) E[d] += 1.0 * t[d], B[d] += 1.0 * t[3 + d];
  return true;
}

struct TpParams {
  int interp, backward, boundaryMode;
  double c, rSphere;
  long long exitCap;
  CplrUnique U;
};

#ifndef TP_MIN_CTAS
#define TP_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(128, TP_MIN_CTAS) move_relativistic_boris_kernel(DevMesh m, DevSpecies sp, TpParams tp, ParticleSoA p, const int *__restrict__ nSlots,
                                                                     const double *__restrict__ bgTile, int *__restrict__ cellCount,
                                                                     DevMoveStats *__restrict__ stats, amps_gpu_exit_record *__restrict__ exitBuf,
                                                                     unsigned long long *__restrict__ exitCount) {
  __shared__ FaceGeo sFace;
  if (threadIdx.x == 0) init_faces(m, sFace);
  __syncthreads();
  const int n = *nSlots;
  const int C = m.cellsPerBlock;
  unsigned int nMoved = 0, nXCell = 0, nXBlock = 0, nLeft = 0, nNotUsed = 0, nWrap = 0, nErr = 0, nSub = 0;
  const double SpeedOfLight = tp.c;
  const bool backward = tp.backward != 0;

  for (int ip = blockIdx.x * blockDim.x + threadIdx.x; ip < n; ip += gridDim.x * blockDim.x) {
    const int oldKey = p.key[ip];
    if (oldKey < 0) continue;
    nMoved++;
    double xInit[3] = {p.x[0][ip], p.x[1][ip], p.x[2][ip]};
    double vInit[3] = {p.v[0][ip], p.v[1][ip], p.v[2][ip]};
    double xFinal[3], vFinal[3];
    const int spec = p.spec[ip] & 0x3f;
    const int startLeaf = oldKey / C;
    const double ElectricCharge = sp.charge[spec], mass = sp.mass[spec];
    double dtTotalIn = (sp.timeStepMode == AMPS_DT_SPECIES_GLOBAL) ? sp.dt[spec] : sp.dt[0];
    int leaf = startLeaf, node = m.leaf[startLeaf].node;
    int outcome = 0;  // 0 running/finished, 1 left domain, 2 not in use, 3 error
    if (dtTotalIn == 0.0) {
      for (int d = 0; d < 3; d++) xFinal[d] = xInit[d], vFinal[d] = vInit[d];
    } else
      while (dtTotalIn > 0.0) {
        nSub++;
        double gamma = 1.0 / sqrt(1.0 - (vInit[0] * vInit[0] + vInit[1] * vInit[1] + vInit[2] * vInit[2]) / (SpeedOfLight * SpeedOfLight));
        double E[3], B[3];
        if (!background_fields(m, tp.interp, bgTile, tp.U, xInit, leaf, E, B)) {
          outcome = 3;
          break;
        }
        double dt;
        const double absB = sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
        if (absB > 1.0E-25) {
          const double PiTimes2 = 6.28318530717958647692528676655900576839433879875021;
          const double g2 = 1.0 / sqrt(1.0 - (vInit[0] * vInit[0] + vInit[1] * vInit[1] + vInit[2] * vInit[2]) / (SpeedOfLight * SpeedOfLight));
          const double GyroFreq = fabs(ElectricCharge) * absB / (PiTimes2 * mass * g2);
          const double dtMax = 1.0 / GyroFreq;
          dt = (dtMax < dtTotalIn) ? dtMax : dtTotalIn;
        } else
          dt = dtTotalIn;
        dtTotalIn -= dt;
        if (backward)
          for (int d = 0; d < 3; d++) vInit[d] = -vInit[d], B[d] = -B[d];
        const double QdT_over_twoM = ElectricCharge * dt / (2.0 * mass);
        double uMinus[3];
        for (int d = 0; d < 3; d++) uMinus[d] = gamma * vInit[d] + QdT_over_twoM * E[d];
        double t[3], s[3], uPrime[3], uPlus[3], l = 0.0;
        gamma = sqrt(1.0 + (uMinus[0] * uMinus[0] + uMinus[1] * uMinus[1] + uMinus[2] * uMinus[2]) / (SpeedOfLight * SpeedOfLight));
        for (int d = 0; d < 3; d++) {
          t[d] = QdT_over_twoM / gamma * B[d];
          l += t[d] * t[d];
        }
        uPrime[0] = uMinus[1] * t[2] - uMinus[2] * t[1];
        uPrime[1] = uMinus[2] * t[0] - uMinus[0] * t[2];
        uPrime[2] = uMinus[0] * t[1] - uMinus[1] * t[0];
        for (int d = 0; d < 3; d++) uPrime[d] += uMinus[d];
        for (int d = 0; d < 3; d++) s[d] = 2.0 * t[d] / (1.0 + l);
        uPlus[0] = uPrime[1] * s[2] - uPrime[2] * s[1];
        uPlus[1] = uPrime[2] * s[0] - uPrime[0] * s[2];
        uPlus[2] = uPrime[0] * s[1] - uPrime[1] * s[0];
        for (int d = 0; d < 3; d++) uPlus[d] += uMinus[d];
        double uFinal[3];
        for (int d = 0; d < 3; d++) uFinal[d] = uPlus[d] + QdT_over_twoM * E[d];
        gamma = sqrt(1.0 + (uFinal[0] * uFinal[0] + uFinal[1] * uFinal[1] + uFinal[2] * uFinal[2]) / (SpeedOfLight * SpeedOfLight));
        for (int d = 0; d < 3; d++) {
          vFinal[d] = uFinal[d] / gamma;
          xFinal[d] = xInit[d] + vFinal[d] * dt;
        }
        if (backward)
          for (int d = 0; d < 3; d++) vFinal[d] = -vFinal[d], vInit[d] = -vInit[d];

        // internal sphere (:270-302): project on the sphere, hand to the callback, delete
        int newNode;
        if (tp.rSphere > 0.0) {
          const double rFinal2 = xFinal[0] * xFinal[0] + xFinal[1] * xFinal[1] + xFinal[2] * xFinal[2];
          if (rFinal2 < tp.rSphere * tp.rSphere) {
            const double r = sqrt(rFinal2);
            for (int d = 0; d < 3; d++) xFinal[d] *= tp.rSphere / r;
            newNode = find_tree_node_plain(m, xFinal, node);
            add_exit_record(exitBuf, exitCount, tp.exitCap, p.ptr[ip], spec, AMPS_EXIT_SPHERE, newNode >= 0 ? m.nodeLeaf[newNode] : -1, xFinal, vFinal);
            outcome = 1;
            break;
          }
        }
        newNode = find_tree_node_plain(m, xFinal, node);
        if (newNode < 0) {
          if (tp.boundaryMode == AMPS_BOUNDARY_DELETE) {
            outcome = 1;
            break;
          }
          // face-intersection search (:320-446; reference defects restated as intended, see the oracle)
          int nIntersectionFace = -1;
          double vEffective[3], r0[3], dtIntersection = -1.0;
          for (int d = 0; d < 3; d++) vEffective[d] = backward ? xInit[d] - xFinal[d] : xFinal[d] - xInit[d];
          for (int nface = 0; nface < 6; nface++) {
            double cx = 0.0, cv = 0.0, rr[3];
            for (int d = 0; d < 3; d++) {
              rr[d] = (backward ? xFinal[d] : xInit[d]) - sFace.x0[nface][d];
              cx += rr[d] * sFace.norm[nface][d];
              cv += vEffective[d] * sFace.norm[nface][d];
            }
            const double dtEffective = backward ? ((cv < 0.0) ? -cx / cv : -1.0) : ((cv > 0.0) ? -cx / cv : -1.0);
            if (dtEffective > 0.0) {
              if ((dtIntersection < 0.0) || ((dtEffective < dtIntersection) && (dtEffective > 0.0))) {
                double cE0 = 0.0, cE1 = 0.0;
                for (int d = 0; d < 3; d++) {
                  const double c = rr[d] + dtEffective * vEffective[d];
                  cE0 += c * sFace.e0[nface][d], cE1 += c * sFace.e1[nface][d];
                }
                if ((cE0 < -m.eps) || (cE0 > sFace.lE0[nface] + m.eps) || (cE1 < -m.eps) || (cE1 > sFace.lE1[nface] + m.eps)) continue;
                nIntersectionFace = nface, dtIntersection = dtEffective;
                for (int d = 0; d < 3; d++) r0[d] = rr[d];
              }
            }
          }
          (void)r0;
          if (nIntersectionFace == -1) {
            outcome = 3;
            break;
          }
          if (backward) {
            for (int d = 0; d < 3; d++) {
              xInit[d] = xFinal[d] + dtIntersection * (xInit[d] - xFinal[d]) - sFace.norm[nIntersectionFace][d] * m.eps;
              vInit[d] = vFinal[d] + dtIntersection * (vInit[d] - vFinal[d]);
            }
          } else {
            for (int d = 0; d < 3; d++) {
              xInit[d] += dtIntersection * (xFinal[d] - xInit[d]) - sFace.norm[nIntersectionFace][d] * m.eps;
              vInit[d] += dtIntersection * (vFinal[d] - vInit[d]);
            }
          }
          newNode = find_tree_node_plain(m, xInit, node);
          if (newNode < 0) {
            for (int d = 0; d < 3; d++) {
              if (m.xGlobalMin[d] >= xInit[d]) xInit[d] = m.xGlobalMin[d] + m.eps;
              if (m.xGlobalMax[d] <= xInit[d]) xInit[d] = m.xGlobalMax[d] - m.eps;
            }
            newNode = find_tree_node_plain(m, xInit, node);
            if (newNode < 0) {
              outcome = 3;
              break;
            }
          }
          if (tp.boundaryMode == AMPS_BOUNDARY_USER_FUNCTION) {
            add_exit_record(exitBuf, exitCount, tp.exitCap, p.ptr[ip], spec, nIntersectionFace, m.nodeLeaf[newNode], xInit, vInit);
            outcome = 1;  // the callback deletes (srcEarth/CutoffRigidity.cpp:129-230)
          } else {
            outcome = 3;  // SPECULAR: _PARTICLE_REJECTED_ON_THE_FACE_ -> exit("not implemented") in the reference
          }
          break;
        }
        if (!(m.nodeFlags[newNode] & AMPS_NODE_USED)) {
          outcome = 2;
          break;
        }
        const int newLeaf = m.nodeLeaf[newNode];
        if (newLeaf < 0) {  // fields of a block that is not allocated on this rank
          outcome = 3;
          break;
        }
        node = newNode, leaf = newLeaf;
        for (int d = 0; d < 3; d++) xInit[d] = xFinal[d], vInit[d] = vFinal[d];
      }

    int newKey = -1;
    if (outcome == 0) {
      int ijk[3];
      if (!find_cell_index(m, xFinal, node, ijk)) {
        outcome = 3;
      } else {
        int newLeaf = leaf;
        const int realLeaf = m.leaf[newLeaf].real;
        if (realLeaf >= 0) {  // periodic ghost -> real (pic_bc_periodic.cpp:100-134)
          const LeafGeo &gg = m.leaf[newLeaf];
          const LeafGeo &rg = m.leaf[realLeaf];
          for (int d = 0; d < 3; d++) {
            xFinal[d] += rg.xmin[d] - gg.xmin[d];
            if (xFinal[d] < rg.xmin[d]) xFinal[d] = rg.xmin[d];
            if (xFinal[d] >= rg.xmax[d]) xFinal[d] = rg.xmax[d] - 1.0E-10 * (rg.xmax[d] - rg.xmin[d]);
          }
          newLeaf = realLeaf;
          nWrap++;
        }
        newKey = newLeaf * C + ijk[0] + m.N[0] * (ijk[1] + m.N[1] * ijk[2]);
        if (newLeaf != startLeaf) nXBlock++;
        else if (newKey != oldKey) nXCell++;
      }
    }
    if (outcome == 1) nLeft++;
    else if (outcome == 2) nNotUsed++;
    else if (outcome == 3) nErr++;
    if (newKey >= 0) {
      p.x[0][ip] = xFinal[0], p.x[1][ip] = xFinal[1], p.x[2][ip] = xFinal[2];
      p.v[0][ip] = vFinal[0], p.v[1][ip] = vFinal[1], p.v[2][ip] = vFinal[2];
      atomicAdd(&cellCount[newKey], 1);
    }
    if (newKey != oldKey) p.key[ip] = newKey;
  }

  {
    unsigned v = nSub;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&stats->n_sub_steps, (unsigned long long)v);
  }
  flush_move_counters(stats, nMoved, nXCell, nXBlock, nLeft, nNotUsed, nWrap, nErr);
}

// ---- structure cache of the AMR stencil (cplr_stencil.cuh): one thread per neighbour slot / per dual cell of a tabulated block ----
__global__ void __launch_bounds__(128) build_cplr_cache_kernel(DevMesh m, int *__restrict__ neib26, const int *__restrict__ tabLeaf, int nTab,
                                                              unsigned char *__restrict__ tab) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nNeib = (long long)m.nLeaves * 27;
  if (t < nNeib) {
    const int leaf = (int)(t / 27), q = (int)(t - 27LL * leaf);
    const int side[3] = {q % 3 - 1, (q / 3) % 3 - 1, q / 9 - 1};
    neib26[t] = cs_neib(m, m.leaf[leaf].node, side);  // m.neib26 is still null here: the lattice probe itself
    return;
  }
  const long long u = t - nNeib;
  const int nEnt = mb_entries(m);
  if (u >= (long long)nTab * nEnt) return;
  const int it = (int)(u / nEnt), e = (int)(u - (long long)it * nEnt);
  const int n0 = m.N[0] + 2, n1 = m.N[1] + 2;
  const int ijkMin[3] = {e % n0 - 1, (e / n0) % n1 - 1, e / (n0 * n1) - 1};
  cs_multiblock_structure(m, m.leaf[tabLeaf[it]].node, ijkMin, tab + (size_t)u * MB_ENTRY);
}
size_t cplr_cache_table_bytes(const DevMesh &m) { return (size_t)(m.N[0] + 2) * (m.N[1] + 2) * (m.N[2] + 2) * MB_ENTRY; }
void launch_build_cplr_cache(const DevMesh &m, int *neib26, const int *tabLeaf, int nTab, unsigned char *tab, cudaStream_t s) {
  DevMesh bare = m;
  bare.neib26 = nullptr, bare.mbSlot = nullptr, bare.mbTab = nullptr;
  const long long n = (long long)m.nLeaves * 27 + (long long)nTab * (m.N[0] + 2) * (m.N[1] + 2) * (m.N[2] + 2);
  build_cplr_cache_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(bare, neib26, tabLeaf, nTab, tab);
}

void launch_move_relativistic_boris(const DevMesh &m, const DevSpecies &sp, int interp, int backward, double c, double rSphere, long long exitCap,
                                    ParticleSoA p, const int *nSlots, long long nUpper, const double *bgTile, const double *uE, const double *uB,
                                    int *cellCount, DevMoveStats *stats, amps_gpu_exit_record *exitBuf, unsigned long long *exitCount,
                                    cudaStream_t s) {
  TpParams tp;
  tp.interp = interp, tp.backward = backward, tp.boundaryMode = sp.boundaryMode, tp.c = c, tp.rSphere = rSphere, tp.exitCap = exitCap;
  tp.U.E = uE, tp.U.B = uB;
  long long g = (nUpper + 127) / 128;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  move_relativistic_boris_kernel<<<(int)g, 128, 0, s>>>(m, sp, tp, p, nSlots, bgTile, cellCount, stats, exitBuf, exitCount);
}

// ------------------------------------------------------------------------------------------------
// a6: PIC::Mover::Boris + BorisSplitAcceleration_default  src/pic/pic_mover_boris.cpp:126-553, :22-123
// ------------------------------------------------------------------------------------------------
// kMarkidis: PIC::Mover::Markidis2010 (pic_mover_boris.cpp:557-835): same skeleton, eqs 22-23 of Markidis et al. 2010 for the velocity
template <bool kMarkidis>
__global__ void __launch_bounds__(128) move_boris_kernel(DevMesh m, DevSpecies sp, TpParams tp, double gravityGM, ParticleSoA p, const int *__restrict__ nSlots,
                                                        const double *__restrict__ bgTile, int *__restrict__ cellCount, DevMoveStats *__restrict__ stats,
                                                        amps_gpu_exit_record *__restrict__ exitBuf, unsigned long long *__restrict__ exitCount) {
  __shared__ FaceGeo sFace;
  if (threadIdx.x == 0) init_faces(m, sFace);
  __syncthreads();
  const int n = *nSlots;
  const int C = m.cellsPerBlock;
  unsigned int nMoved = 0, nXCell = 0, nXBlock = 0, nLeft = 0, nNotUsed = 0, nWrap = 0, nErr = 0;

  for (int ip = blockIdx.x * blockDim.x + threadIdx.x; ip < n; ip += gridDim.x * blockDim.x) {
    const int oldKey = p.key[ip];
    if (oldKey < 0) continue;
    nMoved++;
    double xInit[3] = {p.x[0][ip], p.x[1][ip], p.x[2][ip]};
    double vInit[3] = {p.v[0][ip], p.v[1][ip], p.v[2][ip]};
    double xFinal[3], vFinal[3];
    const int spec = p.spec[ip] & 0x3f;
    const int startLeaf = oldKey / C;
    const int startNode = m.leaf[startLeaf].node;
    const double dtTotal = (sp.timeStepMode == AMPS_DT_SPECIES_GLOBAL) ? sp.dt[spec] : sp.dt[0];
    int outcome = 0, node = -1;

    bool pushed = false;
    if (kMarkidis) {
      double E[3], B[3];
      if (!background_fields(m, tp.interp, bgTile, tp.U, xInit, startLeaf, E, B)) outcome = 3;
      else {
        const double QdT_over_m = sp.charge[spec] * dtTotal / sp.mass[spec];
        const double QdT_over_2m = 0.5 * QdT_over_m;
        double v_prime[3];
        for (int d = 0; d < 3; d++) v_prime[d] = vInit[d] + QdT_over_m * E[d];
        const double Denominator = 1.0 / (1.0 + QdT_over_2m * QdT_over_2m * (B[0] * B[0] + B[1] * B[1] + B[2] * B[2]));
        double n1[3];
        n1[0] = v_prime[1] * B[2] - v_prime[2] * B[1];
        n1[1] = v_prime[2] * B[0] - v_prime[0] * B[2];
        n1[2] = v_prime[0] * B[1] - v_prime[1] * B[0];
        const double n2 = QdT_over_2m * QdT_over_2m * (v_prime[0] * B[0] + v_prime[1] * B[1] + v_prime[2] * B[2]);
        for (int d = 0; d < 3; d++) {
          vFinal[d] = Denominator * (v_prime[d] + QdT_over_2m * n1[d] + n2 * B[d]);
          xFinal[d] = xInit[d] + dtTotal * vFinal[d];
        }
        pushed = true;
      }
    }
    // BorisSplitAcceleration_default: fields in the cell of x (fail-safe: search the block again)
    double acclInit[3] = {0.0, 0.0, 0.0}, rotInit[3] = {0.0, 0.0, 0.0};
    if (!kMarkidis) {
      int fieldLeaf = startLeaf, ijk[3];
      if (!find_cell_index(m, xInit, startNode, ijk)) {
        const int fn = find_tree_node_plain(m, xInit, startNode);
        fieldLeaf = (fn >= 0) ? m.nodeLeaf[fn] : -1;
        if (fieldLeaf < 0 || !find_cell_index(m, xInit, fn, ijk)) outcome = 3;
      }
      double E[3], B[3];
      if (outcome == 0 && !background_fields(m, tp.interp, bgTile, tp.U, xInit, fieldLeaf, E, B)) outcome = 3;
      if (outcome == 0) {
        const double ElectricCharge = sp.charge[spec], mass = sp.mass[spec];
        if (ElectricCharge != 0.0) {
          const double Charge2Mass = ElectricCharge / mass;
          for (int d = 0; d < 3; d++) {
            acclInit[d] += Charge2Mass * E[d];
            rotInit[d] -= Charge2Mass * B[d];
          }
        }
        if (gravityGM != 0.0) {
          const double r2 = xInit[0] * xInit[0] + xInit[1] * xInit[1] + xInit[2] * xInit[2];
          const double r = sqrt(r2);
          for (int d = 0; d < 3; d++) acclInit[d] -= gravityGM / r2 * xInit[d] / r;
        }
      }
    }
    if (!kMarkidis && outcome == 0) {
      double dtTempOverTwo, dtTemp;
      if (tp.backward) dtTemp = -dtTotal, dtTempOverTwo = -dtTotal / 2.0;
      else dtTemp = dtTotal, dtTempOverTwo = dtTotal / 2.0;
      double u[3], h[3], U[3];
      u[0] = vInit[0] + dtTempOverTwo * acclInit[0];
      u[1] = vInit[1] + dtTempOverTwo * acclInit[1];
      u[2] = vInit[2] + dtTempOverTwo * acclInit[2];
      h[0] = -dtTempOverTwo * rotInit[0];
      h[1] = -dtTempOverTwo * rotInit[1];
      h[2] = -dtTempOverTwo * rotInit[2];
      const double h2 = h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
      const double uh = u[0] * h[0] + u[1] * h[1] + u[2] * h[2];
      U[0] = ((1 - h2) * u[0] + 2 * (u[1] * h[2] - h[1] * u[2] + uh * h[0])) / (1 + h2);
      U[1] = ((1 - h2) * u[1] + 2 * (u[2] * h[0] - h[2] * u[0] + uh * h[1])) / (1 + h2);
      U[2] = ((1 - h2) * u[2] + 2 * (u[0] * h[1] - h[0] * u[1] + uh * h[2])) / (1 + h2);
      vFinal[0] = U[0] + dtTempOverTwo * acclInit[0];
      vFinal[1] = U[1] + dtTempOverTwo * acclInit[1];
      vFinal[2] = U[2] + dtTempOverTwo * acclInit[2];
      xFinal[0] = xInit[0] + dtTemp * vFinal[0];
      xFinal[1] = xInit[1] + dtTemp * vFinal[1];
      xFinal[2] = xInit[2] + dtTemp * vFinal[2];
    }
    (void)pushed;
    if (outcome == 0) {
      bool hitSphere = false;
      if (tp.rSphere > 0.0) {
        const double R = tp.rSphere;
        const double dx0 = xFinal[0] - 0.0, dx1 = xFinal[1] - 0.0, dx2 = xFinal[2] - 0.0;
        const double r2 = dx0 * dx0 + dx1 * dx1 + dx2 * dx2;
        if (r2 < R * R) {
          double r = sqrt(r2);
          if (r <= 0.0) r = 1.0;
          xFinal[0] = 0.0 + dx0 * (R / r);
          xFinal[1] = 0.0 + dx1 * (R / r);
          xFinal[2] = 0.0 + dx2 * (R / r);
          const int nn = find_tree_node_plain(m, xFinal, startNode);
          add_exit_record(exitBuf, exitCount, tp.exitCap, p.ptr[ip], spec, AMPS_EXIT_SPHERE, nn >= 0 ? m.nodeLeaf[nn] : -1, xFinal, vFinal);
          outcome = 1;
          hitSphere = true;
        }
      }
      if (!hitSphere) {
        node = find_tree_node_plain(m, xFinal, startNode);
        if (node < 0) {
          if (tp.boundaryMode == AMPS_BOUNDARY_DELETE) outcome = 1;
          else {
            int face, exitNode;
            const int code = domain_exit_vmiddle(m, sFace, tp.boundaryMode, dtTotal, xInit, vInit, xFinal, vFinal, startNode, &face, &exitNode);
            if (code == 0) {
              add_exit_record(exitBuf, exitCount, tp.exitCap, p.ptr[ip], spec, face, m.nodeLeaf[exitNode], xInit, vInit);
              outcome = 1;
            } else
              outcome = 3;  // specular: _PARTICLE_REJECTED_ON_THE_FACE_ -> exit("not implemented") :461
          }
        } else if (!(m.nodeFlags[node] & AMPS_NODE_USED))
          outcome = kMarkidis ? 3 : 2;  // Markidis2010 has no "not in use" return: its block==NULL test exit()s (:803)
      }
    }

    int newKey = -1;
    if (outcome == 0) {
      int ijk[3];
      int newLeaf = m.nodeLeaf[node];
      if (!find_cell_index(m, xFinal, node, ijk)) outcome = 3;
      else if (newLeaf < 0) outcome = kMarkidis ? 3 : 1;  // block not allocated here (:1290-1302 analogue)
      else {
        const int realLeaf = m.leaf[newLeaf].real;
        if (realLeaf >= 0) {
          const LeafGeo &gg = m.leaf[newLeaf];
          const LeafGeo &rg = m.leaf[realLeaf];
          for (int d = 0; d < 3; d++) {
            xFinal[d] += rg.xmin[d] - gg.xmin[d];
            if (xFinal[d] < rg.xmin[d]) xFinal[d] = rg.xmin[d];
            if (xFinal[d] >= rg.xmax[d]) xFinal[d] = rg.xmax[d] - 1.0E-10 * (rg.xmax[d] - rg.xmin[d]);
          }
          newLeaf = realLeaf;
          nWrap++;
        }
        newKey = newLeaf * C + ijk[0] + m.N[0] * (ijk[1] + m.N[1] * ijk[2]);
        if (newLeaf != startLeaf) nXBlock++;
        else if (newKey != oldKey) nXCell++;
      }
    }
    if (outcome == 1) nLeft++;
    else if (outcome == 2) nNotUsed++;
    else if (outcome == 3) nErr++;
    if (newKey >= 0) {
      p.x[0][ip] = xFinal[0], p.x[1][ip] = xFinal[1], p.x[2][ip] = xFinal[2];
      p.v[0][ip] = vFinal[0], p.v[1][ip] = vFinal[1], p.v[2][ip] = vFinal[2];
      atomicAdd(&cellCount[newKey], 1);
    }
    if (newKey != oldKey) p.key[ip] = newKey;
  }
  flush_move_counters(stats, nMoved, nXCell, nXBlock, nLeft, nNotUsed, nWrap, nErr);
}

void launch_move_boris(const DevMesh &m, const DevSpecies &sp, bool markidis, int interp, int backward, double c, double rSphere, long long exitCap,
                       double gravityGM, ParticleSoA p, const int *nSlots, long long nUpper, const double *bgTile, const double *uE, const double *uB, int *cellCount,
                       DevMoveStats *stats, amps_gpu_exit_record *exitBuf, unsigned long long *exitCount, cudaStream_t s) {
  TpParams tp;
  tp.interp = interp, tp.backward = backward, tp.boundaryMode = sp.boundaryMode, tp.c = c, tp.rSphere = rSphere, tp.exitCap = exitCap;
  tp.U.E = uE, tp.U.B = uB;
  long long g = (nUpper + 127) / 128;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  if (markidis) move_boris_kernel<true><<<(int)g, 128, 0, s>>>(m, sp, tp, gravityGM, p, nSlots, bgTile, cellCount, stats, exitBuf, exitCount);
  else move_boris_kernel<false><<<(int)g, 128, 0, s>>>(m, sp, tp, gravityGM, p, nSlots, bgTile, cellCount, stats, exitBuf, exitCount);
}

// ------------------------------------------------------------------------------------------------
// a9: PIC::Mover::Relativistic::GuidingCenter  src/pic/pic_mover_relativistic_guiding_center.cpp
//     InitiateMagneticMoment :19-93, Mover_FirstOrder :96-409 (DELETE boundary; the other branches of the
//     reference exit() or use an undefined face), fields + 15 drift variables through the coupler stencil
//     (GetVarForRelativisticGCA, pic.h:8643-8680) on per-leaf tiles [nCenterLocal][15]
// ------------------------------------------------------------------------------------------------
__global__ void stage_center_table_kernel(DevMesh m, int nVar, const double *__restrict__ var, double *__restrict__ tile) {
  const int leaf = blockIdx.x;
  const int *cuid = m.centerUid + (size_t)leaf * m.nCenterLocal;
  double *dst = tile + (size_t)leaf * m.nCenterLocal * nVar;
  for (int e = threadIdx.x; e < m.nCenterLocal * nVar; e += blockDim.x) {
    const int i = e / nVar, q = e - nVar * i;
    const int u = cuid[i];
    dst[e] = (u >= 0) ? var[nVar * (size_t)u + q] : 0.0;
  }
}
void launch_stage_center_table(const DevMesh &m, int nVar, const double *var, double *tile, cudaStream_t s) {
  stage_center_table_kernel<<<m.nLeaves, 256, 0, s>>>(m, nVar, var, tile);
}

// the coupler stencil at x inside `leaf` (same arithmetic as background_fields), kept for several gathers.
// Single-level neighbourhoods (and the piecewise-constant coupler) use the 8-slot register form: slot s = 4i+2j+k of the
// 2x2x2 centres from nd0, weight 0 for centres outside the domain (adding w*t = 0 in the reference's slot order leaves
// every partial sum unchanged).  Leaves next to another refinement level use the 64-entry list of unique centre ids.
struct BgStencil8 {
  double w[8];
  int nd0, BS0, BS1;
};
// returns 0 = the reference would exit(), 1 = s8 holds the stencil, 2 = `big` holds it (AMR)
__device__ __forceinline__ int background_stencil(const DevMesh &m, int interp, const double x[3], int leaf, BgStencil8 &s8, BgStencil &big) {
  const LeafGeo &lg = m.leaf[leaf];
  if (interp == AMPS_CPLR_CELL_CENTERED_LINEAR && !leaf_single_level(lg))
    return (cplr_linear_stencil_amr(m, x, leaf, big) && big.n > 0 && !big.overflow) ? 2 : 0;
  s8.BS0 = m.TN[0], s8.BS1 = m.TN[0] * m.TN[1];
  if (interp == AMPS_CPLR_CELL_CENTERED_LINEAR) {
    const double iLoc = (x[0] - lg.xmin[0]) / (lg.xmax[0] - lg.xmin[0]) * m.N[0];
    const double jLoc = (x[1] - lg.xmin[1]) / (lg.xmax[1] - lg.xmin[1]) * m.N[1];
    const double kLoc = (x[2] - lg.xmin[2]) / (lg.xmax[2] - lg.xmin[2]) * m.N[2];
    const int i0 = (iLoc < 0.5) ? -1 : (int)(iLoc - 0.50);
    const int j0 = (jLoc < 0.5) ? -1 : (int)(jLoc - 0.50);
    const int k0 = (kLoc < 0.5) ? -1 : (int)(kLoc - 0.50);
    const double w0 = iLoc - (i0 + 0.5), w1 = jLoc - (j0 + 0.5), w2 = kLoc - (k0 + 0.5);
    double w[8];
    w[0] = (1.0 - w0) * (1.0 - w1) * (1.0 - w2);
    w[1] = (1.0 - w0) * (1.0 - w1) * w2;
    w[2] = (1.0 - w0) * w1 * (1.0 - w2);
    w[3] = (1.0 - w0) * w1 * w2;
    w[4] = w0 * (1.0 - w1) * (1.0 - w2);
    w[5] = w0 * (1.0 - w1) * w2;
    w[6] = w0 * w1 * (1.0 - w2);
    w[7] = w0 * w1 * w2;
    // x may lie outside the leaf (GuidingCenter::Mover_FirstOrder :713): indices past the ghost layer are
    // out-of-bounds reads in the reference -> error
    if (!(iLoc >= -1.0e9 && iLoc <= 1.0e9 && jLoc >= -1.0e9 && jLoc <= 1.0e9 && kLoc >= -1.0e9 && kLoc <= 1.0e9)) return 0;
    if (i0 < -m.g[0] || i0 + 1 > m.N[0] + m.g[0] - 1 || j0 < -m.g[1] || j0 + 1 > m.N[1] + m.g[1] - 1 || k0 < -m.g[2] ||
        k0 + 1 > m.N[2] + m.g[2] - 1)
      return 0;
    unsigned valid = 0xffu;
    if (!m.periodic && lg.face) {  // AddCell drops centres outside the domain (pic.h:7235-7245)
      const int o0[3] = {i0, j0, k0};
      const unsigned lowMask[3] = {0x0fu, 0x33u, 0x55u};  // stencil slots whose index along d is o0[d]
#pragma unroll
      for (int d = 0; d < 3; d++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
          const int a = o0[d] + b;
          if (((lg.face >> (2 * d)) & 1 && a < 0) || ((lg.face >> (2 * d + 1)) & 1 && a >= m.N[d])) valid &= b ? lowMask[d] : ~lowMask[d] & 0xffu;
        }
    }
    if (valid == 0u) return 0;
    double norm = 0.0;
    if (valid != 0xffu) {
#pragma unroll
      for (int s = 0; s < 8; s++)
        if (valid & (1u << s)) norm += w[s];
    }
    s8.nd0 = centerLocalNumber(m, i0, j0, k0);
#pragma unroll
    for (int s = 0; s < 8; s++) {
      double ws = w[s];
      if (valid != 0xffu) ws = (valid & (1u << s)) ? ((norm > 0.0) ? w[s] / norm : w[s]) : 0.0;
      s8.w[s] = ws;
    }
    return 1;
  }
  int ijk[3];
  if (!find_cell_index(m, x, lg.node, ijk)) return 0;
  s8.nd0 = centerLocalNumber(m, ijk[0], ijk[1], ijk[2]);
  s8.w[0] = 1.0;
#pragma unroll
  for (int s = 1; s < 8; s++) s8.w[s] = 0.0;
  s8.BS0 = 0, s8.BS1 = 0;  // every slot reads the cell itself; only slot 0 carries weight
  return 1;
}
// T = the leaf's tile (stride doubles per centre, value at +off); U = the unique-node table [nCenters][K] used when the
// stencil holds unique ids (AMR)
template <int K>
__device__ __forceinline__ void background_gather(int kind, const BgStencil8 &s8, const BgStencil &st, const double *__restrict__ T, int stride, int off,
                                                  const double *__restrict__ U, double out[K]) {
#pragma unroll
  for (int q = 0; q < K; q++) out[q] = 0.0;
  if (kind == 1) {
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const double *t = T + (size_t)stride * (s8.nd0 + ((s >> 2) & 1) + ((s >> 1) & 1) * s8.BS0 + (s & 1) * s8.BS1) + off;
      const double ws = s8.w[s];
      if (ws != 0.0) {  // slots without weight are skipped like the cells the reference never adds
#pragma unroll
        for (int q = 0; q < K; q++) out[q] += ws * t[q];
      }
    }
  } else {
    for (int s = 0; s < st.n; s++) {
      const double *t = U + (size_t)K * st.nd[s];
      for (int q = 0; q < K; q++) out[q] += st.w[s] * t[q];
    }
  }
}

__global__ void __launch_bounds__(128) magnetic_moment_init_kernel(DevMesh m, DevSpecies sp, int interp, CplrUnique U, double SpeedOfLight, ParticleSoA p,
                                                                  const int *__restrict__ nSlots, const double *__restrict__ bgTile,
                                                                  DevMoveStats *__restrict__ stats) {
  const int n = *nSlots;
  const int C = m.cellsPerBlock;
  unsigned int nErr = 0;
  for (int ip = blockIdx.x * blockDim.x + threadIdx.x; ip < n; ip += gridDim.x * blockDim.x) {
    const int key = p.key[ip];
    if (key < 0) continue;
    const double x[3] = {p.x[0][ip], p.x[1][ip], p.x[2][ip]};
    const double v[3] = {p.v[0][ip], p.v[1][ip], p.v[2][ip]};
    const int leaf = key / C;
    BgStencil st;
    BgStencil8 s8;
    const int kind = background_stencil(m, interp, x, leaf, s8, st);
    if (!kind) {
      nErr++;
      continue;
    }
    const double *T = bgTile + (size_t)leaf * m.nCenterLocal * 6;
    double B[3], E[3];
    background_gather<3>(kind, s8, st, T, 6, 3, U.B, B);
    background_gather<3>(kind, s8, st, T, 6, 0, U.E, E);
    const double AbsB = sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]) + 1E-15;
    double vE[3];
    vE[0] = E[1] * B[2] - E[2] * B[1];
    vE[1] = E[2] * B[0] - E[0] * B[2];
    vE[2] = E[0] * B[1] - E[1] * B[0];
    if (AbsB > 0.0)
      for (int d = 0; d < 3; d++) vE[d] /= AbsB * AbsB;
    const double vE_norm = sqrt(vE[0] * vE[0] + vE[1] * vE[1] + vE[2] * vE[2]);
    const double v_norm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const double c2 = SpeedOfLight * SpeedOfLight;
    const double kappa = 1 / sqrt(1 - vE_norm * vE_norm / (c2));
    const double gamma = 1 / sqrt(1 - v_norm * v_norm / (c2));
    double mu = 0.0;
    double B_star[3], v_star[3];
    const double gamma_star = gamma / kappa;
    double vE_cross_E[3];
    const double B_dot_vE = B[0] * vE[0] + B[1] * vE[1] + B[2] * vE[2];
    vE_cross_E[0] = vE[1] * E[2] - vE[2] * E[1];
    vE_cross_E[1] = vE[2] * E[0] - vE[0] * E[2];
    vE_cross_E[2] = vE[0] * E[1] - vE[1] * E[0];
    for (int d = 0; d < 3; d++) {
      B_star[d] = kappa * (B[d] - vE_cross_E[d] / c2);
      if (vE_norm > 0) B_star[d] -= (kappa - 1) * B_dot_vE * vE[d] / (vE_norm * vE_norm);
      v_star[d] = v[d] - vE[d];
    }
    const double Bstar_norm = sqrt(B_star[0] * B_star[0] + B_star[1] * B_star[1] + B_star[2] * B_star[2]);
    if (Bstar_norm > 0.0) {
      const double vstar_norm = sqrt(v_star[0] * v_star[0] + v_star[1] * v_star[1] + v_star[2] * v_star[2]);
      const double vstar_par = (v_star[0] * B_star[0] + v_star[1] * B_star[1] + v_star[2] * B_star[2]) / Bstar_norm;
      const double m0 = sp.mass[p.spec[ip] & 0x3f];
      mu = 0.5 * (gamma_star * gamma_star) * m0 * (vstar_norm * vstar_norm - vstar_par * vstar_par) / Bstar_norm;
    }
    p.mu[ip] = mu;
  }
  flush_move_counters(stats, 0, 0, 0, 0, 0, 0, nErr);
}
void launch_magnetic_moment_init(const DevMesh &m, const DevSpecies &sp, int interp, double c, ParticleSoA p, const int *nSlots, long long nUpper,
                                 const double *bgTile, const double *uE, const double *uB, DevMoveStats *stats, cudaStream_t s) {
  long long g = (nUpper + 127) / 128;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  CplrUnique U;
  U.E = uE, U.B = uB;
  magnetic_moment_init_kernel<<<(int)g, 128, 0, s>>>(m, sp, interp, U, c, p, nSlots, bgTile, stats);
}

// target = p.mu or p.vpar
__global__ void magnetic_moment_set_kernel(ParticleSoA p, double *__restrict__ target, const int *__restrict__ nSlots,
                                           const double *__restrict__ muByPtr, long long nMu) {
  const int n = *nSlots;
  for (int ip = blockIdx.x * blockDim.x + threadIdx.x; ip < n; ip += gridDim.x * blockDim.x) {
    const int pt = p.ptr[ip];
    if (pt >= 0 && pt < nMu) target[ip] = muByPtr[pt];
  }
}
void launch_magnetic_moment_set(ParticleSoA p, double *target, const int *nSlots, long long nUpper, const double *muByPtr, long long nMu,
                                cudaStream_t s) {
  long long g = (nUpper + 255) / 256;
  if (g < 1) g = 1;
  if (g > 148 * 16) g = 148 * 16;
  magnetic_moment_set_kernel<<<(int)g, 256, 0, s>>>(p, target, nSlots, muByPtr, nMu);
}

// (no minimum of resident CTAs: with 3, 4 or 5 per SM in the launch bounds the kernel measured 6 % slower / no faster: its 3 200
// instructions per particle of IEEE divisions and gathers do not gain from occupancy, profiles/r2_gca_ncu_summary.txt)
__global__ void __launch_bounds__(128) move_relativistic_gca_kernel(DevMesh m, DevSpecies sp, TpParams tp, ParticleSoA p, const int *__restrict__ nSlots,
                                                                   const double *__restrict__ bgTile, const double *__restrict__ gcaTile,
                                                                   const double *__restrict__ uVar, int *__restrict__ cellCount, DevMoveStats *__restrict__ stats,
                                                                   amps_gpu_exit_record *__restrict__ exitBuf, unsigned long long *__restrict__ exitCount) {
  const int n = *nSlots;
  const int C = m.cellsPerBlock;
  unsigned int nMoved = 0, nXCell = 0, nXBlock = 0, nLeft = 0, nNotUsed = 0, nWrap = 0, nErr = 0;
  const double c2 = tp.c * tp.c;

  for (int ip = blockIdx.x * blockDim.x + threadIdx.x; ip < n; ip += gridDim.x * blockDim.x) {
    const int oldKey = p.key[ip];
    if (oldKey < 0) continue;
    nMoved++;
    const double xInit[3] = {p.x[0][ip], p.x[1][ip], p.x[2][ip]};
    const double vInit[3] = {p.v[0][ip], p.v[1][ip], p.v[2][ip]};
    double xFinal[3], vFinal[3] = {0.0, 0.0, 0.0};
    const int spec = p.spec[ip] & 0x3f;
    const int startLeaf = oldKey / C;
    const int startNode = m.leaf[startLeaf].node;
    const double ElectricCharge = sp.charge[spec], mass = sp.mass[spec];
    const double dtTotal = (sp.timeStepMode == AMPS_DT_SPECIES_GLOBAL) ? sp.dt[spec] : sp.dt[0];
    int outcome = 0, node = -1;
    const double Mr = p.mu[ip];
    double uPar = 0.0;
    double bHat[3] = {0.0, 0.0, 0.0};

    {
      const double vNorm = sqrt(vInit[0] * vInit[0] + vInit[1] * vInit[1] + vInit[2] * vInit[2]);
      const double lfac = 1 / sqrt(1.0 - vNorm * vNorm / c2);
      BgStencil st;
      BgStencil8 s8;
      const int kind = background_stencil(m, tp.interp, xInit, startLeaf, s8, st);
      if (!kind) outcome = 3;
      else {
        double B[3], E[3], var15[15];
        background_gather<3>(kind, s8, st, bgTile + (size_t)startLeaf * m.nCenterLocal * 6, 6, 3, tp.U.B, B);
        background_gather<3>(kind, s8, st, bgTile + (size_t)startLeaf * m.nCenterLocal * 6, 6, 0, tp.U.E, E);
        background_gather<15>(kind, s8, st, gcaTile + (size_t)startLeaf * m.nCenterLocal * 15, 15, 0, uVar, var15);
        const double *b_dot_grad_b = var15, *vE_dot_grad_b = var15 + 3, *b_dot_grad_vE = var15 + 6, *vE_dot_grad_vE = var15 + 9, *grad_kappaB = var15 + 12;
        const double bNorm = sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
        double ePar = 0.0;
        if (bNorm > 0.0) {
          for (int d = 0; d < 3; d++) {
            bHat[d] = B[d] / bNorm;
            ePar += E[d] * bHat[d];
          }
        }
        const double vPar = vInit[0] * bHat[0] + vInit[1] * bHat[1] + vInit[2] * bHat[2];
        uPar = lfac * vPar;
        double vE[3], vENorm = 0.0;
        vE[0] = E[1] * bHat[2] - E[2] * bHat[1];
        vE[1] = E[2] * bHat[0] - E[0] * bHat[2];
        vE[2] = E[0] * bHat[1] - E[1] * bHat[0];
        if (bNorm > 0.0) {
          for (int d = 0; d < 3; d++) {
            vE[d] = vE[d] / bNorm;
            vENorm += vE[d] * vE[d];
          }
        }
        vENorm = sqrt(vENorm);
        const double kappa = 1 / sqrt(1 - vENorm * vENorm / c2);
        const double gamma = sqrt(1.0 + (uPar * uPar + 2.0 * Mr * bNorm / mass) / c2) * kappa;
        double utmp1[3] = {0.0, 0.0, 0.0}, utmp2[3], utmp3[3];
        double temp = bNorm / (kappa * kappa);
        if (bNorm > 0.0)
          for (int d = 0; d < 3; d++) utmp1[d] = bHat[d] / temp;
        for (int d = 0; d < 3; d++) {
          utmp2[d] = Mr / (gamma * ElectricCharge) * grad_kappaB[d] +
                     mass / ElectricCharge * (uPar * uPar / gamma * b_dot_grad_b[d] + uPar * vE_dot_grad_b[d] + uPar * b_dot_grad_vE[d] + gamma * vE_dot_grad_vE[d]);
        }
        for (int d = 0; d < 3; d++) utmp2[d] = utmp2[d] + uPar * ePar / (gamma)*vE[d];
        double u[3];
        utmp3[0] = utmp1[1] * utmp2[2] - utmp1[2] * utmp2[1];
        utmp3[1] = utmp1[2] * utmp2[0] - utmp1[0] * utmp2[2];
        utmp3[2] = utmp1[0] * utmp2[1] - utmp1[1] * utmp2[0];
        for (int d = 0; d < 3; d++) {
          u[d] = vE[d] + utmp3[d];
          u[d] += uPar / gamma * bHat[d];
        }
        double dupardt = ElectricCharge / mass * ePar;
        temp = Mr / (mass * gamma);
        for (int d = 0; d < 3; d++) {
          dupardt += -temp * bHat[d] * grad_kappaB[d] + vE[d] * (uPar * b_dot_grad_b[d] + gamma * vE_dot_grad_b[d]);
          xFinal[d] = xInit[d] + dtTotal * u[d];
        }
        uPar += dupardt * dtTotal;
      }
    }

    if (outcome == 0 && tp.rSphere > 0.0) {
      const double rFinal = sqrt(xFinal[0] * xFinal[0] + xFinal[1] * xFinal[1] + xFinal[2] * xFinal[2]);
      if (rFinal < tp.rSphere) {
        add_exit_record(exitBuf, exitCount, tp.exitCap, p.ptr[ip], spec, AMPS_EXIT_SPHERE, startLeaf, xInit, vInit);
        outcome = 1;
      }
    }
    int newLeaf = -1;
    if (outcome == 0) {
      node = find_tree_node_plain(m, xFinal, startNode);
      if (node < 0) outcome = 1;  // DELETE
      else if ((newLeaf = m.nodeLeaf[node]) < 0) outcome = 3;  // "the block is empty" :348
    }
    if (outcome == 0) {
      BgStencil st;
      BgStencil8 s8;
      const int kind = background_stencil(m, tp.interp, xFinal, newLeaf, s8, st);
      if (!kind) outcome = 3;
      else {
        double B[3], E[3];
        background_gather<3>(kind, s8, st, bgTile + (size_t)newLeaf * m.nCenterLocal * 6, 6, 3, tp.U.B, B);
        background_gather<3>(kind, s8, st, bgTile + (size_t)newLeaf * m.nCenterLocal * 6, 6, 0, tp.U.E, E);
        const double bNorm = sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
        if (bNorm > 0.0)  // otherwise bHat of the first stage stays, as in the reference
          for (int d = 0; d < 3; d++) bHat[d] = B[d] / bNorm;
        double vE[3], vENorm = 0.0;
        vE[0] = E[1] * bHat[2] - E[2] * bHat[1];
        vE[1] = E[2] * bHat[0] - E[0] * bHat[2];
        vE[2] = E[0] * bHat[1] - E[1] * bHat[0];
        if (bNorm > 0.0) {
          for (int d = 0; d < 3; d++) {
            vE[d] = vE[d] / bNorm;
            vENorm += vE[d] * vE[d];
          }
        }
        vENorm = sqrt(vENorm);
        const double kappa = 1 / sqrt(1 - vENorm * vENorm / c2);
        const double gamma = sqrt(1.0 + (uPar * uPar + 2.0 * Mr * bNorm / mass) / c2) * kappa;
        const double vPar = uPar / gamma;
        double res = (1 - 1 / (gamma * gamma)) * c2 - vPar * vPar;
        if (bNorm == 0 && res < 0) res = 0.0;
        if (bNorm > 0 && res < 0) outcome = 1;  // DeleteParticle, _PARTICLE_LEFT_THE_DOMAIN_ :308-311
        else {
          const double vPerp = sqrt(res);
          const double diff = (1.0 - bHat[0]) * (1.0 - bHat[0]) + (0.0 - bHat[1]) * (0.0 - bHat[1]) + (0.0 - bHat[2]) * (0.0 - bHat[2]);
          const double e0 = (diff > 0.0) ? 1.0 : 0.0, e1 = (diff > 0.0) ? 0.0 : 1.0, e2 = 0.0;
          double ePerp[3];
          ePerp[0] = e1 * bHat[2] - e2 * bHat[1];
          ePerp[1] = e2 * bHat[0] - e0 * bHat[2];
          ePerp[2] = e0 * bHat[1] - e1 * bHat[0];
          for (int d = 0; d < 3; d++) vFinal[d] += vPerp * ePerp[d] + vPar * bHat[d];
        }
      }
    }

    int newKey = -1;
    if (outcome == 0) {
      int ijk[3];
      if (!find_cell_index(m, xFinal, node, ijk)) outcome = 3;
      else {
        const int realLeaf = m.leaf[newLeaf].real;
        if (realLeaf >= 0) {
          const LeafGeo &gg = m.leaf[newLeaf];
          const LeafGeo &rg = m.leaf[realLeaf];
          for (int d = 0; d < 3; d++) {
            xFinal[d] += rg.xmin[d] - gg.xmin[d];
            if (xFinal[d] < rg.xmin[d]) xFinal[d] = rg.xmin[d];
            if (xFinal[d] >= rg.xmax[d]) xFinal[d] = rg.xmax[d] - 1.0E-10 * (rg.xmax[d] - rg.xmin[d]);
          }
          newLeaf = realLeaf;
          nWrap++;
        }
        newKey = newLeaf * C + ijk[0] + m.N[0] * (ijk[1] + m.N[1] * ijk[2]);
        if (newLeaf != startLeaf) nXBlock++;
        else if (newKey != oldKey) nXCell++;
      }
    }
    if (outcome == 1) nLeft++;
    else if (outcome == 2) nNotUsed++;
    else if (outcome == 3) nErr++;
    if (newKey >= 0) {
      p.x[0][ip] = xFinal[0], p.x[1][ip] = xFinal[1], p.x[2][ip] = xFinal[2];
      p.v[0][ip] = vFinal[0], p.v[1][ip] = vFinal[1], p.v[2][ip] = vFinal[2];
      atomicAdd(&cellCount[newKey], 1);
    }
    if (newKey != oldKey) p.key[ip] = newKey;
  }
  flush_move_counters(stats, nMoved, nXCell, nXBlock, nLeft, nNotUsed, nWrap, nErr);
}

void launch_move_relativistic_gca(const DevMesh &m, const DevSpecies &sp, int interp, double c, double rSphere, long long exitCap, ParticleSoA p,
                                  const int *nSlots, long long nUpper, const double *bgTile, const double *gcaTile, const double *uE, const double *uB,
                                  const double *uVar, int *cellCount, DevMoveStats *stats, amps_gpu_exit_record *exitBuf,
                                  unsigned long long *exitCount, cudaStream_t s) {
  TpParams tp;
  tp.interp = interp, tp.backward = 0, tp.boundaryMode = sp.boundaryMode, tp.c = c, tp.rSphere = rSphere, tp.exitCap = exitCap;
  tp.U.E = uE, tp.U.B = uB;
  long long g = (nUpper + 127) / 128;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  move_relativistic_gca_kernel<<<(int)g, 128, 0, s>>>(m, sp, tp, p, nSlots, bgTile, gcaTile, uVar, cellCount, stats, exitBuf, exitCount);
}

// ------------------------------------------------------------------------------------------------
// a8: PIC::Mover::GuidingCenter  src/pic/pic_mover_guiding_center.cpp  (coupler mode, relativity off)
//     InitiateMagneticMoment :85-144, GuidingCenterMotion_default :146-289, Mover_SecondOrder :292-619,
//     Mover_FirstOrder :622-849.  The reference writes |B| as pow(B.B,0.5); sqrt is used here (glibc's pow differs
//     from the correctly rounded root in ~0.1% of the arguments, so x and v agree with the CPU oracle to a few ulp
//     instead of bit for bit -- the parity tests state the tolerance).
// ------------------------------------------------------------------------------------------------
struct GcTables {
  const double *bg;     // [leaf][nCenterLocal][6]  E, B
  const double *gradB;  // [leaf][nCenterLocal][9]
  const double *uE, *uB, *uGradB;  // unique-node tables (AMR stencils)
  // cfg.gc_fields_ecsim: the ECSIM arrays instead (unique corners / centres), read through the corner and centre stencils
  const double *ecsimE = nullptr;  // [nCorners][3] current E
  const double *ecsimB = nullptr;  // [nCenters][3] B_cur
  int globalStencilFull = 0;
};

// ---- the ECSIM field getters of the guiding-centre movers (reference built with the ECSIM field solver) ----
// ECSIM::GetMagneticField, pic_field_solver_ecsim.cpp:7456-7469: CellCentered::Linear stencil (same-level branch: the trilinear stencil of
// the block's own centres incl. the ghost layer, centres outside the global box dropped, Normalize()) on B_cur
__device__ __forceinline__ bool ecsim_get_B(const DevMesh &m, const GcTables &T, const double x[3], int leaf, double B[3]) {
  const LeafGeo &lg = m.leaf[leaf];
  const double iLoc = (x[0] - lg.xmin[0]) / (lg.xmax[0] - lg.xmin[0]) * m.N[0];
  const double jLoc = (x[1] - lg.xmin[1]) / (lg.xmax[1] - lg.xmin[1]) * m.N[1];
  const double kLoc = (x[2] - lg.xmin[2]) / (lg.xmax[2] - lg.xmin[2]) * m.N[2];
  const int i0 = (iLoc < 0.5) ? -1 : (int)(iLoc - 0.50);
  const int j0 = (jLoc < 0.5) ? -1 : (int)(jLoc - 0.50);
  const int k0 = (kLoc < 0.5) ? -1 : (int)(kLoc - 0.50);
  B[0] = B[1] = B[2] = 0.0;
  if (i0 < -m.g[0] || i0 + 1 > m.N[0] + m.g[0] - 1 || j0 < -m.g[1] || j0 + 1 > m.N[1] + m.g[1] - 1 || k0 < -m.g[2] || k0 + 1 > m.N[2] + m.g[2] - 1)
    return false;  // outside the block's tile (out-of-bounds read in the reference)
  const double w0 = iLoc - (i0 + 0.5), w1 = jLoc - (j0 + 0.5), w2 = kLoc - (k0 + 0.5);
  double w[8];
  w[0] = (1.0 - w0) * (1.0 - w1) * (1.0 - w2);
  w[1] = (1.0 - w0) * (1.0 - w1) * w2;
  w[2] = (1.0 - w0) * w1 * (1.0 - w2);
  w[3] = (1.0 - w0) * w1 * w2;
  w[4] = w0 * (1.0 - w1) * (1.0 - w2);
  w[5] = w0 * (1.0 - w1) * w2;
  w[6] = w0 * w1 * (1.0 - w2);
  w[7] = w0 * w1 * w2;
  unsigned valid = 0xffu;
  if (!m.periodic && lg.face) {
    if ((lg.face & 1) && i0 < 0) valid &= 0xf0u;
    if ((lg.face & 2) && i0 + 1 >= m.N[0]) valid &= 0x0fu;
    if ((lg.face & 4) && j0 < 0) valid &= 0xccu;
    if ((lg.face & 8) && j0 + 1 >= m.N[1]) valid &= 0x33u;
    if ((lg.face & 16) && k0 < 0) valid &= 0xaau;
    if ((lg.face & 32) && k0 + 1 >= m.N[2]) valid &= 0x55u;
  }
  double norm = 0.0;
#pragma unroll
  for (int s = 0; s < 8; s++)
    if (valid & (1u << s)) norm += w[s];
  const int *cu = m.centerUid + (size_t)leaf * m.nCenterLocal;
  const bool skipNorm = T.globalStencilFull && valid == 0xffu;  // see DevSpecies::globalStencilFull
#pragma unroll
  for (int s = 0; s < 8; s++)
    if (valid & (1u << s)) {
      const double ws = skipNorm ? w[s] : w[s] / norm;
      const double *t = T.ecsimB + 3 * (size_t)cu[centerLocalNumber(m, i0 + ((s >> 2) & 1), j0 + ((s >> 1) & 1), k0 + (s & 1))];
      B[0] += ws * t[0], B[1] += ws * t[1], B[2] += ws * t[2];
    }
  return true;
}
// ECSIM::GetElectricField, :7440-7453: CornerBased::InitStencil (pic_interpolation_routines.cpp:1074-1194, Normalize()) on the current E
__device__ __forceinline__ bool ecsim_get_E(const DevMesh &m, const GcTables &T, const double x[3], int leaf, double E[3]) {
  const LeafGeo &lg = m.leaf[leaf];
  double xl[3];
  int iX[3];
  E[0] = E[1] = E[2] = 0.0;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double dx = (lg.xmax[d] - lg.xmin[d]) / m.N[d];
    double xs = x[d];
    if ((xs < lg.xmin[d]) || (xs > lg.xmax[d])) return false;
    if (fabs(xs - lg.xmax[d]) < 1e-10 * dx) xs = lg.xmax[d] - 1e-10 * dx;
    xl[d] = (xs - lg.xmin[d]) / dx;
    iX[d] = (int)(xl[d]);
    xl[d] -= iX[d];
  }
  double w[8], norm = 0.0;
  // stencil order of the reference: i, j, k loops, k fastest
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const int i = (s >> 2) & 1, j = (s >> 1) & 1, k = s & 1;
    w[s] = (i ? xl[0] : 1.0 - xl[0]) * (j ? xl[1] : 1.0 - xl[1]) * (k ? xl[2] : 1.0 - xl[2]);
    norm += w[s];
  }
  const int *cu = m.cornerUid + (size_t)leaf * m.nCornerLocal;
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const double ws = w[s] / norm;
    const double *t = T.ecsimE + 3 * (size_t)cu[cornerLocalNumber(m, iX[0] + ((s >> 2) & 1), iX[1] + ((s >> 1) & 1), iX[2] + (s & 1))];
    E[0] += ws * t[0], E[1] += ws * t[1], E[2] += ws * t[2];
  }
  return true;
}
// ECSIM::GetMagneticFieldGradient, :7473-7547: central differences over half a cell of the start block, one-sided where a probe leaves the domain
__device__ __forceinline__ bool ecsim_get_gradB(const DevMesh &m, const GcTables &T, const double x[3], int leaf, double gradB[9]) {
  const LeafGeo &lg = m.leaf[leaf];
  double B0[3];
  if (!ecsim_get_B(m, T, x, leaf, B0)) return false;
#pragma unroll 1
  for (int idim = 0; idim < 3; idim++) {
    double xp[3] = {x[0], x[1], x[2]}, xm[3] = {x[0], x[1], x[2]};
    const double dx = 0.5 * (lg.xmax[idim] - lg.xmin[idim]) / m.N[idim];
    xp[idim] += dx, xm[idim] -= dx;
    const int np = find_tree_node_plain(m, xp, lg.node), nm = find_tree_node_plain(m, xm, lg.node);
    const bool hasP = np >= 0, hasM = nm >= 0;
    double Bp[3], Bm[3];
    if (hasP && (m.nodeLeaf[np] < 0 || !ecsim_get_B(m, T, xp, m.nodeLeaf[np], Bp))) return false;
    if (hasM && (m.nodeLeaf[nm] < 0 || !ecsim_get_B(m, T, xm, m.nodeLeaf[nm], Bm))) return false;
    for (int c = 0; c < 3; c++) {
      double g = 0.0;
      if (hasP && hasM) g = (Bp[c] - Bm[c]) / (2.0 * dx);
      else if (hasP) g = (Bp[c] - B0[c]) / dx;
      else if (hasM) g = (B0[c] - Bm[c]) / dx;
      gradB[3 * c + idim] = g;
    }
  }
  return true;
}

__device__ __forceinline__ bool gc_initiate_magnetic_moment(const DevMesh &m, const DevSpecies &sp, int interp, const GcTables &T, int spec,
                                                            const double x[3], double v[3], int leaf, double &muOut) {
  double B[3];
  if (T.ecsimB != nullptr) {  // pic_mover_guiding_center.cpp:103-104
    if (!ecsim_get_B(m, T, x, leaf, B)) return false;
  } else {
    BgStencil st;
    BgStencil8 s8;
    const int kind = background_stencil(m, interp, x, leaf, s8, st);
    if (!kind) return false;
    background_gather<3>(kind, s8, st, T.bg + (size_t)leaf * m.nCenterLocal * 6, 6, 3, T.uB, B);
  }
  const double AbsB = sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]) + 1E-15;
  double v_par = 0.0, mu = 0.0;
  const double b[3] = {B[0] / AbsB, B[1] / AbsB, B[2] / AbsB};
  if (AbsB > 0.0) {
    v_par = v[0] * b[0] + v[1] * b[1] + v[2] * b[2];
    const double v2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double gamma2 = 1.0;
    const double m0 = sp.mass[spec];
    mu = 0.5 * gamma2 * m0 * (v2 - v_par * v_par) / AbsB;
  }
  v[0] = v_par * b[0];
  v[1] = v_par * b[1];
  v[2] = v_par * b[2];
  muOut = mu;
  return true;
}

__device__ __forceinline__ bool gc_motion(const DevMesh &m, const DevSpecies &sp, int interp, int idealMhd, const GcTables &T, double Vguide_perp[3],
                                          double &ForceParal, double &AbsBOut, double bOut[3], const double *PParal, int spec, double mu,
                                          const double x[3], const double v[3], int leaf) {
  double E[3], B[3], gradB[9];
  if (T.ecsimB != nullptr) {  // pic_mover_guiding_center.cpp:179-184
    if (!ecsim_get_B(m, T, x, leaf, B) || !ecsim_get_E(m, T, x, leaf, E) || !ecsim_get_gradB(m, T, x, leaf, gradB)) return false;
  } else {
    BgStencil st;
    BgStencil8 s8;
    const int kind = background_stencil(m, interp, x, leaf, s8, st);
    if (!kind) return false;
    const double *tb = T.bg + (size_t)leaf * m.nCenterLocal * 6;
    background_gather<3>(kind, s8, st, tb, 6, 0, T.uE, E);
    background_gather<3>(kind, s8, st, tb, 6, 3, T.uB, B);
    background_gather<9>(kind, s8, st, T.gradB + (size_t)leaf * m.nCenterLocal * 9, 9, 0, T.uGradB, gradB);
  }
  const double AbsB = sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]) + 1E-15;
  double b[3], gradAbsB[3];
  b[0] = B[0] / AbsB;
  b[1] = B[1] / AbsB;
  b[2] = B[2] / AbsB;
  gradAbsB[0] = b[0] * gradB[0] + b[1] * gradB[3] + b[2] * gradB[6];
  gradAbsB[1] = b[0] * gradB[1] + b[1] * gradB[4] + b[2] * gradB[7];
  gradAbsB[2] = b[0] * gradB[2] + b[1] * gradB[5] + b[2] * gradB[8];
  const double q = sp.charge[spec], m0 = sp.mass[spec], gamma = 1.0;
  const double p_par = (PParal == nullptr) ? gamma * m0 * (v[0] * b[0] + v[1] * b[1] + v[2] * b[2]) : *PParal;
  double msc, vec[3];
  double V[3] = {0.0, 0.0, 0.0};
  V[0] += (E[1] * b[2] - E[2] * b[1]) / AbsB;
  V[1] += (E[2] * b[0] - E[0] * b[2]) / AbsB;
  V[2] += (E[0] * b[1] - E[1] * b[0]) / AbsB;
  msc = mu / (q * gamma) / AbsB;
  V[0] += msc * (b[1] * gradAbsB[2] - b[2] * gradAbsB[1]);
  V[1] += msc * (b[2] * gradAbsB[0] - b[0] * gradAbsB[2]);
  V[2] += msc * (b[0] * gradAbsB[1] - b[1] * gradAbsB[0]);
  msc = p_par * p_par / (q * gamma * m0) / AbsB / AbsB;
  vec[0] = b[0] * gradB[0] + b[1] * gradB[1] + b[2] * gradB[2];
  vec[1] = b[0] * gradB[3] + b[1] * gradB[4] + b[2] * gradB[5];
  vec[2] = b[0] * gradB[6] + b[1] * gradB[7] + b[2] * gradB[8];
  V[0] += msc * (b[1] * vec[2] - b[2] * vec[1]);
  V[1] += msc * (b[2] * vec[0] - b[0] * vec[2]);
  V[2] += msc * (b[0] * vec[1] - b[1] * vec[0]);
  if (idealMhd) ForceParal = -mu / gamma * (gradAbsB[0] * b[0] + gradAbsB[1] * b[1] + gradAbsB[2] * b[2]);
  else ForceParal = q * (E[0] * b[0] + E[1] * b[1] + E[2] * b[2]) - mu / gamma * (gradAbsB[0] * b[0] + gradAbsB[1] * b[1] + gradAbsB[2] * b[2]);
  for (int d = 0; d < 3; d++) Vguide_perp[d] = V[d], bOut[d] = b[d];
  AbsBOut = AbsB;
  return true;
}

// GuidingCenter::InitiateMagneticMoment for every resident particle (InitiateParticle, pic_pbuffer.cpp:988): mu and v || B
__global__ void __launch_bounds__(128) gc_magnetic_moment_init_kernel(DevMesh m, DevSpecies sp, int interp, GcTables T, ParticleSoA p,
                                                                     const int *__restrict__ nSlots, DevMoveStats *__restrict__ stats) {
  const int n = *nSlots;
  unsigned int nErr = 0;
  for (int ip = blockIdx.x * blockDim.x + threadIdx.x; ip < n; ip += gridDim.x * blockDim.x) {
    const int key = p.key[ip];
    if (key < 0) continue;
    const double x[3] = {p.x[0][ip], p.x[1][ip], p.x[2][ip]};
    double v[3] = {p.v[0][ip], p.v[1][ip], p.v[2][ip]};
    double mu;
    if (!gc_initiate_magnetic_moment(m, sp, interp, T, p.spec[ip] & 0x3f, x, v, key / m.cellsPerBlock, mu)) {
      nErr++;
      continue;
    }
    p.mu[ip] = mu;
    p.v[0][ip] = v[0], p.v[1][ip] = v[1], p.v[2][ip] = v[2];
  }
  flush_move_counters(stats, 0, 0, 0, 0, 0, 0, nErr);
}
void launch_gc_magnetic_moment_init(const DevMesh &m, const DevSpecies &sp, int interp, ParticleSoA p, const int *nSlots, long long nUpper,
                                    const double *bgTile, const double *uE, const double *uB, DevMoveStats *stats, cudaStream_t s,
                                    const double *ecsimE, const double *ecsimB) {
  long long g = (nUpper + 127) / 128;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  GcTables T;
  T.bg = bgTile, T.gradB = nullptr, T.uE = uE, T.uB = uB, T.uGradB = nullptr;
  T.ecsimE = ecsimE, T.ecsimB = ecsimB, T.globalStencilFull = sp.globalStencilFull;
  gc_magnetic_moment_init_kernel<<<(int)g, 128, 0, s>>>(m, sp, interp, T, p, nSlots, stats);
}

template <bool kSecondOrder>
__global__ void __launch_bounds__(128) move_guiding_center_kernel(DevMesh m, DevSpecies sp, TpParams tp, int idealMhd, GcTables T, ParticleSoA p,
                                                                 const int *__restrict__ nSlots, int *__restrict__ cellCount,
                                                                 DevMoveStats *__restrict__ stats, amps_gpu_exit_record *__restrict__ exitBuf,
                                                                 unsigned long long *__restrict__ exitCount) {
  const int n = *nSlots;
  const int C = m.cellsPerBlock;
  unsigned int nMoved = 0, nXCell = 0, nXBlock = 0, nLeft = 0, nNotUsed = 0, nWrap = 0, nErr = 0;

  for (int ip = blockIdx.x * blockDim.x + threadIdx.x; ip < n; ip += gridDim.x * blockDim.x) {
    const int oldKey = p.key[ip];
    if (oldKey < 0) continue;
    nMoved++;
    double x[3] = {p.x[0][ip], p.x[1][ip], p.x[2][ip]};
    double v[3] = {p.v[0][ip], p.v[1][ip], p.v[2][ip]};
    const unsigned char specByte = p.spec[ip];
    const int spec = specByte & 0x3f;
    const int startLeaf = oldKey / C;
    const int startNode = m.leaf[startLeaf].node;
    const double m0 = sp.mass[spec];
    const double dtTotal = (sp.timeStepMode == AMPS_DT_SPECIES_GLOBAL) ? sp.dt[spec] : sp.dt[0];
    int outcome = 0, node = -1;
    double xFinal[3], vFinal[3];
    double mu = p.mu[ip];

    if (!kSecondOrder) {
      if (!(specByte & 0x40)) {  // TestInitFlag == false (:640-644)
        p.spec[ip] = specByte | 0x40;
        if (!gc_initiate_magnetic_moment(m, sp, tp.interp, T, spec, x, v, startLeaf, mu)) outcome = 3;
        else p.mu[ip] = mu;
      }
      double Vg[3], Fpar = 0.0, AbsBInit, bInit[3], pp = 0.0;
      if (outcome == 0 && !gc_motion(m, sp, tp.interp, idealMhd, T, Vg, Fpar, AbsBInit, bInit, nullptr, spec, mu, x, v, startLeaf)) outcome = 3;
      if (outcome == 0) {
        double misc = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (v[0] * bInit[0] + v[1] * bInit[1] + v[2] * bInit[2] < 0.0) misc *= -1.0;
        for (int d = 0; d < 3; d++) v[d] = misc * bInit[d];
        pp = m0 * (v[0] * bInit[0] + v[1] * bInit[1] + v[2] * bInit[2]);
        for (int d = 0; d < 3; d++) x[d] += dtTotal * (Vg[d] + v[d]);
        pp += dtTotal * Fpar;
        node = find_tree_node_plain(m, x, -1);  // FindBlock
        if (node < 0) outcome = 1;
      }
      if (outcome == 0) {
        double bFinal[3];
        bool got;
        if (T.ecsimB != nullptr) {  // the NEW block in the ECSIM branch (:727-729)
          got = m.nodeLeaf[node] >= 0 && ecsim_get_B(m, T, x, m.nodeLeaf[node], bFinal);
        } else {
          BgStencil st;
          BgStencil8 s8;
          const int kind = background_stencil(m, tp.interp, x, startLeaf, s8, st);  // the START block, as written (:713)
          got = kind != 0;
          if (got) background_gather<3>(kind, s8, st, T.bg + (size_t)startLeaf * m.nCenterLocal * 6, 6, 3, T.uB, bFinal);
        }
        if (!got) outcome = 3;
        else {
          const double l0 = sqrt(bFinal[0] * bFinal[0] + bFinal[1] * bFinal[1] + bFinal[2] * bFinal[2]);
          if (l0 > 0.0) {
            const double l = 1.0 / l0;
            for (int d = 0; d < 3; d++) bFinal[d] *= l;
          }
          const double misc = pp / m0;
          for (int d = 0; d < 3; d++) v[d] = misc * bFinal[d];
          if (tp.rSphere > 0.0 && x[0] * x[0] + x[1] * x[1] + x[2] * x[2] < tp.rSphere * tp.rSphere) outcome = 1;  // no callback (:748-758)
          else {
            node = find_tree_node_plain(m, x, startNode);
            if (node < 0) outcome = 3;
          }
        }
      }
      for (int d = 0; d < 3; d++) xFinal[d] = x[d], vFinal[d] = v[d];
    } else {
      double VgInit[3], FparInit = 0.0, AbsBInit, bInit[3];
      if (!gc_motion(m, sp, tp.interp, idealMhd, T, VgInit, FparInit, AbsBInit, bInit, nullptr, spec, mu, x, v, startLeaf)) outcome = 3;
      double pInit = 0.0, pMiddle = 0.0, xMiddle[3];
      if (outcome == 0) {
        double misc = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (v[0] * bInit[0] + v[1] * bInit[1] + v[2] * bInit[2] < 0) misc *= -1.0;
        v[0] = misc * bInit[0];
        v[1] = misc * bInit[1];
        v[2] = misc * bInit[2];
        pInit = m0 * (v[0] * bInit[0] + v[1] * bInit[1] + v[2] * bInit[2]);
        const double dtTemp = dtTotal / 2.0;
        for (int d = 0; d < 3; d++) xMiddle[d] = x[d] + dtTemp * (VgInit[d] + v[d]);
        pMiddle = pInit + dtTemp * FparInit;
        node = find_tree_node_plain(m, xMiddle, -1);  // FindBlock
        if (node < 0) outcome = 1;
        else if (m.nodeLeaf[node] < 0) outcome = 3;
      }
      if (outcome == 0) {
        double VgMid[3], FparMid = 0.0, AbsBMid, bMid[3];
        const double vMiddle0[3] = {0.0, 0.0, 0.0};
        if (!gc_motion(m, sp, tp.interp, idealMhd, T, VgMid, FparMid, AbsBMid, bMid, &pMiddle, spec, mu, xMiddle, vMiddle0, m.nodeLeaf[node])) outcome = 3;
        else {
          double misc = pMiddle / m0;
          double vMiddle[3];
          for (int d = 0; d < 3; d++) vMiddle[d] = misc * bMid[d];
          for (int d = 0; d < 3; d++) xFinal[d] = x[d] + dtTotal * (VgMid[d] + vMiddle[d]);
          const double pFinal = pInit + dtTotal * FparMid;
          misc = pFinal / m0;
          for (int d = 0; d < 3; d++) vFinal[d] = misc * bMid[d];
          bool hit = false;
          if (tp.rSphere > 0.0) {
            const double rFinal2 = xFinal[0] * xFinal[0] + xFinal[1] * xFinal[1] + xFinal[2] * xFinal[2];
            if (rFinal2 < tp.rSphere * tp.rSphere) {
              const double r = sqrt(rFinal2);
              for (int d = 0; d < 3; d++) xFinal[d] *= tp.rSphere / r;
              const int nn = find_tree_node_plain(m, xFinal, startNode);
              add_exit_record(exitBuf, exitCount, tp.exitCap, p.ptr[ip], spec, AMPS_EXIT_SPHERE, nn >= 0 ? m.nodeLeaf[nn] : -1, xFinal, vFinal);
              outcome = 1, hit = true;
            }
          }
          if (!hit) {
            node = find_tree_node_plain(m, xFinal, startNode);
            if (node < 0) outcome = 1;
          }
        }
      }
    }

    int newKey = -1;
    if (outcome == 0) {
      int ijk[3];
      int newLeaf = m.nodeLeaf[node];
      if (!find_cell_index(m, xFinal, node, ijk) || newLeaf < 0) outcome = 3;
      else {
        const int realLeaf = m.leaf[newLeaf].real;
        if (realLeaf >= 0) {
          const LeafGeo &gg = m.leaf[newLeaf];
          const LeafGeo &rg = m.leaf[realLeaf];
          for (int d = 0; d < 3; d++) {
            xFinal[d] += rg.xmin[d] - gg.xmin[d];
            if (xFinal[d] < rg.xmin[d]) xFinal[d] = rg.xmin[d];
            if (xFinal[d] >= rg.xmax[d]) xFinal[d] = rg.xmax[d] - 1.0E-10 * (rg.xmax[d] - rg.xmin[d]);
          }
          newLeaf = realLeaf;
          nWrap++;
        }
        newKey = newLeaf * C + ijk[0] + m.N[0] * (ijk[1] + m.N[1] * ijk[2]);
        if (newLeaf != startLeaf) nXBlock++;
        else if (newKey != oldKey) nXCell++;
      }
    }
    if (outcome == 1) nLeft++;
    else if (outcome == 2) nNotUsed++;
    else if (outcome == 3) nErr++;
    if (newKey >= 0) {
      p.x[0][ip] = xFinal[0], p.x[1][ip] = xFinal[1], p.x[2][ip] = xFinal[2];
      p.v[0][ip] = vFinal[0], p.v[1][ip] = vFinal[1], p.v[2][ip] = vFinal[2];
      atomicAdd(&cellCount[newKey], 1);
    }
    if (newKey != oldKey) p.key[ip] = newKey;
  }
  flush_move_counters(stats, nMoved, nXCell, nXBlock, nLeft, nNotUsed, nWrap, nErr);
}

void launch_move_guiding_center(const DevMesh &m, const DevSpecies &sp, int order, int interp, int idealMhd, double rSphere, long long exitCap,
                                ParticleSoA p, const int *nSlots, long long nUpper, const double *bgTile, const double *gradBTile, const double *uE,
                                const double *uB, const double *uGradB, int *cellCount, DevMoveStats *stats, amps_gpu_exit_record *exitBuf,
                                unsigned long long *exitCount, cudaStream_t s, const double *ecsimE, const double *ecsimB) {
  TpParams tp;
  tp.interp = interp, tp.backward = 0, tp.boundaryMode = sp.boundaryMode, tp.c = 0.0, tp.rSphere = rSphere, tp.exitCap = exitCap;
  GcTables T;
  T.bg = bgTile, T.gradB = gradBTile, T.uE = uE, T.uB = uB, T.uGradB = uGradB;
  T.ecsimE = ecsimE, T.ecsimB = ecsimB, T.globalStencilFull = sp.globalStencilFull;
  long long g = (nUpper + 127) / 128;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  if (order == 2) move_guiding_center_kernel<true><<<(int)g, 128, 0, s>>>(m, sp, tp, idealMhd, T, p, nSlots, cellCount, stats, exitBuf, exitCount);
  else move_guiding_center_kernel<false><<<(int)g, 128, 0, s>>>(m, sp, tp, idealMhd, T, p, nSlots, cellCount, stats, exitBuf, exitCount);
}


// ------------------------------------------------------------------------------------------------
// f2: PIC::GYROKINETIC::Mover_FirstOrder / Mover_SecondOrder  src/pic/gyro/gyro_mover.cpp:383-720 (coupler fields)
// reduced state (x, v_parallel, mu); v = b v_parallel + v_drift is rebuilt at the final position (CommitReducedStateAndVelocity)
// ------------------------------------------------------------------------------------------------
// EvalRHS, :245-336; false where the reference reads outside the block's table
__device__ __forceinline__ bool gyro_eval_rhs(const DevMesh &m, const DevSpecies &sp, int interp, const GcTables &T, const double x[3], int leaf,
                                              double vpar, double mu, int spec, double &absB, double b[3], double vdrift[3], double &dvpar_dt) {
  absB = 0.0;
  b[0] = 0.0, b[1] = 0.0, b[2] = 0.0;
  vdrift[0] = 0.0, vdrift[1] = 0.0, vdrift[2] = 0.0;
  dvpar_dt = 0.0;
  BgStencil st;
  BgStencil8 s8;
  const int kind = background_stencil(m, interp, x, leaf, s8, st);
  if (!kind) return false;
  double E[3], B[3], gradB[9];
  const double *tb = T.bg + (size_t)leaf * m.nCenterLocal * 6;
  background_gather<3>(kind, s8, st, tb, 6, 0, T.uE, E);
  background_gather<3>(kind, s8, st, tb, 6, 3, T.uB, B);
  background_gather<9>(kind, s8, st, T.gradB + (size_t)leaf * m.nCenterLocal * 9, 9, 0, T.uGradB, gradB);
  const double absB2_ = B[0] * B[0] + B[1] * B[1] + B[2] * B[2];
  if (absB2_ <= 0.0) return true;
  absB = sqrt(absB2_);
  const double inv = 1.0 / absB;
  b[0] = B[0] * inv, b[1] = B[1] * inv, b[2] = B[2] * inv;
  const double mm = sp.mass[spec], q = sp.charge[spec];
  const double absB2 = absB * absB;
  const double invAbsB2 = 1.0 / absB2;
  double gradAbsB[3];
  {
    const double invAbsB = 1.0 / absB;
    for (int j = 0; j < 3; j++) gradAbsB[j] = (B[0] * gradB[0 * 3 + j] + B[1] * gradB[1 * 3 + j] + B[2] * gradB[2 * 3 + j]) * invAbsB;
  }
  const double Epar = E[0] * b[0] + E[1] * b[1] + E[2] * b[2];
  const double bDotGradAbsB = b[0] * gradAbsB[0] + b[1] * gradAbsB[1] + b[2] * gradAbsB[2];
  dvpar_dt = (q / mm) * Epar - (mu / mm) * bDotGradAbsB;
  const double ExB[3] = {E[1] * B[2] - E[2] * B[1], E[2] * B[0] - E[0] * B[2], E[0] * B[1] - E[1] * B[0]};
  vdrift[0] = ExB[0] * invAbsB2;
  vdrift[1] = ExB[1] * invAbsB2;
  vdrift[2] = ExB[2] * invAbsB2;
  if (q != 0.0 && mu != 0.0) {
    const double BxG[3] = {B[1] * gradAbsB[2] - B[2] * gradAbsB[1], B[2] * gradAbsB[0] - B[0] * gradAbsB[2], B[0] * gradAbsB[1] - B[1] * gradAbsB[0]};
    const double c = (mu / q) * invAbsB2;
    vdrift[0] += c * BxG[0];
    vdrift[1] += c * BxG[1];
    vdrift[2] += c * BxG[2];
  }
  double BB[3];
  BB[0] = B[0] * gradB[0 * 3 + 0] + B[1] * gradB[0 * 3 + 1] + B[2] * gradB[0 * 3 + 2];
  BB[1] = B[0] * gradB[1 * 3 + 0] + B[1] * gradB[1 * 3 + 1] + B[2] * gradB[1 * 3 + 2];
  BB[2] = B[0] * gradB[2 * 3 + 0] + B[1] * gradB[2 * 3 + 1] + B[2] * gradB[2 * 3 + 2];
  if (q != 0.0 && vpar != 0.0) {
    const double BxBB[3] = {B[1] * BB[2] - B[2] * BB[1], B[2] * BB[0] - B[0] * BB[2], B[0] * BB[1] - B[1] * BB[0]};
    const double invAbsB4 = 1.0 / (absB2 * absB2);
    const double c = (mm * vpar * vpar / q) * invAbsB4;
    vdrift[0] += c * BxBB[0];
    vdrift[1] += c * BxBB[1];
    vdrift[2] += c * BxBB[2];
  }
  if (!isfinite(vdrift[0]) || !isfinite(vdrift[1]) || !isfinite(vdrift[2]) || !isfinite(dvpar_dt)) {
    vdrift[0] = 0.0, vdrift[1] = 0.0, vdrift[2] = 0.0;
    dvpar_dt = 0.0;
  }
  return true;
}

template <bool kSecondOrder>
__global__ void __launch_bounds__(128) move_gyrokinetic_kernel(DevMesh m, DevSpecies sp, TpParams tp, GcTables T, ParticleSoA p,
                                                              const int *__restrict__ nSlots, int *__restrict__ cellCount,
                                                              DevMoveStats *__restrict__ stats) {
  const int n = *nSlots;
  const int C = m.cellsPerBlock;
  unsigned int nMoved = 0, nXCell = 0, nXBlock = 0, nLeft = 0, nNotUsed = 0, nWrap = 0, nErr = 0;
  for (int ip = blockIdx.x * blockDim.x + threadIdx.x; ip < n; ip += gridDim.x * blockDim.x) {
    const int oldKey = p.key[ip];
    if (oldKey < 0) continue;
    nMoved++;
    const double x0[3] = {p.x[0][ip], p.x[1][ip], p.x[2][ip]};
    const int spec = p.spec[ip] & 0x3f;
    const int startLeaf = oldKey / C;
    const int startNode = m.leaf[startLeaf].node;
    const double dtTotal = (sp.timeStepMode == AMPS_DT_SPECIES_GLOBAL) ? sp.dt[spec] : sp.dt[0];
    const double mu = p.mu[ip];
    const double vpar0 = p.vpar[ip];
    int outcome = 0, node = -1;
    double x[3], vparNew = vpar0, vFinal[3] = {0.0, 0.0, 0.0};

    double absB, b[3], vdrift[3], dvpar_dt;
    if (!gyro_eval_rhs(m, sp, tp.interp, T, x0, startLeaf, vpar0, mu, spec, absB, b, vdrift, dvpar_dt)) outcome = 3;
    if (outcome == 0) {
      if (!kSecondOrder) {
        for (int d = 0; d < 3; d++) x[d] = x0[d] + dtTotal * (vdrift[d] + b[d] * vpar0);
        vparNew = vpar0 + dtTotal * dvpar_dt;
      } else {
        double xHalf[3];
        for (int d = 0; d < 3; d++) xHalf[d] = x0[d] + 0.5 * dtTotal * (vdrift[d] + b[d] * vpar0);
        const double vparHalf = vpar0 + 0.5 * dtTotal * dvpar_dt;
        const int nodeHalf = find_tree_node_plain(m, xHalf, -1);  // FindBlock
        if (nodeHalf < 0) outcome = 1;
        else if (m.nodeLeaf[nodeHalf] < 0) outcome = 3;
        else {
          double absBH, bH[3], vdriftH[3], dvpar_dtH;
          if (!gyro_eval_rhs(m, sp, tp.interp, T, xHalf, m.nodeLeaf[nodeHalf], vparHalf, mu, spec, absBH, bH, vdriftH, dvpar_dtH)) outcome = 3;
          else {
            for (int d = 0; d < 3; d++) x[d] = x0[d] + dtTotal * (vdriftH[d] + bH[d] * vparHalf);
            vparNew = vpar0 + dtTotal * dvpar_dtH;
          }
        }
      }
    }
    if (outcome == 0) {
      node = find_tree_node_plain(m, x, -1);  // FindBlock
      if (node < 0) outcome = 1;
      else if (m.nodeLeaf[node] < 0) outcome = 3;
    }
    if (outcome == 0) {
      // EvalRHS at the final point for the stored drift; CommitReducedStateAndVelocity evaluates b there once more (same values)
      double absB1, b1[3], vdrift1[3], dummy;
      if (!gyro_eval_rhs(m, sp, tp.interp, T, x, m.nodeLeaf[node], vparNew, mu, spec, absB1, b1, vdrift1, dummy)) outcome = 3;
      else {
        for (int d = 0; d < 3; d++) vFinal[d] = b1[d] * vparNew + vdrift1[d];
        if (tp.rSphere > 0.0 && x[0] * x[0] + x[1] * x[1] + x[2] * x[2] < tp.rSphere * tp.rSphere) outcome = 1;
        else {
          node = find_tree_node_plain(m, x, startNode);
          if (node < 0) outcome = 3;
        }
      }
    }
    int newKey = -1;
    if (outcome == 0) {
      int ijk[3];
      int newLeaf = m.nodeLeaf[node];
      if (!find_cell_index(m, x, node, ijk) || newLeaf < 0) outcome = 3;
      else {
        const int realLeaf = m.leaf[newLeaf].real;
        if (realLeaf >= 0) {  // periodic ghost -> real (pic_bc_periodic.cpp:100-134)
          const LeafGeo &gg = m.leaf[newLeaf];
          const LeafGeo &rg = m.leaf[realLeaf];
          for (int d = 0; d < 3; d++) {
            x[d] += rg.xmin[d] - gg.xmin[d];
            if (x[d] < rg.xmin[d]) x[d] = rg.xmin[d];
            if (x[d] >= rg.xmax[d]) x[d] = rg.xmax[d] - 1.0E-10 * (rg.xmax[d] - rg.xmin[d]);
          }
          newLeaf = realLeaf;
          nWrap++;
        }
        newKey = newLeaf * C + ijk[0] + m.N[0] * (ijk[1] + m.N[1] * ijk[2]);
        if (newLeaf != startLeaf) nXBlock++;
        else if (newKey != oldKey) nXCell++;
      }
    }
    if (outcome == 1) nLeft++;
    else if (outcome == 2) nNotUsed++;
    else if (outcome == 3) nErr++;
    if (newKey >= 0) {
      p.x[0][ip] = x[0], p.x[1][ip] = x[1], p.x[2][ip] = x[2];
      p.v[0][ip] = vFinal[0], p.v[1][ip] = vFinal[1], p.v[2][ip] = vFinal[2];
      p.vpar[ip] = vparNew;
      atomicAdd(&cellCount[newKey], 1);
    }
    if (newKey != oldKey) p.key[ip] = newKey;
  }
  flush_move_counters(stats, nMoved, nXCell, nXBlock, nLeft, nNotUsed, nWrap, nErr);
}

void launch_move_gyrokinetic(const DevMesh &m, const DevSpecies &sp, int order, int interp, double rSphere, ParticleSoA p, const int *nSlots,
                             long long nUpper, const double *bgTile, const double *gradBTile, const double *uE, const double *uB,
                             const double *uGradB, int *cellCount, DevMoveStats *stats, cudaStream_t s) {
  TpParams tp;
  tp.interp = interp, tp.backward = 0, tp.boundaryMode = sp.boundaryMode, tp.c = 0.0, tp.rSphere = rSphere, tp.exitCap = 0;
  GcTables T;
  T.bg = bgTile, T.gradB = gradBTile, T.uE = uE, T.uB = uB, T.uGradB = uGradB;
  long long g = (nUpper + 127) / 128;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  if (order == 2) move_gyrokinetic_kernel<true><<<(int)g, 128, 0, s>>>(m, sp, tp, T, p, nSlots, cellCount, stats);
  else move_gyrokinetic_kernel<false><<<(int)g, 128, 0, s>>>(m, sp, tp, T, p, nSlots, cellCount, stats);
}

// ------------------------------------------------------------------------------------------------
// f4: the species moments on the corners (the _PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_ part of ProcessCell /
// UpdateJMassMatrix, src/pic/pic_field_solver_ecsim.cpp:2270-2300, :2384-2392, :3874-3879): per species s and corner c
//   spec[c][s][0..9] += sum_p m~ W_c {1, vx, vy, vz, vx vx, vy vy, vz vz, vx vy, vy vz, vx vz} / CellVolume.
// One warp per cell of the sorted store; for one species at a time every lane keeps the 8 x 10 sums of its particles in
// registers, the warp folds them through shared memory and issues 80 REDs per (cell, species).
// ------------------------------------------------------------------------------------------------
constexpr int SM_WARPS = 2;
__global__ void __launch_bounds__(32 * SM_WARPS) species_moments_kernel(DevMesh m, DevSpecies sp, ParticleSoA p, const int *__restrict__ cellStart,
                                                                        double *__restrict__ spec) {
  __shared__ double sAcc[SM_WARPS][32][81];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warpGlobal = blockIdx.x * SM_WARPS + wib, nWarps = gridDim.x * SM_WARPS;
  const int C = m.cellsPerBlock, nS = sp.n;
  const int nIdx = m.nDepReal * C;
  for (int idx = warpGlobal; idx < nIdx; idx += nWarps) {
    const int rl = idx / C;
    const int leaf = m.depLeaf[rl];
    const int cell = leaf * C + (idx - rl * C);
    const int begin = cellStart[cell], end = cellStart[cell + 1];
    if (begin == end) continue;
    const LeafGeo &lg = m.leaf[leaf];
    const int cin = cell - leaf * C;
    const int kc = cin / (m.N[0] * m.N[1]);
    const int jc = (cin - kc * m.N[0] * m.N[1]) / m.N[0];
    const int ic = cin - kc * m.N[0] * m.N[1] - jc * m.N[0];
    int uidLane = 0;
    if (lane < 8) {
      const int cx = ((lane + 1) >> 1) & 1, cy = (lane >> 1) & 1, cz = (lane >> 2) & 1;  // cell-corner order (a3)
      uidLane = m.cornerUid[(size_t)leaf * m.nCornerLocal + cornerLocalNumber(m, ic + cx, jc + cy, kc + cz)];
    }
    // which species are present in the cell
    unsigned present = 0;
    for (int ip = begin + lane; ip < end; ip += 32) present |= 1u << (p.spec[ip] & 0x3f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) present |= __shfl_xor_sync(0xffffffffu, present, o);
    for (int s = 0; s < nS; s++) {
      if (!(present & (1u << s))) continue;
      double acc[80];
#pragma unroll
      for (int q = 0; q < 80; q++) acc[q] = 0.0;
      for (int ip = begin + lane; ip < end; ip += 32) {
        if ((p.spec[ip] & 0x3f) != s) continue;
        const double mass = sp.mass[s] * (sp.weight[s] * p.w[ip]);
        const double v0 = p.v[0][ip] * sp.length_conv, v1 = p.v[1][ip] * sp.length_conv, v2 = p.v[2][ip] * sp.length_conv;
        double xl[3];
        {
          const double xx[3] = {p.x[0][ip], p.x[1][ip], p.x[2][ip]};
#pragma unroll
          for (int d = 0; d < 3; d++) {  // CornerBased::InitStencil (pic_interpolation_routines.cpp:1090-1098)
            double xs = xx[d];
            const double xmx = lg.xmax[d], dxc = lg.dxc[d];
            if (fabs(xs - xmx) < 1e-10 * dxc) xs = xmx - 1e-10 * dxc;
            double r = (xs - lg.xmin[d]) / dxc;
            r -= (int)r;
            xl[d] = r;
          }
        }
        const double X[2] = {1.0 - xl[0], xl[0]}, Y[2] = {1.0 - xl[1], xl[1]}, Z[2] = {1.0 - xl[2], xl[2]};
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const int cx = ((c + 1) >> 1) & 1, cy = (c >> 1) & 1, cz = (c >> 2) & 1;
          const double t = mass * (X[cx] * Y[cy] * Z[cz]);
          const double t0 = t * v0, t1 = t * v1, t2 = t * v2;
          acc[10 * c + 0] += t;
          acc[10 * c + 1] += t0;
          acc[10 * c + 2] += t1;
          acc[10 * c + 3] += t2;
          acc[10 * c + 4] += t0 * v0;
          acc[10 * c + 5] += t1 * v1;
          acc[10 * c + 6] += t2 * v2;
          acc[10 * c + 7] += t0 * v1;
          acc[10 * c + 8] += t1 * v2;
          acc[10 * c + 9] += t0 * v2;
        }
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 80; q++) sAcc[wib][lane][q] = acc[q];
      __syncwarp();
      for (int q = lane; q < 80; q += 32) {
        double sum = 0.0;
        for (int l = 0; l < 32; l++) sum += sAcc[wib][l][q];
        const int c = q / 10, k = q - 10 * c;
        const int ui = __shfl_sync(__activemask(), uidLane, c);
        atomicAdd(spec + ((size_t)ui * nS + s) * 10 + k, sum * lg.invV);
      }
      __syncwarp();
    }
  }
}

void launch_species_moments(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *cellStart, double *spec, int nSM, cudaStream_t s) {
  cudaMemsetAsync(spec, 0, sizeof(double) * (size_t)m.nCorners * 10 * sp.n, s);
  species_moments_kernel<<<nSM * 8, 32 * SM_WARPS, 0, s>>>(m, sp, p, cellStart, spec);
}

// ------------------------------------------------------------------------------------------------
// f3: PIC::Sampling::ProcessCell  src/pic/pic.cpp:705-990 on the sorted store: per cell and species 13 sums (weight, number,
// number density, w v, w v^2, w |v|, w v_i v_(i+1)) ADDED to the collecting buffer sample[cell][species][13].  One warp per cell.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_cells_kernel(DevMesh m, DevSpecies sp, ParticleSoA p, const int *__restrict__ cellStart,
                                                          double *__restrict__ sample, unsigned long long *__restrict__ nSampled) {
  const int lane = threadIdx.x & 31;
  const int warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
  const int C = m.cellsPerBlock, nS = sp.n, nCells = m.nLeaves * C;
  for (int cell = warpGlobal; cell < nCells; cell += nWarps) {
    const int begin = cellStart[cell], end = cellStart[cell + 1];
    if (begin == end) continue;
    const LeafGeo &lg = m.leaf[cell / C];
    const double Measure = ((lg.xmax[0] - lg.xmin[0]) / m.N[0]) * ((lg.xmax[1] - lg.xmin[1]) / m.N[1]) * ((lg.xmax[2] - lg.xmin[2]) / m.N[2]);
    unsigned present = 0;
    for (int ip = begin + lane; ip < end; ip += 32) present |= 1u << (p.spec[ip] & 0x3f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) present |= __shfl_xor_sync(0xffffffffu, present, o);
    for (int s = 0; s < nS; s++) {
      if (!(present & (1u << s))) continue;
      double a[13];
#pragma unroll
      for (int q = 0; q < 13; q++) a[q] = 0.0;
      for (int ip = begin + lane; ip < end; ip += 32) {
        if ((p.spec[ip] & 0x3f) != s) continue;
        const double w = sp.weight[s] * p.w[ip];
        const double v0 = p.v[0][ip], v1 = p.v[1][ip], v2 = p.v[2][ip];
        a[0] += w, a[1] += 1.0, a[2] += w / Measure;
        a[3] += v0 * w, a[4] += v1 * w, a[5] += v2 * w;
        a[6] += (v0 * v0) * w, a[7] += (v1 * v1) * w, a[8] += (v2 * v2) * w;
        a[9] += sqrt(v0 * v0 + v1 * v1 + v2 * v2) * w;
        a[10] += (v0 * v1) * w, a[11] += (v1 * v2) * w, a[12] += (v2 * v0) * w;
      }
#pragma unroll
      for (int q = 0; q < 13; q++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
      if (lane == 0) {
        double *d = sample + ((size_t)cell * nS + s) * 13;  // the cell belongs to this warp: no atomics
#pragma unroll
        for (int q = 0; q < 13; q++) d[q] += a[q];
        atomicAdd(&nSampled[s], (unsigned long long)(a[1] + 0.5));
      }
    }
  }
}
void launch_sample_cells(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *cellStart, double *sample, unsigned long long *nSampled,
                         int nSM, cudaStream_t s) {
  sample_cells_kernel<<<nSM * 8, 256, 0, s>>>(m, sp, p, cellStart, sample, nSampled);
}

// ------------------------------------------------------------------------------------------------
// f4: ECSIM::CorrectParticleLocation  src/pic/pic_field_solver_ecsim.cpp:4440-4688
// Species 0 is displaced along -grad(phi)/(4 pi rho_e) (phi on the cell centres, rho_e = species-0 density on the closest
// corner, at most 0.1 cell) and re-filed; other species and cells at a block side without an (in use) neighbour keep their
// position.  A particle still filed in a periodic ghost block is lost from every list in the reference (exchangeParticleLocal
// :4366-4438 overwrites those lists) and is dropped here.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double cpl_interp2D(double vmm, double vpm, double vpp, double vmp, double dx, double dy) {  // :216-232
  return vmm * (1 - dx) * (1 - dy) + vpm * dx * (1 - dy) + vpp * dx * dy + vmp * (1 - dx) * dy;
}
__global__ void __launch_bounds__(128) correct_particle_location_kernel(DevMesh m, DevSpecies sp, ParticleSoA p, const int *__restrict__ nSlots,
                                                                        const double *__restrict__ phi, const double *__restrict__ spec,
                                                                        const unsigned *__restrict__ neibMask, double qom0,
                                                                        int *__restrict__ cellCount, unsigned long long *__restrict__ counters) {
  const int n = *nSlots;
  const int C = m.cellsPerBlock, nS = sp.n;
  const double Pi = 3.14159265358979323846264338327950288419716939937510582;
  unsigned nDisp = 0, nDel = 0, nErr = 0;
  for (int ip = blockIdx.x * blockDim.x + threadIdx.x; ip < n; ip += gridDim.x * blockDim.x) {
    const int key = p.key[ip];
    if (key < 0) continue;
    const int leaf = key / C;
    const LeafGeo &lg = m.leaf[leaf];
    if (m.periodic && lg.face != 0) {  // not walked, then overwritten by exchangeParticleLocal
      p.key[ip] = -1;
      nDel++;
      continue;
    }
    if ((p.spec[ip] & 0x3f) != 0) {  // xFinal = xInit: same block, same cell
      atomicAdd(&cellCount[key], 1);
      continue;
    }
    const int cin = key - leaf * C;
    int index[3];
    index[2] = cin / (m.N[0] * m.N[1]);
    index[1] = (cin - index[2] * m.N[0] * m.N[1]) / m.N[0];
    index[0] = cin - index[2] * m.N[0] * m.N[1] - index[1] * m.N[0];
    double dx[3], xNode[3], xCell[3];
    for (int d = 0; d < 3; d++) {
      dx[d] = (lg.xmax[d] - lg.xmin[d]) / m.N[d] * sp.length_conv;
      xNode[d] = lg.xmin[d] + dx[d] * index[d];
      xCell[d] = lg.xmin[d] + dx[d] * (index[d] + 0.5);
    }
    {
      // isBoundaryCell != 0 (:6963-6999): a neighbour across the block sides this cell touches is missing or not in use
      const unsigned mask = neibMask[leaf];
      bool atBoundary = false;
      if (mask) {
        int lo[3], hi[3];
        for (int d = 0; d < 3; d++) {
          lo[d] = fabs(xCell[d] - 0.5 * dx[d] - lg.xmin[d]) < m.eps;
          hi[d] = fabs(xCell[d] + 0.5 * dx[d] - lg.xmax[d]) < m.eps;
        }
        for (int q = 0; q < 27 && !atBoundary; q++) {
          if (q == 13 || !(mask & (1u << q))) continue;
          const int s[3] = {q % 3 - 1, (q / 3) % 3 - 1, q / 9 - 1};
          bool touched = true;
          for (int d = 0; d < 3; d++)
            if ((s[d] < 0 && !lo[d]) || (s[d] > 0 && !hi[d])) touched = false;
          atBoundary = touched;
        }
      }
      if (atBoundary) {
        atomicAdd(&cellCount[key], 1);
        continue;
      }
    }
    const double xInit[3] = {p.x[0][ip], p.x[1][ip], p.x[2][ip]};
    double xRel[3];
    int iClosestNode[3];
    for (int d = 0; d < 3; d++) {
      xRel[d] = (xInit[d] - xNode[d]) / dx[d];
      iClosestNode[d] = (int)(index[d] + round(xRel[d]));
    }
    for (int d = 0; d < 3; d++) xRel[d] = xRel[d] >= 0.5 ? xRel[d] - 0.5 : xRel[d] + 0.5;
    // Phi[ix-1..ix][iy-1..iy][iz-1..iz] with ix = iClosestNode+1: the centres iClosestNode-1, iClosestNode of the block
    const int *cuid = m.centerUid + (size_t)leaf * m.nCenterLocal;
    double P[2][2][2];
    for (int a = 0; a < 2; a++)
      for (int b = 0; b < 2; b++)
        for (int c = 0; c < 2; c++) {
          const int u = cuid[centerLocalNumber(m, iClosestNode[0] - 1 + a, iClosestNode[1] - 1 + b, iClosestNode[2] - 1 + c)];
          P[a][b][c] = (u >= 0) ? phi[u] : 0.0;
        }
    double GradPhi[3];
    GradPhi[0] = cpl_interp2D(P[1][0][0] - P[0][0][0], P[1][1][0] - P[0][1][0], P[1][1][1] - P[0][1][1], P[1][0][1] - P[0][0][1], xRel[1], xRel[2]);
    GradPhi[1] = cpl_interp2D(P[0][1][0] - P[0][0][0], P[1][1][0] - P[1][0][0], P[1][1][1] - P[1][0][1], P[0][1][1] - P[0][0][1], xRel[0], xRel[2]);
    GradPhi[2] = cpl_interp2D(P[0][0][1] - P[0][0][0], P[1][0][1] - P[1][0][0], P[1][1][1] - P[1][1][0], P[0][1][1] - P[0][1][0], xRel[0], xRel[1]);
    for (int d = 0; d < 3; d++) GradPhi[d] /= dx[d];
    const int cu = m.cornerUid[(size_t)leaf * m.nCornerLocal + cornerLocalNumber(m, iClosestNode[0], iClosestNode[1], iClosestNode[2])];
    const double eChargeDens = spec[(size_t)cu * nS * 10] * qom0;  // SpeciesDataIndex[0] + Rho_
    const double eps = 0.9;
    double displacement[3], temp;
    if (eChargeDens != 0) temp = 1. / (4. * Pi * eChargeDens);
    else temp = 0;
    for (int d = 0; d < 3; d++) displacement[d] = -eps * GradPhi[d] * temp;
    const double epsLimit = 0.1;
    if (fabs(displacement[0] / dx[0]) > epsLimit || fabs(displacement[1] / dx[1]) > epsLimit || fabs(displacement[2] / dx[2]) > epsLimit) {
      // (pow(d,2) of the reference: d*d is the correctly rounded square)
      const double dl = sqrt(displacement[0] * displacement[0] + displacement[1] * displacement[1] + displacement[2] * displacement[2]);
      for (int d = 0; d < 3; d++) displacement[d] *= epsLimit * dx[0] / dl;
    }
    double xFinal[3];
    for (int d = 0; d < 3; d++) xFinal[d] = xInit[d] + displacement[d];
    nDisp++;
    const int newNode = find_tree_node_plain(m, xFinal, lg.node);
    int newKey = -1;
    if (newNode < 0 || m.nodeLeaf[newNode] < 0) {
      nDel++;  // DeleteParticle: outside the domain, or a node without a block
    } else {
      int ijk[3];
      if (!find_cell_index(m, xFinal, newNode, ijk)) {
        nErr++;  // exit("cannot find the cell") in the reference
        p.key[ip] = -1;
        continue;
      }
      newKey = m.nodeLeaf[newNode] * C + ijk[0] + m.N[0] * (ijk[1] + m.N[1] * ijk[2]);
      p.x[0][ip] = xFinal[0], p.x[1][ip] = xFinal[1], p.x[2][ip] = xFinal[2];
      atomicAdd(&cellCount[newKey], 1);
    }
    if (newKey != key) p.key[ip] = newKey;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nDisp += __shfl_xor_sync(0xffffffffu, nDisp, o);
    nDel += __shfl_xor_sync(0xffffffffu, nDel, o);
    nErr += __shfl_xor_sync(0xffffffffu, nErr, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (nDisp) atomicAdd(&counters[0], (unsigned long long)nDisp);
    if (nDel) atomicAdd(&counters[1], (unsigned long long)nDel);
    if (nErr) atomicAdd(&counters[2], (unsigned long long)nErr);
  }
}

void launch_correct_particle_location(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *nSlots, long long nUpper, const double *phi,
                                      const double *spec, const unsigned *neibMask, double qom0, int *cellCount, unsigned long long *counters,
                                      cudaStream_t s) {
  cudaMemsetAsync(counters, 0, 3 * sizeof(unsigned long long), s);
  cudaMemsetAsync(cellCount, 0, sizeof(int) * (size_t)m.nLeaves * m.cellsPerBlock, s);  // the histogram the sort / migration use
  long long g = (nUpper + 127) / 128;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  correct_particle_location_kernel<<<(int)g, 128, 0, s>>>(m, sp, p, nSlots, phi, spec, neibMask, qom0, cellCount, counters);
}

}  // namespace amps
