// mover.cu -- particle push kernels (sm_100a).  COMPILED WITH --fmad=false:
// the position update and the cell search must round exactly like the reference's CPU build
// (no FMA contraction; IEEE fp64 division), else a particle within 1 ulp of a face lands in a
// different cell than on the CPU (hard part 1).
//
// Kernels
//   stage_tiles_kernel     a2  SetBlock_E/SetBlock_B   src/pic/pic_mover.cpp:86-166
//   move_lapenta_kernel    a5  PIC::Mover::Lapenta2017 src/pic/pic_mover_boris.cpp:876-1393
//                          a3  CornerBased::InitStencil             pic_interpolation_routines.cpp:1074-1194
//                          a4  CellCentered::Linear::InitStencil    pic_interpolation_routines.cpp:224-330,820-907
//                          a14 findTreeNode / FindCellIndex         meshAMRgeneric.h:2793-2882, 2256-2323
//                          a16 periodic wrap                        pic_bc_periodic.cpp:86-178
//
// Mapping: one CTA per (leaf block, slice of its particle range); the block's E (corner) and
// B (centre) tiles incl. ghost layer are brought into shared memory by two 1-D TMA bulk copies
// (cp.async.bulk -> UBLKCP) completing on an mbarrier; one thread per particle, SoA loads are
// unit-stride.  The new (block,cell) key is histogrammed here so the counting sort needs no
// extra pass over the particles.
#include "amps_dev.cuh"
#include "mover_common.cuh"
#include "tma.cuh"

namespace amps {

// self test: out[0] += number of (a,b) pairs for which the helpers differ from IEEE '/'
__global__ void division_selftest_kernel(const double *__restrict__ a, const double *__restrict__ b, int n, unsigned long long *out) {
  unsigned long long bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double bb = b[i];
    const Recip rc = make_recip(bb);
    double w[8], ref[8];
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const double aa = a[(i + s * 7919) % n];
      w[s] = aa;
      ref[s] = aa / bb;
      if (div_rn(aa, bb, rc) != ref[s] && !(ref[s] != ref[s])) bad++;
    }
    normalize8(w, bb);
    if (bb > 0.0 && bb != 1.0) {
#pragma unroll
      for (int s = 0; s < 8; s++)
        if (w[s] != ref[s] && !(ref[s] != ref[s])) bad++;
    }
  }
  if (bad) atomicAdd(out, bad);
}
void launch_division_selftest(const double *a, const double *b, int n, unsigned long long *out, cudaStream_t s) {
  division_selftest_kernel<<<148 * 4, 256, 0, s>>>(a, b, n, out);
}

// ------------------------------------------------------------------------------------------------
// a2: gather unique-node fields into per-leaf tiles (incl. ghost layers); missing node -> 0
// ------------------------------------------------------------------------------------------------
__global__ void stage_tiles_kernel(DevMesh m, bool cornerB, const double *__restrict__ E_half, const double *__restrict__ B_prev, const double *__restrict__ B_cur,
                                   double *__restrict__ eTile, double *__restrict__ bPrevTile, double *__restrict__ bCurTile) {
  const int leaf = blockIdx.x;
  if (E_half) {
    const int *uid = m.cornerUid + (size_t)leaf * m.nCornerLocal;
    double *dst = eTile + (size_t)leaf * m.eTileStride;
    for (int i = threadIdx.x; i < m.nCornerLocal; i += blockDim.x) {
      int u = uid[i];
      double a = 0.0, b = 0.0, c = 0.0;
      if (u >= 0) a = E_half[3 * (size_t)u], b = E_half[3 * (size_t)u + 1], c = E_half[3 * (size_t)u + 2];
      dst[3 * i] = a, dst[3 * i + 1] = b, dst[3 * i + 2] = c;
    }
  }
  // B lives on the centre nodes, or on the corner nodes with _PIC_FIELD_SOLVER_B_CORNER_BASED_ (OffsetB_corner,
  // pic_field_solver_ecsim.cpp:534-538)
  const int *cuid = cornerB ? m.cornerUid + (size_t)leaf * m.nCornerLocal : m.centerUid + (size_t)leaf * m.nCenterLocal;
  const int nLocal = cornerB ? m.nCornerLocal : m.nCenterLocal;
  for (int pass = 0; pass < 2; pass++) {
    const double *src = pass ? B_cur : B_prev;
    double *dst = (pass ? bCurTile : bPrevTile);
    if (!src) continue;
    dst += (size_t)leaf * m.bTileStride;
    for (int i = threadIdx.x; i < nLocal; i += blockDim.x) {
      int u = cuid[i];
      double a = 0.0, b = 0.0, c = 0.0;
      if (u >= 0) a = src[3 * (size_t)u], b = src[3 * (size_t)u + 1], c = src[3 * (size_t)u + 2];
      dst[3 * i] = a, dst[3 * i + 1] = b, dst[3 * i + 2] = c;
    }
  }
}

void launch_stage_tiles(const DevMesh &m, bool cornerB, const double *E_half, const double *B_prev, const double *B_cur, double *eTile,
                        double *bPrevTile, double *bCurTile, cudaStream_t s) {
  stage_tiles_kernel<<<m.nLeaves, 256, 0, s>>>(m, cornerB, E_half, B_prev, B_cur, eTile, bPrevTile, bCurTile);
}

// ------------------------------------------------------------------------------------------------
// a5: Lapenta2017
// ------------------------------------------------------------------------------------------------
struct BlockConst {
  double dxc[3], span[3], dxCell[3];
  Recip rDxc[3], rSpan[3], rCell[3], rRef[3];
  double qdt2m[AMPS_GPU_MAX_SPECIES], dt[AMPS_GPU_MAX_SPECIES];
};

template <bool kSmemTiles, bool kCornerB>
__global__ void __launch_bounds__(256, 2) move_lapenta_kernel(DevMesh m, DevSpecies sp, ParticleSoA p, const int *__restrict__ cellStart,
                                                          const double *__restrict__ eTileG, const double *__restrict__ bTileG,
                                                          int *__restrict__ cellCount, DevMoveStats *__restrict__ stats, int slices,
                                                          amps_gpu_exit_record *__restrict__ exitBuf, unsigned long long *__restrict__ exitCount,
                                                          long long exitCap, const unsigned char *__restrict__ redoMask,
                                                          const int *__restrict__ redoLeafList, const int *__restrict__ nRedoLeaves) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t mbar;
  __shared__ LeafGeo sLeaf;
  __shared__ FaceGeo sFace;
  __shared__ BlockConst sC;

  unsigned int nMoved = 0, nXCell = 0, nXBlock = 0, nLeft = 0, nNotUsed = 0, nWrap = 0, nErr = 0;
  // work item = (leaf block, slice).  Normal launch: one item per CTA.  Second pass behind move_lapenta_fast_kernel
  // (redoLeafList != nullptr): a small grid strides over the blocks that hold flagged particles (usually none).
  const int nWork = redoLeafList ? *nRedoLeaves * slices : (int)gridDim.x;
  uint32_t mbarParity = 0;
  if (kSmemTiles && threadIdx.x == 0) {
    mbar_init(&mbar, 1);
    fence_mbar_init();
  }
  for (int work = blockIdx.x; work < nWork; work += gridDim.x) {
  const int slot = work / slices, slice = work - slot * slices;
  const int leaf = redoLeafList ? redoLeafList[slot] : slot;
  const int C = m.cellsPerBlock;
  const int begin = cellStart[(size_t)leaf * C], end = cellStart[(size_t)(leaf + 1) * C];
  const long long len = (long long)end - begin;
  const int b = begin + (int)(len * slice / slices), e = begin + (int)(len * (slice + 1) / slices);
  if (b >= e) continue;
  __syncthreads();  // the previous work item is done with sLeaf / sC / the tiles

  const double *sE, *sB;
  if (kSmemTiles) {
    double *tE = reinterpret_cast<double *>(smem_raw);
    double *tB = tE + m.eTileStride;
    if (threadIdx.x == 0) sLeaf = m.leaf[leaf];
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t bytesE = (uint32_t)m.eTileStride * 8u, bytesB = (uint32_t)m.bTileStride * 8u;
      mbar_expect_tx(&mbar, bytesE + bytesB);
      bulk_g2s(tE, eTileG + (size_t)leaf * m.eTileStride, bytesE, &mbar);
      bulk_g2s(tB, bTileG + (size_t)leaf * m.bTileStride, bytesB, &mbar);
    }
    mbar_wait(&mbar, mbarParity);
    mbarParity ^= 1u;
    sE = tE, sB = tB;
  } else {
    if (threadIdx.x == 0) sLeaf = m.leaf[leaf];
    __syncthreads();
    sE = eTileG + (size_t)leaf * m.eTileStride;
    sB = bTileG + (size_t)leaf * m.bTileStride;
  }

  const LeafGeo &lg = sLeaf;
  // per-block constants of the stencils and their shared reciprocals (computed once per work item)
  if (threadIdx.x < 3) {
    const int d = threadIdx.x;
    sC.dxc[d] = (lg.xmax[d] - lg.xmin[d]) / m.N[d];  // CornerBased::InitStencil :1086-1088
    sC.span[d] = lg.xmax[d] - lg.xmin[d];
    sC.dxCell[d] = m.dxRoot[d] / (1 << lg.level) / double(m.N[d]);  // FindCellIndex :2275
    sC.rDxc[d] = make_recip(sC.dxc[d]);
    sC.rSpan[d] = make_recip(sC.span[d]);
    sC.rCell[d] = make_recip(sC.dxCell[d]);
    sC.rRef[d] = make_recip(m.dxMaxRef[d]);
  } else if (threadIdx.x == 64) {
    if (sp.boundaryMode != AMPS_BOUNDARY_DELETE) init_faces(m, sFace);
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + AMPS_GPU_MAX_SPECIES) {
    const int sidx = threadIdx.x - 32;
    const double dts = (sp.timeStepMode == AMPS_DT_SPECIES_GLOBAL) ? sp.dt[sidx] : sp.dt[0];
    const double QdT_over_m = (sidx < sp.n) ? sp.charge[sidx] * dts / sp.mass[sidx] : 0.0;  // :1036
    sC.qdt2m[sidx] = 0.5 * QdT_over_m;
    sC.dt[sidx] = dts;
  }
  __syncthreads();
  double dxc[3], span[3];
  Recip rDxc[3], rSpan[3], rRef[3], rCell[3];
  bool blockOk = true;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    dxc[d] = sC.dxc[d], span[d] = sC.span[d];
    rDxc[d] = sC.rDxc[d], rSpan[d] = sC.rSpan[d], rRef[d] = sC.rRef[d], rCell[d] = sC.rCell[d];
    blockOk = blockOk && rDxc[d].ok && rSpan[d].ok && rRef[d].ok && rCell[d].ok;
  }
  const int CS0 = 1 + m.TN[0], CS1 = (1 + m.TN[0]) * (1 + m.TN[1]);  // corner strides
  const int BS0 = m.TN[0], BS1 = m.TN[0] * m.TN[1];                  // centre strides

  for (int ip = b + threadIdx.x; ip < e; ip += blockDim.x) {
    if (redoMask && !redoMask[ip]) continue;
    double xInit[3], vInit[3], xFinal[3], vFinal[3];
    xInit[0] = p.x[0][ip], xInit[1] = p.x[1][ip], xInit[2] = p.x[2][ip];
    vInit[0] = p.v[0][ip], vInit[1] = p.v[1][ip], vInit[2] = p.v[2][ip];
    const int spec = p.spec[ip] & 0x3f;
    const int oldKey = p.key[ip];
    const double dtTotal = sC.dt[spec];
    nMoved++;
    bool err = false;

    // ---- a3: corner stencil for E (mutates xInit: snap to xmax-1e-10dx) ----
    int iX[3];
    double xLoc[3];
    double off[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      if ((xInit[d] < lg.xmin[d]) || (xInit[d] > lg.xmax[d])) err = true;  // reference: exit("the point is out of block")
      if (fabs(xInit[d] - lg.xmax[d]) < 1e-10 * dxc[d]) xInit[d] = lg.xmax[d] - 1e-10 * dxc[d];
      off[d] = xInit[d] - lg.xmin[d];
    }
    div_rn3(off, dxc, rDxc, blockOk, xLoc);
#pragma unroll
    for (int d = 0; d < 3; d++) {
      iX[d] = (int)(xLoc[d]);
      xLoc[d] -= iX[d];
    }
    if (err) {
      nErr++;
      p.key[ip] = -1;
      continue;
    }
    double E[3] = {0.0, 0.0, 0.0}, B[3] = {0.0, 0.0, 0.0};
    {
      // stencil order: iStencil, jStencil, kStencil with k fastest (:1112)
      const double ax0 = 1.0 - xLoc[0], ax1 = xLoc[0], ay0 = 1.0 - xLoc[1], ay1 = xLoc[1], az0 = 1.0 - xLoc[2], az1 = xLoc[2];
      double w[8];
      w[0] = ax0 * ay0 * az0;  // (0,0,0)
      w[1] = ax0 * ay0 * az1;  // (0,0,1)
      w[2] = ax0 * ay1 * az0;  // (0,1,0)
      w[3] = ax0 * ay1 * az1;  // (0,1,1)
      w[4] = ax1 * ay0 * az0;  // (1,0,0)
      w[5] = ax1 * ay0 * az1;  // (1,0,1)
      w[6] = ax1 * ay1 * az0;  // (1,1,0)
      w[7] = ax1 * ay1 * az1;  // (1,1,1)
      double norm = 0.0;
#pragma unroll
      for (int s = 0; s < 8; s++) norm += w[s];
      normalize8(w, norm);  // Stencil.Normalize(): w[s] /= norm
      const int nd0 = cornerLocalNumber(m, iX[0], iX[1], iX[2]);
#pragma unroll
      for (int s = 0; s < 8; s++) {
        const int nd = nd0 + ((s >> 2) & 1) + ((s >> 1) & 1) * CS0 + (s & 1) * CS1;
        const double *t = sE + 3 * nd;
        E[0] += w[s] * t[0];
        E[1] += w[s] * t[1];
        E[2] += w[s] * t[2];
        if (kCornerB) {  // _PIC_FIELD_SOLVER_B_CORNER_BASED_: same stencil on the corner B (:952-965)
          const double *tb = sB + 3 * nd;
          B[0] += w[s] * tb[0];
          B[1] += w[s] * tb[1];
          B[2] += w[s] * tb[2];
        }
      }
    }
    // ---- a4: cell-centred linear stencil for B (uniform / same-level branch) ----
    if (!kCornerB) {
      double qLoc[3];
      div_rn3(off, span, rSpan, blockOk, qLoc);  // (x-xmin)/(xmax-xmin), :278-280
      const double iLoc = qLoc[0] * m.N[0];
      const double jLoc = qLoc[1] * m.N[1];
      const double kLoc = qLoc[2] * m.N[2];
      const int i0 = (iLoc < 0.5) ? -1 : (int)(iLoc - 0.50);
      const int j0 = (jLoc < 0.5) ? -1 : (int)(jLoc - 0.50);
      const int k0 = (kLoc < 0.5) ? -1 : (int)(kLoc - 0.50);
      const double w0 = iLoc - (i0 + 0.5), w1 = jLoc - (j0 + 0.5), w2 = kLoc - (k0 + 0.5);
      double w[8];
      // loop order i,j,k with k fastest; weight by code i+2j+4k (:850-881)
      w[0] = (1.0 - w0) * (1.0 - w1) * (1.0 - w2);  // i0 j0 k0
      w[1] = (1.0 - w0) * (1.0 - w1) * w2;          // i0 j0 k1
      w[2] = (1.0 - w0) * w1 * (1.0 - w2);          // i0 j1 k0
      w[3] = (1.0 - w0) * w1 * w2;                  // i0 j1 k1
      w[4] = w0 * (1.0 - w1) * (1.0 - w2);          // i1 j0 k0
      w[5] = w0 * (1.0 - w1) * w2;                  // i1 j0 k1
      w[6] = w0 * w1 * (1.0 - w2);                  // i1 j1 k0
      w[7] = w0 * w1 * w2;                          // i1 j1 k1
      // AddCell drops centres outside the global box in non-periodic mode (pic.h:7235-7245)
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
      // Normalize(): the reference tests the GLOBAL StencilTable->Length (:903): it runs unless ComputeNetCharge has left an 8-cell
      // stencil there (sp.globalStencilFull); a stencil that lost cells at the domain boundary is always normalised
      if (!(sp.globalStencilFull && valid == 0xffu)) normalize8(w, norm);
      const int nd0 = centerLocalNumber(m, i0, j0, k0);
#pragma unroll
      for (int s = 0; s < 8; s++) {
        if (valid & (1u << s)) {
          const int nd = nd0 + ((s >> 2) & 1) + ((s >> 1) & 1) * BS0 + (s & 1) * BS1;
          const double *t = sB + 3 * nd;
          B[0] += w[s] * t[0];
          B[1] += w[s] * t[1];
          B[2] += w[s] * t[2];
        }
      }
    }

    // ---- velocity / position update (:1036-1081) ----
    {
      const double QdT_over_2m = sC.qdt2m[spec];  // 0.5*(chargeQ*dtTotal/mass)
      const double QdT_over_2m_squared = QdT_over_2m * QdT_over_2m;
      double BB[3][3], P[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        P[i] = -QdT_over_2m * B[i];
#pragma unroll
        for (int j = 0; j <= i; j++) {
          BB[i][j] = QdT_over_2m_squared * B[i] * B[j];
          BB[j][i] = BB[i][j];
        }
      }
      const double c0 = 1.0 / (1.0 + QdT_over_2m_squared * (B[0] * B[0] + B[1] * B[1] + B[2] * B[2]));
      double alpha[3][3];
      alpha[0][0] = c0 * (1.0 + BB[0][0]);
      alpha[0][1] = c0 * (-P[2] + BB[0][1]);
      alpha[0][2] = c0 * (P[1] + BB[0][2]);
      alpha[1][0] = c0 * (P[2] + BB[1][0]);
      alpha[1][1] = c0 * (1.0 + BB[1][1]);
      alpha[1][2] = c0 * (-P[0] + BB[1][2]);
      alpha[2][0] = c0 * (-P[1] + BB[2][0]);
      alpha[2][1] = c0 * (P[0] + BB[2][1]);
      alpha[2][2] = c0 * (1.0 + BB[2][2]);
#pragma unroll
      for (int d = 0; d < 3; d++) {
        double vp = 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) vp += alpha[d][j] * (vInit[j] + QdT_over_2m * E[j]);
        vFinal[d] = 2.0 * vp - vInit[d];
      }
#pragma unroll
      for (int d = 0; d < 3; d++) xFinal[d] = xInit[d] + dtTotal * vFinal[d];
    }

    // ---- a14/a15: new block (:1121-1272) ----
    int node = find_tree_node(m, xFinal, lg, rRef, blockOk);
    int newKey = -1;
    if (node < 0) {
      // left the domain (:1159-1265): DELETE, or the face search + exit record for the host callback
      if (sp.boundaryMode == AMPS_BOUNDARY_DELETE) {
        nLeft++;
      } else {
        // copies keep the hot path's x/v in registers (the out-of-line exit search takes its arguments by address)
        int face, exitNode;
        double ex[3] = {xInit[0], xInit[1], xInit[2]}, ev[3] = {vInit[0], vInit[1], vInit[2]};
        double fx[3] = {xFinal[0], xFinal[1], xFinal[2]}, fv[3] = {vFinal[0], vFinal[1], vFinal[2]};
        const int code = domain_exit_vmiddle(m, sFace, sp.boundaryMode, dtTotal, ex, ev, fx, fv, lg.node, &face, &exitNode);
        if (code == 0) {
          add_exit_record(exitBuf, exitCount, exitCap, p.ptr[ip], spec, face, m.nodeLeaf[exitNode], ex, ev);
          nLeft++;
        } else {
          nErr++;  // specular: _PARTICLE_REJECTED_ON_THE_FACE_ -> exit("not implemented") :1262-1263
        }
      }
    } else if (!(m.nodeFlags[node] & AMPS_NODE_USED)) {
      nNotUsed++;
    } else {
      int newLeaf = m.nodeLeaf[node];
      if (newLeaf < 0) {
        nLeft++;  // block not allocated on this rank (:1290-1302)
      } else {
        // FindCellIndex (:2256-2323)
        const bool same = (newLeaf == leaf);
        const int lev = same ? lg.level : m.nodeLevel[node];
        int ijk[3];
        bool out = false;
#pragma unroll
        for (int d = 0; d < 3; d++) {
          const double lo = same ? lg.xmin[d] : m.nxmin[3 * node + d];
          const double hi = same ? lg.xmax[d] : m.nxmax[3 * node + d];
          if ((xFinal[d] < lo) || (hi < xFinal[d])) out = true;
          int c;
          if (lev == lg.level) {
            c = (int)div_rn(xFinal[d] - lo, sC.dxCell[d], rCell[d]);
          } else {
            const double dx = m.dxRoot[d] / (1 << lev) / double(m.N[d]);
            c = (int)div_slow(xFinal[d] - lo, dx);
          }
          if (c == m.N[d]) c = m.N[d] - 1;
          ijk[d] = c;
        }
        if (out) {
          nErr++;  // reference: exit("cannot find the cell")
        } else {
          // a16: periodic ghost block -> real block (pic_bc_periodic.cpp:100-134); the cell index is inherited
          const int realLeaf = same ? -1 : m.leaf[newLeaf].real;
          if (realLeaf >= 0) {
            const LeafGeo &gg = m.leaf[newLeaf];
            const LeafGeo &rg = m.leaf[realLeaf];
#pragma unroll
            for (int d = 0; d < 3; d++) {
              const double sh = rg.xmin[d] - gg.xmin[d];
              xFinal[d] += sh;
              if (xFinal[d] < rg.xmin[d]) xFinal[d] = rg.xmin[d];
              if (xFinal[d] >= rg.xmax[d]) xFinal[d] = rg.xmax[d] - 1.0E-10 * (rg.xmax[d] - rg.xmin[d]);
            }
            newLeaf = realLeaf;
            nWrap++;
          }
          newKey = newLeaf * C + ijk[0] + m.N[0] * (ijk[1] + m.N[1] * ijk[2]);
          if (newLeaf != leaf) nXBlock++;
          else if (newKey != oldKey) nXCell++;
        }
      }
    }

    if (newKey >= 0) {
      p.x[0][ip] = xFinal[0], p.x[1][ip] = xFinal[1], p.x[2][ip] = xFinal[2];
      p.v[0][ip] = vFinal[0], p.v[1][ip] = vFinal[1], p.v[2][ip] = vFinal[2];
      atomicAdd(&cellCount[newKey], 1);
    }
    if (newKey != oldKey) p.key[ip] = newKey;
  }
  }  // work items

  flush_move_counters(stats, nMoved, nXCell, nXBlock, nLeft, nNotUsed, nWrap, nErr);
}

void launch_move_lapenta(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *cellStart, const double *eTile, const double *bTile,
                         int *cellCount, DevMoveStats *stats, int slices, amps_gpu_exit_record *exitBuf, unsigned long long *exitCount,
                         long long exitCap, const unsigned char *redoMask, const int *redoLeafList, const int *nRedoLeaves, cudaStream_t s) {
  const size_t smem = (size_t)(m.eTileStride + m.bTileStride) * sizeof(double);
  int grid = m.nLeaves * slices;
  if (redoLeafList && grid > 148 * 2) grid = 148 * 2;  // second pass: persistent CTAs over the (short) list of flagged blocks
  const bool cornerB = sp.bMode == AMPS_B_CORNER_BASED;
  if (smem <= 200 * 1024) {
    static OncePerDevice once;
    if (once.first()) {
      cudaFuncSetAttribute(move_lapenta_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(move_lapenta_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    if (cornerB)
      move_lapenta_kernel<true, true><<<grid, 256, smem, s>>>(m, sp, p, cellStart, eTile, bTile, cellCount, stats, slices, exitBuf, exitCount, exitCap, redoMask, redoLeafList, nRedoLeaves);
    else
      move_lapenta_kernel<true, false><<<grid, 256, smem, s>>>(m, sp, p, cellStart, eTile, bTile, cellCount, stats, slices, exitBuf, exitCount, exitCap, redoMask, redoLeafList, nRedoLeaves);
  } else if (cornerB) {
    move_lapenta_kernel<false, true><<<grid, 256, 0, s>>>(m, sp, p, cellStart, eTile, bTile, cellCount, stats, slices, exitBuf, exitCount, exitCap, redoMask, redoLeafList, nRedoLeaves);
  } else {
    move_lapenta_kernel<false, false><<<grid, 256, 0, s>>>(m, sp, p, cellStart, eTile, bTile, cellCount, stats, slices, exitBuf, exitCount, exitCap, redoMask, redoLeafList, nRedoLeaves);
  }
}

}  // namespace amps
