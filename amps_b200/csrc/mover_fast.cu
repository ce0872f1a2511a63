// mover_fast.cu -- the production path of PIC::Mover::Lapenta2017 (a5): FMA contraction ON, reciprocal
// multiplies instead of IEEE divisions, no Stencil.Normalize() (the trilinear weights sum to 1 within 4 ulp).
//
// Why this is still bit-exact where the contract demands it.  Every quantity computed here differs from the
// reference's rounding by a few ulp (|dx'| <~ 1e-14 |x'|).  The (block,cell) key is the integer part of
// (x'-xmin)/dx; a few-ulp change of x' can only change it when x' lies within ~1e-12 of a cell face.  A particle
// whose x' ends within GUARD = 1e-8 cell widths of a cell face (block faces and the domain boundary are cell
// faces), or that does anything but "land in a used, locally allocated block" (leaves the domain, hits a block
// that is not in use, needs the truncated boundary stencil, lies outside its block on entry), is NOT finished
// here: it is flagged and the exact kernel (mover.cu: no contraction, correctly rounded quotients, bit-identical
// to the CPU) pushes it from its untouched state.  The flagged fraction is ~6e-8 per particle in a periodic box.
// Result: keys, crossing counters and deletions bit-exact; x', v' within ~1e-14 relative of the reference
// (tolerance of the contract: 1e-10).  cfg.exact_arithmetic = 1 routes every particle through the exact kernel.
#include "amps_dev.cuh"
#include "mover_common.cuh"
#include "tma.cuh"

namespace amps {

constexpr double MOVE_GUARD = 1.0e-8;
#ifndef FAST_CTAS
#define FAST_CTAS 2
#endif

struct FastBlockConst {
  double dxc[3], invDxc[3], invSpanN[3], invCell[3], invRef[3];
  double qdt2m[AMPS_GPU_MAX_SPECIES], dt[AMPS_GPU_MAX_SPECIES];
};

// kCube10: 8^3-cell blocks with one ghost layer (tiles of 11^3 corners / 10^3 centres): the strides of the two stencils are
// compile-time constants, so the 8 + 8 node addresses are immediate offsets of one base
template <bool kSmemTiles, bool kCornerB, bool kCube10>
__global__ void __launch_bounds__(256, FAST_CTAS) move_lapenta_fast_kernel(DevMesh m, DevSpecies sp, ParticleSoA p, const int *__restrict__ cellStart,
                                                                   const double *__restrict__ eTileG, const double *__restrict__ bTileG,
                                                                   int *__restrict__ cellCount, DevMoveStats *__restrict__ stats, int slices,
                                                                   unsigned char *__restrict__ redoMask, int *__restrict__ leafRedo,
                                                                   int *__restrict__ redoLeafList, int *__restrict__ nRedoLeaves) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t mbar;
  __shared__ LeafGeo sLeaf;
  __shared__ FastBlockConst sC;

  const int leaf = blockIdx.x / slices, slice = blockIdx.x - leaf * slices;
  const int C = m.cellsPerBlock;
  const int begin = cellStart[(size_t)leaf * C], end = cellStart[(size_t)(leaf + 1) * C];
  const long long len = (long long)end - begin;
  const int b = begin + (int)(len * slice / slices), e = begin + (int)(len * (slice + 1) / slices);
  if (b >= e) return;

  const double *sE, *sB;
  if (kSmemTiles) {
    double *tE = reinterpret_cast<double *>(smem_raw);
    double *tB = tE + m.eTileStride;
    if (threadIdx.x == 0) {
      sLeaf = m.leaf[leaf];
      mbar_init(&mbar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t bytesE = (uint32_t)m.eTileStride * 8u, bytesB = (uint32_t)m.bTileStride * 8u;
      mbar_expect_tx(&mbar, bytesE + bytesB);
      bulk_g2s(tE, eTileG + (size_t)leaf * m.eTileStride, bytesE, &mbar);
      bulk_g2s(tB, bTileG + (size_t)leaf * m.bTileStride, bytesB, &mbar);
    }
    sE = tE, sB = tB;
  } else {
    if (threadIdx.x == 0) sLeaf = m.leaf[leaf];
    __syncthreads();
    sE = eTileG + (size_t)leaf * m.eTileStride;
    sB = bTileG + (size_t)leaf * m.bTileStride;
  }
  const LeafGeo &lg = sLeaf;
  if (threadIdx.x < 3) {
    const int d = threadIdx.x;
    const double span = lg.xmax[d] - lg.xmin[d];
    sC.dxc[d] = span / m.N[d];
    sC.invDxc[d] = 1.0 / sC.dxc[d];
    sC.invSpanN[d] = (double)m.N[d] / span;
    sC.invCell[d] = 1.0 / (m.dxRoot[d] / (1 << lg.level) / double(m.N[d]));
    sC.invRef[d] = 1.0 / m.dxMaxRef[d];
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + AMPS_GPU_MAX_SPECIES) {
    const int sidx = threadIdx.x - 32;
    const double dts = (sp.timeStepMode == AMPS_DT_SPECIES_GLOBAL) ? sp.dt[sidx] : sp.dt[0];
    sC.qdt2m[sidx] = (sidx < sp.n) ? 0.5 * (sp.charge[sidx] * dts / sp.mass[sidx]) : 0.0;
    sC.dt[sidx] = dts;
  }
  __syncthreads();
  if (kSmemTiles) mbar_wait(&mbar, 0);

  const int CS0 = kCube10 ? 11 : 1 + m.TN[0], CS1 = kCube10 ? 121 : (1 + m.TN[0]) * (1 + m.TN[1]);  // corner strides
  const int BS0 = kCube10 ? 10 : m.TN[0], BS1 = kCube10 ? 100 : m.TN[0] * m.TN[1];                  // centre strides
  const bool openFace = (!m.periodic) && lg.face != 0;
  // a particle that stays in its block needs no node table: the flags of the start block are read once per CTA
  const int startFlags = m.nodeFlags[lg.node];

  unsigned int nMoved = 0, nXCell = 0, nXBlock = 0, nWrap = 0, nRedo = 0;

  // software pipeline: the loads of the NEXT particle are in flight while the current one is pushed (the kernel is
  // bound by the latency of these loads otherwise: one thread owns every 256th particle of the slice)
  double nx0 = 0.0, nx1 = 0.0, nx2 = 0.0, nv0 = 0.0, nv1 = 0.0, nv2 = 0.0;
  int nspec = 0, nkey = 0;
  {
    const int ip = b + threadIdx.x;
    if (ip < e) {
      nx0 = p.x[0][ip], nx1 = p.x[1][ip], nx2 = p.x[2][ip];
      nv0 = p.v[0][ip], nv1 = p.v[1][ip], nv2 = p.v[2][ip];
      nspec = p.spec[ip], nkey = p.key[ip];
    }
  }
  for (int ip = b + threadIdx.x; ip < e; ip += blockDim.x) {
    double x0 = nx0, x1 = nx1, x2 = nx2;
    const double v0 = nv0, v1 = nv1, v2 = nv2;
    const int spec = nspec & 0x3f;
    const int oldKey = nkey;
    {
      const int np = ip + blockDim.x;
      if (np < e) {
        nx0 = p.x[0][np], nx1 = p.x[1][np], nx2 = p.x[2][np];
        nv0 = p.v[0][np], nv1 = p.v[1][np], nv2 = p.v[2][np];
        nspec = p.spec[np], nkey = p.key[np];
      }
    }
    bool redo = false;

    // ---- a3: corner stencil (the snap to xmax-1e-10dx compares inputs only: same decision as the reference) ----
    double xl[3];
    int iX[3];
    {
      double xx[3] = {x0, x1, x2};
#pragma unroll
      for (int d = 0; d < 3; d++) {
        if ((xx[d] < lg.xmin[d]) || (xx[d] > lg.xmax[d])) redo = true;
        if (fabs(xx[d] - lg.xmax[d]) < 1e-10 * sC.dxc[d]) xx[d] = lg.xmax[d] - 1e-10 * sC.dxc[d];
        const double t = (xx[d] - lg.xmin[d]) * sC.invDxc[d];
        iX[d] = (int)t;
        xl[d] = t - iX[d];
      }
      x0 = xx[0], x1 = xx[1], x2 = xx[2];
    }
    double E0 = 0.0, E1 = 0.0, E2 = 0.0, B0 = 0.0, B1 = 0.0, B2 = 0.0;
    if (!redo) {
      const double ax0 = 1.0 - xl[0], ax1 = xl[0], ay0 = 1.0 - xl[1], ay1 = xl[1], az0 = 1.0 - xl[2], az1 = xl[2];
      const double a00 = ax0 * ay0, a01 = ax0 * ay1, a10 = ax1 * ay0, a11 = ax1 * ay1;
      const double w[8] = {a00 * az0, a00 * az1, a01 * az0, a01 * az1, a10 * az0, a10 * az1, a11 * az0, a11 * az1};
      const int nd0 = kCube10 ? (iX[0] + 1 + 11 * (iX[1] + 1 + (iX[2] + 1) * 11)) : cornerLocalNumber(m, iX[0], iX[1], iX[2]);
#pragma unroll
      for (int s = 0; s < 8; s++) {
        const int nd = nd0 + ((s >> 2) & 1) + ((s >> 1) & 1) * CS0 + (s & 1) * CS1;
        const double *t = sE + 3 * nd;
        E0 = fma(w[s], t[0], E0), E1 = fma(w[s], t[1], E1), E2 = fma(w[s], t[2], E2);
        if (kCornerB) {
          const double *tb = sB + 3 * nd;
          B0 = fma(w[s], tb[0], B0), B1 = fma(w[s], tb[1], B1), B2 = fma(w[s], tb[2], B2);
        }
      }
    }
    // ---- a4: cell-centred stencil for B (same-level branch); truncated boundary stencils go to the exact kernel ----
    if (!kCornerB && !redo) {
      const double iLoc = (x0 - lg.xmin[0]) * sC.invSpanN[0], jLoc = (x1 - lg.xmin[1]) * sC.invSpanN[1], kLoc = (x2 - lg.xmin[2]) * sC.invSpanN[2];
      const int i0 = (iLoc < 0.5) ? -1 : (int)(iLoc - 0.50);
      const int j0 = (jLoc < 0.5) ? -1 : (int)(jLoc - 0.50);
      const int k0 = (kLoc < 0.5) ? -1 : (int)(kLoc - 0.50);
      if (openFace && (((lg.face & 1) && i0 < 0) || ((lg.face & 2) && i0 + 1 >= m.N[0]) || ((lg.face & 4) && j0 < 0) ||
                       ((lg.face & 8) && j0 + 1 >= m.N[1]) || ((lg.face & 16) && k0 < 0) || ((lg.face & 32) && k0 + 1 >= m.N[2]))) {
        redo = true;
      } else {
        const double w0 = iLoc - (i0 + 0.5), w1 = jLoc - (j0 + 0.5), w2 = kLoc - (k0 + 0.5);
        const double a00 = (1.0 - w0) * (1.0 - w1), a01 = (1.0 - w0) * w1, a10 = w0 * (1.0 - w1), a11 = w0 * w1;
        const double w[8] = {a00 * (1.0 - w2), a00 * w2, a01 * (1.0 - w2), a01 * w2, a10 * (1.0 - w2), a10 * w2, a11 * (1.0 - w2), a11 * w2};
        const int nd0 = kCube10 ? (i0 + 1 + 10 * (j0 + 1 + (k0 + 1) * 10)) : centerLocalNumber(m, i0, j0, k0);
#pragma unroll
        for (int s = 0; s < 8; s++) {
          const int nd = nd0 + ((s >> 2) & 1) + ((s >> 1) & 1) * BS0 + (s & 1) * BS1;
          const double *t = sB + 3 * nd;
          B0 = fma(w[s], t[0], B0), B1 = fma(w[s], t[1], B1), B2 = fma(w[s], t[2], B2);
        }
      }
    }

    double xf[3], vf[3];
    int newKey = -1, keyLeaf = leaf;  // keyLeaf = newKey / C without the division
    bool wrapped = false;
    if (!redo) {
      // ---- velocity / position update (:1036-1081) ----
      const double beta = sC.qdt2m[spec], dtTotal = sC.dt[spec];
      const double s2 = beta * beta;
      const double P0 = -beta * B0, P1 = -beta * B1, P2 = -beta * B2;
      const double c0 = __drcp_rn(1.0 + s2 * (B0 * B0 + B1 * B1 + B2 * B2));
      const double u0 = v0 + beta * E0, u1 = v1 + beta * E1, u2 = v2 + beta * E2;
      const double bu = s2 * (B0 * u0 + B1 * u1 + B2 * u2);  // beta^2 B (B.u)
      // alpha u = c0 ( u - P x u ... ) written out: alpha = c0 (I + [P]x' + beta^2 B B^T), rows as in the reference
      const double r0 = c0 * (u0 + (-P2 * u1 + P1 * u2) + bu * B0);
      const double r1 = c0 * (u1 + (P2 * u0 - P0 * u2) + bu * B1);
      const double r2 = c0 * (u2 + (-P1 * u0 + P0 * u1) + bu * B2);
      vf[0] = 2.0 * r0 - v0, vf[1] = 2.0 * r1 - v1, vf[2] = 2.0 * r2 - v2;
      xf[0] = fma(dtTotal, vf[0], x0), xf[1] = fma(dtTotal, vf[1], x1), xf[2] = fma(dtTotal, vf[2], x2);

      // ---- a14: new block ----
      int ix[3];
#pragma unroll
      for (int d = 0; d < 3; d++) ix[d] = (int)floor((xf[d] - m.xGlobalMin[d]) * sC.invRef[d]);
      const bool in = ix[0] >= lg.imin[0] && ix[0] < lg.imin[0] + lg.isize && ix[1] >= lg.imin[1] && ix[1] < lg.imin[1] + lg.isize &&
                      ix[2] >= lg.imin[2] && ix[2] < lg.imin[2] + lg.isize;
      int node = in ? lg.node : find_node_ix(m, ix[0], ix[1], ix[2]);
      // the three node-table entries of a block crosser are requested together (one exposed latency instead of three)
      int nFlags = startFlags, nLeaf = leaf, nLevel = lg.level;
      if (!in && node >= 0) nFlags = m.nodeFlags[node], nLeaf = m.nodeLeaf[node], nLevel = m.nodeLevel[node];
      if (node < 0 || !(nFlags & AMPS_NODE_USED)) redo = true;
      int newLeaf = redo ? -1 : nLeaf;
      if (newLeaf < 0) redo = true;
      if (!redo) {
        int ijk[3];
        const bool sameLevel = in || nLevel == lg.level;
#pragma unroll
        for (int d = 0; d < 3; d++) {
          const double lo = in ? lg.xmin[d] : m.nxmin[3 * node + d];
          const double hi = in ? lg.xmax[d] : m.nxmax[3 * node + d];
          const double inv = sameLevel ? sC.invCell[d] : 1.0 / (m.dxRoot[d] / (1 << nLevel) / double(m.N[d]));
          const double t = (xf[d] - lo) * inv;
          const double fl = floor(t);
          const double fr = t - fl;
          // within GUARD of a cell face (or outside the block found from the approximate position): exact kernel
          if (!(xf[d] > lo && xf[d] < hi) || fr < MOVE_GUARD || fr > 1.0 - MOVE_GUARD) redo = true;
          ijk[d] = (int)fl;
        }
        if (!redo) {
          const int realLeaf = in ? -1 : m.leaf[newLeaf].real;
          if (realLeaf >= 0) {  // a16: periodic ghost block -> real block; the guard keeps x' away from the clamps
            const LeafGeo &gg = m.leaf[newLeaf];
            const LeafGeo &rg = m.leaf[realLeaf];
#pragma unroll
            for (int d = 0; d < 3; d++) xf[d] += rg.xmin[d] - gg.xmin[d];
            newLeaf = realLeaf;
            wrapped = true;
          }
          newKey = newLeaf * C + ijk[0] + m.N[0] * (ijk[1] + m.N[1] * ijk[2]);
          keyLeaf = newLeaf;
        }
      }
    }

    if (redo) {
      redoMask[ip] = 1;
      nRedo++;
      continue;
    }
    nMoved++;
    if (wrapped) nWrap++;
    if (keyLeaf != leaf) nXBlock++;
    else if (newKey != oldKey) nXCell++;
    p.x[0][ip] = xf[0], p.x[1][ip] = xf[1], p.x[2][ip] = xf[2];
    p.v[0][ip] = vf[0], p.v[1][ip] = vf[1], p.v[2][ip] = vf[2];
    atomicAdd(&cellCount[newKey], 1);
    if (newKey != oldKey) p.key[ip] = newKey;
  }

  // one counter per leaf tells the exact kernel whether it has anything to do there
  {
    unsigned v = nRedo;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) {
      if (atomicAdd(&leafRedo[leaf], (int)v) == 0) redoLeafList[atomicAdd(nRedoLeaves, 1)] = leaf;  // first flag in this block
    }
  }
  flush_move_counters(stats, nMoved, nXCell, nXBlock, 0, 0, nWrap, 0);
}

void launch_move_lapenta_fast(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *cellStart, const double *eTile, const double *bTile,
                              int *cellCount, DevMoveStats *stats, int slices, unsigned char *redoMask, int *leafRedo, int *redoLeafList,
                              int *nRedoLeaves, cudaStream_t s) {
  const size_t smem = (size_t)(m.eTileStride + m.bTileStride) * sizeof(double);
  const int grid = m.nLeaves * slices;
  const bool cornerB = sp.bMode == AMPS_B_CORNER_BASED;
  const bool cube10 = m.TN[0] == 10 && m.TN[1] == 10 && m.TN[2] == 10 && m.g[0] == 1 && m.g[1] == 1 && m.g[2] == 1;
#define AMPS_FAST_ARGS m, sp, p, cellStart, eTile, bTile, cellCount, stats, slices, redoMask, leafRedo, redoLeafList, nRedoLeaves
  if (smem <= 200 * 1024) {
    static OncePerDevice once;
    if (once.first()) {
      cudaFuncSetAttribute(move_lapenta_fast_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(move_lapenta_fast_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(move_lapenta_fast_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(move_lapenta_fast_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    if (cube10) {
      if (cornerB) move_lapenta_fast_kernel<true, true, true><<<grid, 256, smem, s>>>(AMPS_FAST_ARGS);
      else move_lapenta_fast_kernel<true, false, true><<<grid, 256, smem, s>>>(AMPS_FAST_ARGS);
    } else {
      if (cornerB) move_lapenta_fast_kernel<true, true, false><<<grid, 256, smem, s>>>(AMPS_FAST_ARGS);
      else move_lapenta_fast_kernel<true, false, false><<<grid, 256, smem, s>>>(AMPS_FAST_ARGS);
    }
  } else if (cornerB) {
    move_lapenta_fast_kernel<false, true, false><<<grid, 256, 0, s>>>(AMPS_FAST_ARGS);
  } else {
    move_lapenta_fast_kernel<false, false, false><<<grid, 256, 0, s>>>(AMPS_FAST_ARGS);
  }
#undef AMPS_FAST_ARGS
}

}  // namespace amps
