// deposit.cu -- ECSIM current + mass-matrix deposition (sm_100a, fp64, FMA contraction allowed:
// results are compared to the reference within 1e-10 relative, summation order differs anyway).
//
//   a10 ECSIM::ProcessCell          src/pic/pic_field_solver_ecsim.cpp:1881-2438
//   a11 ECSIM::UpdateJMassMatrix    src/pic/pic_field_solver_ecsim.cpp:3244-3995
//   a12 ProcessJMassMatrix          src/pic/pic_field_solver_ecsim.cpp:1383-1395 (implicit: all copies of a
//                                   shared / periodic corner are ONE unique corner on the device)
//
// Per cell the mass matrix is a small fp64 contraction
//        MM[pair(c,c')][3x3] = sum_p (W_c W_c')_p * (k alpha)_p ,   k = q~ beta / V,   36 pairs c' <= c
//        J [c][3]            = sum_p  W_c,p * (q~ alpha v / V)_p
// i.e. 348 accumulators per cell fed by 20 numbers per particle.  The kernel is bound by the fp64
// pipe, not by HBM (57 B/particle in, ~4 KB/cell out), so everything is arranged to issue DFMAs
// from registers:
//   phase 1  thread <-> particle: B gather, alpha, corner weights -> 20 doubles/particle in shared memory
//   phase 2  warp w owns pairs 9w..9w+8 and corners 2w,2w+1 (87 accumulators per lane, compile-time
//            register tile); lane <-> particle slot, so each lane streams its particles' factors from
//            shared memory (unit stride, conflict free) and issues 9 DMUL + 87 DFMA per particle
//   flush    "halving" butterfly: 5 shuffle rounds reduce 96 registers x 32 lanes to 3 totals per lane
//            (93 DADD instead of 5 x 87), then one fp64 RED per value into the unique-corner arrays.
#include "amps_dev.cuh"

namespace amps {

__constant__ int cIndexMatrix[8][8] = {{0, 2, 8, 6, 18, 20, 26, 24},  {1, 0, 6, 7, 19, 18, 24, 25},
                                       {4, 3, 0, 1, 22, 21, 18, 19},  {3, 5, 2, 0, 21, 23, 20, 18},
                                       {9, 11, 17, 15, 0, 2, 8, 6},   {10, 9, 15, 16, 1, 0, 6, 7},
                                       {13, 12, 9, 10, 4, 3, 0, 1},   {12, 14, 11, 9, 3, 5, 2, 0}};
// cell-corner order (0,0,0)(1,0,0)(1,1,0)(0,1,0)(0,0,1)(1,0,1)(1,1,1)(0,1,1)
__constant__ int cCornerOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
__constant__ int cPairI[36] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 7, 7};
__constant__ int cPairJ[36] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 6, 0, 1, 2, 3, 4, 5, 6, 7};

constexpr int DEP_THREADS = 128;
constexpr int DEP_WARPS = DEP_THREADS / 32;
constexpr int DEP_CHUNK = 256;  // particles staged per pass (a cell normally fits in one pass)
constexpr int NF = 20;          // factors per particle: W[8], k*alpha[9], q~ alpha v / V [3]
constexpr int NACC = 96;        // 81 mass-matrix + 6 current accumulators per lane, padded to 3*32

__host__ __device__ constexpr int pair_i(int p) {
  int i = 0;
  while ((i + 1) * (i + 2) / 2 <= p) i++;
  return i;
}
__host__ __device__ constexpr int pair_j(int p) { return p - pair_i(p) * (pair_i(p) + 1) / 2; }

__device__ __forceinline__ void atomicMaxPositiveDouble(unsigned long long *addr, double v) {
  // v >= 0 and not NaN: the bit patterns of non-negative doubles order like unsigned integers
  atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

// phase 2 for warp role WARP: pairs 9*WARP .. 9*WARP+8 and corners 2*WARP, 2*WARP+1
template <int WARP>
__device__ __forceinline__ void accumulate(const double *__restrict__ sF, int nsub, int lane, double (&acc)[NACC]) {
  for (int j = 0; j < nsub; j++) {
    const int q = lane + 32 * j;
    double W[8], a[9], qv[3];
#pragma unroll
    for (int c = 0; c < 8; c++) W[c] = sF[c * DEP_CHUNK + q];
#pragma unroll
    for (int k = 0; k < 9; k++) a[k] = sF[(8 + k) * DEP_CHUNK + q];
#pragma unroll
    for (int d = 0; d < 3; d++) qv[d] = sF[(17 + d) * DEP_CHUNK + q];
#pragma unroll
    for (int pp = 0; pp < 9; pp++) {
      constexpr int base = WARP * 9;
      const double u = W[pair_i(base + pp)] * W[pair_j(base + pp)];
#pragma unroll
      for (int k = 0; k < 9; k++) acc[pp * 9 + k] = fma(u, a[k], acc[pp * 9 + k]);
    }
#pragma unroll
    for (int c2 = 0; c2 < 2; c2++)
#pragma unroll
      for (int d = 0; d < 3; d++) acc[81 + c2 * 3 + d] = fma(W[2 * WARP + c2], qv[d], acc[81 + c2 * 3 + d]);
  }
}

// butterfly with halving: after the 5 rounds lane l holds the totals of outputs base(l)+{0,1,2},
// base(l) = 48*b4 + 24*b3 + 12*b2 + 6*b1 + 3*b0
template <int H>
__device__ __forceinline__ void halve(double (&v)[NACC], int lane, int offset) {
  const bool hi = (lane & offset) != 0;
#pragma unroll
  for (int i = 0; i < H; i++) {
    const double send = hi ? v[i] : v[i + H];
    const double keep = hi ? v[i + H] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, offset);
  }
}

__global__ void __launch_bounds__(DEP_THREADS, 2) deposit_kernel(DevMesh m, DevSpecies sp, ParticleSoA p, const int *__restrict__ cellStart,
                                                                const double *__restrict__ bCurTile, double *__restrict__ J,
                                                                double *__restrict__ M, double *__restrict__ energyOut,
                                                                unsigned long long *__restrict__ cflBits) {
  __shared__ double sF[NF * DEP_CHUNK];
  __shared__ double sB[27 * 3];  // B_cur on the 3x3x3 centres around the cell
  __shared__ double sRed[DEP_WARPS][AMPS_GPU_MAX_SPECIES];
  __shared__ int sCnt[DEP_WARPS][AMPS_GPU_MAX_SPECIES];

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int C = m.cellsPerBlock;
  const long long nCells = (long long)m.nLeaves * C;

  double energyThread = 0.0;
  double cflThread = 0.0;  // thread s < n keeps the running max of species s

  for (long long cell = blockIdx.x; cell < nCells; cell += gridDim.x) {
    const int begin = cellStart[cell], end = cellStart[cell + 1];
    if (begin == end) continue;  // ProcessCell returns false: nothing is flushed
    const int leaf = (int)(cell / C);
    const int cin = (int)(cell - (long long)leaf * C);
    const LeafGeo &lg = m.leaf[leaf];
    if (m.periodic && lg.face != 0) continue;  // periodic "ghost" (boundary) blocks are skipped, :3815-3825

    const int kc = cin / (m.N[0] * m.N[1]);
    const int jc = (cin - kc * m.N[0] * m.N[1]) / m.N[0];
    const int ic = cin - kc * m.N[0] * m.N[1] - jc * m.N[0];

    double dx[3], invdxc[3], xmn[3], xmx[3];
    double CellVolume = 1;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      xmn[d] = lg.xmin[d], xmx[d] = lg.xmax[d];
      const double dxc = (xmx[d] - xmn[d]) / m.N[d];
      dx[d] = dxc * sp.length_conv;
      invdxc[d] = 1.0 / dxc;
    }
#pragma unroll
    for (int d = 0; d < 3; d++) CellVolume *= dx[d];
    const double invV = 1.0 / CellVolume;
    const int face = m.periodic ? 0 : lg.face;

    __syncthreads();  // previous cell fully consumed (sF, sB, sRed)
    // stage the 27 centre values of B_cur the cell's stencils can touch
    if (t < 81) {
      const int n = t / 3, d = t - 3 * n;
      const int di = n % 3 - 1, dj = (n / 3) % 3 - 1, dk = n / 9 - 1;
      const double *bT = bCurTile + (size_t)leaf * m.bTileStride;
      sB[t] = __ldg(bT + 3 * centerLocalNumber(m, ic + di, jc + dj, kc + dk) + d);
    }

    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = 0.0;
    double vmean[2] = {0.0, 0.0};  // per-thread |v| sums are kept per species via the loop below
    double vmSpec[AMPS_GPU_MAX_SPECIES];
    int cntSpec[AMPS_GPU_MAX_SPECIES];
    (void)vmean;
#pragma unroll
    for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++) vmSpec[s] = 0.0, cntSpec[s] = 0;

    for (int base = begin; base < end; base += DEP_CHUNK) {
      const int np = min(DEP_CHUNK, end - base);
      const int nsub = (np + 31) >> 5;
      if (base != begin) __syncthreads();
      __syncthreads();  // sB visible
      // ---------------- phase 1: per-particle factors ----------------
      for (int q = t; q < nsub * 32; q += DEP_THREADS) {
        double F[NF];
        if (q < np) {
          const int ip = base + q;
          const double x0 = p.x[0][ip], x1 = p.x[1][ip], x2 = p.x[2][ip];
          double v0 = p.v[0][ip], v1 = p.v[1][ip], v2 = p.v[2][ip];
          const int spec = p.spec[ip];
          const double LocalParticleWeight = sp.weight[spec] * p.w[ip];
          // local coordinates: CornerBased::InitStencil (:1090-1098 of pic_interpolation_routines.cpp)
          double xl[3];
          {
            const double xx[3] = {x0, x1, x2};
#pragma unroll
            for (int d = 0; d < 3; d++) {
              double xs = xx[d];
              const double dxc = (xmx[d] - xmn[d]) / m.N[d];
              if (fabs(xs - xmx[d]) < 1e-10 * dxc) xs = xmx[d] - 1e-10 * dxc;
              double r = (xs - xmn[d]) * invdxc[d];
              r -= (int)r;
              xl[d] = r;
            }
          }
          const double ax0 = 1.0 - xl[0], ax1 = xl[0], ay0 = 1.0 - xl[1], ay1 = xl[1], az0 = 1.0 - xl[2], az1 = xl[2];
          F[0] = ax0 * ay0 * az0;
          F[1] = ax1 * ay0 * az0;
          F[2] = ax1 * ay1 * az0;
          F[3] = ax0 * ay1 * az0;
          F[4] = ax0 * ay0 * az1;
          F[5] = ax1 * ay0 * az1;
          F[6] = ax1 * ay1 * az1;
          F[7] = ax0 * ay1 * az1;

          // B at the particle: cell-centred trilinear stencil on B_cur (:2100-2129); the stencil cell
          // offsets relative to this cell are -1/0 (lower half) or 0/+1 (upper half) per dimension
          double B0 = 0.0, B1 = 0.0, B2 = 0.0;
          {
            // iLoc - ic in [0,1): position inside the cell in cell units (same quantity as xl up to rounding)
            int o[3];
            double w[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
              o[d] = (xl[d] < 0.5) ? 0 : 1;           // first stencil cell = cell-1+o
              w[d] = xl[d] + 0.5 - (double)o[d];      // weight of the upper stencil cell
            }
            double ws[8];
            ws[0] = (1.0 - w[0]) * (1.0 - w[1]) * (1.0 - w[2]);
            ws[1] = (1.0 - w[0]) * (1.0 - w[1]) * w[2];
            ws[2] = (1.0 - w[0]) * w[1] * (1.0 - w[2]);
            ws[3] = (1.0 - w[0]) * w[1] * w[2];
            ws[4] = w[0] * (1.0 - w[1]) * (1.0 - w[2]);
            ws[5] = w[0] * (1.0 - w[1]) * w[2];
            ws[6] = w[0] * w[1] * (1.0 - w[2]);
            ws[7] = w[0] * w[1] * w[2];
            unsigned valid = 0xffu;
            if (face) {  // AddCell drops centres outside the global box (pic.h:7235-7245)
              const int i0 = ic - 1 + o[0], j0 = jc - 1 + o[1], k0 = kc - 1 + o[2];
              if ((face & 1) && i0 < 0) valid &= 0xf0u;
              if ((face & 2) && i0 + 1 >= m.N[0]) valid &= 0x0fu;
              if ((face & 4) && j0 < 0) valid &= 0xccu;
              if ((face & 8) && j0 + 1 >= m.N[1]) valid &= 0x33u;
              if ((face & 16) && k0 < 0) valid &= 0xaau;
              if ((face & 32) && k0 + 1 >= m.N[2]) valid &= 0x55u;
            }
            double norm = 0.0;
#pragma unroll
            for (int s = 0; s < 8; s++) {
              if (!(valid & (1u << s))) ws[s] = 0.0;
              norm += ws[s];
            }
            const double inv = (norm > 0.0) ? 1.0 / norm : 1.0;
#pragma unroll
            for (int s = 0; s < 8; s++) {
              const int n = (o[0] + ((s >> 2) & 1)) + 3 * (o[1] + ((s >> 1) & 1)) + 9 * (o[2] + (s & 1));
              const double wn = ws[s] * inv;
              B0 = fma(wn, sB[3 * n], B0);
              B1 = fma(wn, sB[3 * n + 1], B1);
              B2 = fma(wn, sB[3 * n + 2], B2);
            }
          }
          B0 *= sp.B_conv, B1 *= sp.B_conv, B2 *= sp.B_conv;
          v0 *= sp.length_conv, v1 *= sp.length_conv, v2 *= sp.length_conv;
          const double chargeQ = sp.charge[spec] * LocalParticleWeight;
          const double mass = sp.mass[spec] * LocalParticleWeight;
          const double QdT_over_m = chargeQ * sp.dtTotal / mass;
          const double QdT_over_2m = 0.5 * QdT_over_m;
          const double QdT_over_2m_squared = QdT_over_2m * QdT_over_2m;
          const double invc = 1.0 / sp.LightSpeed;
          B0 *= invc, B1 *= invc, B2 *= invc;
          const double P0 = -QdT_over_2m * B0, P1 = -QdT_over_2m * B1, P2 = -QdT_over_2m * B2;
          const double c0 = 1.0 / (1.0 + QdT_over_2m_squared * (B0 * B0 + B1 * B1 + B2 * B2));
          const double s2 = QdT_over_2m_squared;
          double al[9];
          al[0] = c0 * (1.0 + s2 * B0 * B0);
          al[1] = c0 * (-P2 + s2 * B0 * B1);
          al[2] = c0 * (P1 + s2 * B0 * B2);
          al[3] = c0 * (P2 + s2 * B1 * B0);
          al[4] = c0 * (1.0 + s2 * B1 * B1);
          al[5] = c0 * (-P0 + s2 * B1 * B2);
          al[6] = c0 * (-P1 + s2 * B2 * B0);
          al[7] = c0 * (P0 + s2 * B2 * B1);
          al[8] = c0 * (1.0 + s2 * B2 * B2);
          const double kk = chargeQ * QdT_over_2m * invV;  // matrixConst (:2311)
#pragma unroll
          for (int k = 0; k < 9; k++) F[8 + k] = kk * al[k];
          const double qV = chargeQ * invV;  // Jg/CellVolume (:2367)
          F[17] = qV * (al[0] * v0 + al[1] * v1 + al[2] * v2);
          F[18] = qV * (al[3] * v0 + al[4] * v1 + al[5] * v2);
          F[19] = qV * (al[6] * v0 + al[7] * v1 + al[8] * v2);

          const double vsqr = v0 * v0 + v1 * v1 + v2 * v2;
          energyThread += 0.5 * mass * vsqr;
          const double vabs = sqrt(vsqr) * sp.dt[0];
#pragma unroll
          for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++)
            if (s == spec) vmSpec[s] += vabs, cntSpec[s]++;
        } else {
#pragma unroll
          for (int k = 0; k < NF; k++) F[k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < NF; k++) sF[k * DEP_CHUNK + q] = F[k];
      }
      __syncthreads();
      // ---------------- phase 2: register-tile accumulation, role = warp ----------------
      switch (warp) {
        case 0: accumulate<0>(sF, nsub, lane, acc); break;
        case 1: accumulate<1>(sF, nsub, lane, acc); break;
        case 2: accumulate<2>(sF, nsub, lane, acc); break;
        default: accumulate<3>(sF, nsub, lane, acc); break;
      }
    }

    // ---------------- flush ----------------
    halve<48>(acc, lane, 16);
    halve<24>(acc, lane, 8);
    halve<12>(acc, lane, 4);
    halve<6>(acc, lane, 2);
    halve<3>(acc, lane, 1);
    {
      const int *uidT = m.cornerUid + (size_t)leaf * m.nCornerLocal;
      int uidLane = 0;
      if (lane < 8) uidLane = uidT[cornerLocalNumber(m, ic + cCornerOff[lane][0], jc + cCornerOff[lane][1], kc + cCornerOff[lane][2])];
      const int obase = 48 * ((lane >> 4) & 1) + 24 * ((lane >> 3) & 1) + 12 * ((lane >> 2) & 1) + 6 * ((lane >> 1) & 1) + 3 * (lane & 1);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const int o = obase + i;
        int ci = 0, cj = 0, k = 0;
        const bool isM = o < 81, isJ = (o >= 81 && o < 87);
        if (isM) {
          const int pair = warp * 9 + o / 9;
          k = o - 9 * (o / 9);
          ci = cPairI[pair], cj = cPairJ[pair];
        } else if (isJ) {
          ci = 2 * warp + (o - 81) / 3;
          k = (o - 81) % 3;
          cj = ci;
        }
        const int ui = __shfl_sync(0xffffffffu, uidLane, ci);
        const int uj = __shfl_sync(0xffffffffu, uidLane, cj);
        const double val = acc[i];
        if (isM) {
          atomicAdd(M + (size_t)ui * 243 + 9 * cIndexMatrix[ci][cj] + k, val);
          if (ci != cj) atomicAdd(M + (size_t)uj * 243 + 9 * cIndexMatrix[cj][ci] + k, val);
        } else if (isJ) {
          atomicAdd(J + (size_t)ui * 3 + k, val);
        }
      }
    }

    // ---------------- per-cell cfl: vmean[s] / (count[s] * |dx|)  (:2357-2359) ----------------
    for (int s = 0; s < sp.n; s++) {
      double vm = 0.0;
      int c = 0;
#pragma unroll
      for (int s2 = 0; s2 < AMPS_GPU_MAX_SPECIES; s2++)
        if (s2 == s) vm = vmSpec[s2], c = cntSpec[s2];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vm += __shfl_xor_sync(0xffffffffu, vm, o);
      c = __reduce_add_sync(0xffffffffu, c);
      if (lane == 0) sRed[warp][s] = vm, sCnt[warp][s] = c;
    }
    __syncthreads();
    if (t < sp.n) {
      double vm = 0.0;
      int c = 0;
      for (int w = 0; w < DEP_WARPS; w++) vm += sRed[w][t], c += sCnt[w][t];
      if (c > 0) {  // 0/0 = NaN never wins the reference's '>' comparison
        const double cfl = vm / (c * sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]));
        if (cfl > cflThread) cflThread = cfl;
      }
    }
  }

  // energy: the reference adds the cell energy once per corner => x8 (:3860)
  {
    double e = energyThread;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0 && e != 0.0) atomicAdd(energyOut, 8.0 * e);
    if (t < sp.n && cflThread > 0.0) atomicMaxPositiveDouble(&cflBits[t], cflThread);
  }
}

void launch_deposit(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *cellStart, const double *bCurTile, double *J, double *M,
                    double *energy, unsigned long long *cflBits, cudaStream_t s, long long *launches) {
  const long long nCells = (long long)m.nLeaves * m.cellsPerBlock;
  // zero J, M (SetCornerNodeAssociatedDataValue, :3266-3267) and the diagnostics
  cudaMemsetAsync(J, 0, sizeof(double) * 3 * (size_t)m.nCorners, s);
  cudaMemsetAsync(M, 0, sizeof(double) * 243 * (size_t)m.nCorners, s);
  cudaMemsetAsync(energy, 0, sizeof(double), s);
  cudaMemsetAsync(cflBits, 0, sizeof(unsigned long long) * AMPS_GPU_MAX_SPECIES, s);
  long long grid = 148LL * 2 * 8;
  if (grid > nCells) grid = nCells;
  deposit_kernel<<<(int)grid, DEP_THREADS, 0, s>>>(m, sp, p, cellStart, bCurTile, J, M, energy, cflBits);
  (*launches)++;
}

}  // namespace amps
