// deposit.cu -- ECSIM current + mass-matrix deposition (sm_100a, fp64, FMA contraction allowed:
// results are compared to the reference within 1e-10 relative, summation order differs anyway).
//
//   a10 ECSIM::ProcessCell          src/pic/pic_field_solver_ecsim.cpp:1881-2438
//   a11 ECSIM::UpdateJMassMatrix    src/pic/pic_field_solver_ecsim.cpp:3244-3995
//   a12 ProcessJMassMatrix          src/pic/pic_field_solver_ecsim.cpp:1383-1395 (implicit: all copies of a
//                                   shared / periodic corner are ONE unique corner on the device)
//
// Per cell the mass matrix is a small fp64 contraction over the cell's particles
//        MM[pair(c,c')][3x3] = sum_p (W_c W_c')_p * (k alpha)_p ,   k = q~ beta / V,   36 pairs c' <= c
//        J [c][3]            = sum_p  W_c,p * (q~ alpha v / V)_p
// i.e. 348 accumulators per cell fed by 20 numbers per particle.  The kernel is bound by the fp64
// pipe, not by HBM (57 B/particle in, ~4 KB/cell out), so the layout is chosen to keep the DFMA
// pipe fed with few shared-memory wavefronts and almost no cross-lane reduction:
//   phase 1  thread <-> particle: B gather, alpha, corner weights -> one 176-byte row per particle in
//            shared memory (row stride 22 doubles: 16-byte vector accesses are bank-conflict free)
//   phase 2  108 "MM" threads = 18 register tiles (2 pairs x 9) x 6 particle slices; per particle a
//            thread issues 4 LDS.128 + 5 LDS.64 (alpha row broadcast across the tiles of a slice) for
//            2 DMUL + 18 DFMA.  16 "J" threads = 8 corners x 2 slices accumulate the current.
//   reduce   the 6 (2) slice partials meet in shared memory, thread o sums output o
//   flush    one fp64 RED per value into the unique-corner arrays J[nCorners][3], M[nCorners][243]
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
constexpr int DEP_CHUNK = 192;   // particles staged per pass (a 128-ppc cell fits in one pass)
constexpr int ROW = 22;          // doubles per particle row: W[0..7] ka[8..16] pad qv[18..20] pad
constexpr int OFF_A = 8, OFF_QV = 18;
constexpr int N_TILES = 18, N_SLICES = 6, N_MM = N_TILES * N_SLICES;  // 108 mass-matrix threads
constexpr int N_JSL = 2, J_FIRST = N_MM, N_J = 8 * N_JSL;            // 16 current threads (108..123)
constexpr int N_OUT = 324 + 24;
constexpr int RED_STRIDE = 352;  // doubles per slice in the reduction buffer

__device__ __forceinline__ void atomicMaxPositiveDouble(unsigned long long *addr, double v) {
  // v >= 0 and not NaN: the bit patterns of non-negative doubles order like unsigned integers
  atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

__global__ void __launch_bounds__(DEP_THREADS, 4) deposit_kernel(DevMesh m, DevSpecies sp, ParticleSoA p, const int *__restrict__ cellStart,
                                                                const double *__restrict__ bCurTile, double *__restrict__ J,
                                                                double *__restrict__ M, double *__restrict__ energyOut,
                                                                unsigned long long *__restrict__ cflBits) {
  // sF doubles as the reduction buffer after phase 2 (N_SLICES*RED_STRIDE <= DEP_CHUNK*ROW)
  __shared__ __align__(16) double sF[DEP_CHUNK * ROW];
  __shared__ double sB[27 * 3];  // B_cur on the 3x3x3 centres around the cell
  __shared__ double sRedV[DEP_THREADS / 32][AMPS_GPU_MAX_SPECIES];
  __shared__ int sRedC[DEP_THREADS / 32][AMPS_GPU_MAX_SPECIES];
  __shared__ int sUid[8];
  static_assert(N_SLICES * RED_STRIDE <= DEP_CHUNK * ROW, "reduction buffer must fit in the factor buffer");

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int C = m.cellsPerBlock;
  const long long nCells = (long long)m.nLeaves * C;

  // fixed phase-2 role of this thread
  const bool mmThread = t < N_MM, jThread = (t >= J_FIRST && t < J_FIRST + N_J);
  int tile = 0, slice = 0;
  int wi0 = 0, wj0 = 0, wi1 = 0, wj1 = 0;  // row offsets of the W operands of the two pairs
  if (mmThread) {
    tile = t % N_TILES, slice = t / N_TILES;
    wi0 = cPairI[2 * tile], wj0 = cPairJ[2 * tile], wi1 = cPairI[2 * tile + 1], wj1 = cPairJ[2 * tile + 1];
  } else if (jThread) {
    tile = (t - J_FIRST) % 8, slice = (t - J_FIRST) / 8;  // tile = corner
  }

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

    double dx[3], dxc[3], invdxc[3], xmn[3], xmx[3];
    double CellVolume = 1;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      xmn[d] = lg.xmin[d], xmx[d] = lg.xmax[d];
      dxc[d] = (xmx[d] - xmn[d]) / m.N[d];
      dx[d] = dxc[d] * sp.length_conv;
      invdxc[d] = 1.0 / dxc[d];
    }
#pragma unroll
    for (int d = 0; d < 3; d++) CellVolume *= dx[d];
    const double invV = 1.0 / CellVolume;
    const double invc = 1.0 / sp.LightSpeed;
    const int face = m.periodic ? 0 : lg.face;

    __syncthreads();  // previous cell fully consumed (sF, sB, sRed*, sUid)
    // stage the 27 centre values of B_cur the cell's stencils can touch, and the 8 corner ids
    if (t < 81) {
      const int n = t / 3, d = t - 3 * n;
      const int di = n % 3 - 1, dj = (n / 3) % 3 - 1, dk = n / 9 - 1;
      const double *bT = bCurTile + (size_t)leaf * m.bTileStride;
      sB[t] = __ldg(bT + 3 * centerLocalNumber(m, ic + di, jc + dj, kc + dk) + d);
    } else if (t >= 96 && t < 104) {
      const int c = t - 96;
      const int *uidT = m.cornerUid + (size_t)leaf * m.nCornerLocal;
      sUid[c] = uidT[cornerLocalNumber(m, ic + cCornerOff[c][0], jc + cCornerOff[c][1], kc + cCornerOff[c][2])];
    }

    double acc[18];
#pragma unroll
    for (int i = 0; i < 18; i++) acc[i] = 0.0;
    double vmSpec[AMPS_GPU_MAX_SPECIES];
    int cntSpec[AMPS_GPU_MAX_SPECIES];
#pragma unroll
    for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++) vmSpec[s] = 0.0, cntSpec[s] = 0;

    for (int base = begin; base < end; base += DEP_CHUNK) {
      const int np = min(DEP_CHUNK, end - base);
      __syncthreads();  // sB/sUid visible; previous chunk consumed
      // ---------------- phase 1: per-particle factors ----------------
      for (int q = t; q < np; q += DEP_THREADS) {
        const int ip = base + q;
        const double x0 = p.x[0][ip], x1 = p.x[1][ip], x2 = p.x[2][ip];
        double v0 = p.v[0][ip], v1 = p.v[1][ip], v2 = p.v[2][ip];
        const int spec = p.spec[ip];
        const double LocalParticleWeight = sp.weight[spec] * p.w[ip];
        double *row = sF + q * ROW;
        // local coordinates: CornerBased::InitStencil (pic_interpolation_routines.cpp:1090-1098)
        double xl[3];
        {
          const double xx[3] = {x0, x1, x2};
#pragma unroll
          for (int d = 0; d < 3; d++) {
            double xs = xx[d];
            if (fabs(xs - xmx[d]) < 1e-10 * dxc[d]) xs = xmx[d] - 1e-10 * dxc[d];
            double r = (xs - xmn[d]) * invdxc[d];
            r -= (int)r;
            xl[d] = r;
          }
        }
        {
          const double ax0 = 1.0 - xl[0], ax1 = xl[0], ay0 = 1.0 - xl[1], ay1 = xl[1], az0 = 1.0 - xl[2], az1 = xl[2];
          const double a00 = ax0 * ay0, a10 = ax1 * ay0, a11 = ax1 * ay1, a01 = ax0 * ay1;
          reinterpret_cast<double2 *>(row)[0] = make_double2(a00 * az0, a10 * az0);
          reinterpret_cast<double2 *>(row)[1] = make_double2(a11 * az0, a01 * az0);
          reinterpret_cast<double2 *>(row)[2] = make_double2(a00 * az1, a10 * az1);
          reinterpret_cast<double2 *>(row)[3] = make_double2(a11 * az1, a01 * az1);
        }
        // B at the particle: cell-centred trilinear stencil on B_cur (:2100-2129); relative to this cell the
        // stencil cells are -1/0 (particle in the lower half) or 0/+1 (upper half) per dimension
        double B0 = 0.0, B1 = 0.0, B2 = 0.0;
        {
          int o[3];
          double w[3];
#pragma unroll
          for (int d = 0; d < 3; d++) {
            o[d] = (xl[d] < 0.5) ? 0 : 1;
            w[d] = xl[d] + 0.5 - (double)o[d];  // weight of the upper stencil cell
          }
          double ws[8];
          {
            const double b00 = (1.0 - w[0]) * (1.0 - w[1]), b01 = (1.0 - w[0]) * w[1], b10 = w[0] * (1.0 - w[1]), b11 = w[0] * w[1];
            ws[0] = b00 * (1.0 - w[2]), ws[1] = b00 * w[2], ws[2] = b01 * (1.0 - w[2]), ws[3] = b01 * w[2];
            ws[4] = b10 * (1.0 - w[2]), ws[5] = b10 * w[2], ws[6] = b11 * (1.0 - w[2]), ws[7] = b11 * w[2];
          }
          double inv = 1.0;
          if (face) {  // AddCell drops centres outside the global box (pic.h:7235-7245), the rest is re-normalised
            unsigned valid = 0xffu;
            const int i0 = ic - 1 + o[0], j0 = jc - 1 + o[1], k0 = kc - 1 + o[2];
            if ((face & 1) && i0 < 0) valid &= 0xf0u;
            if ((face & 2) && i0 + 1 >= m.N[0]) valid &= 0x0fu;
            if ((face & 4) && j0 < 0) valid &= 0xccu;
            if ((face & 8) && j0 + 1 >= m.N[1]) valid &= 0x33u;
            if ((face & 16) && k0 < 0) valid &= 0xaau;
            if ((face & 32) && k0 + 1 >= m.N[2]) valid &= 0x55u;
            double norm = 0.0;
#pragma unroll
            for (int s = 0; s < 8; s++) {
              if (!(valid & (1u << s))) ws[s] = 0.0;
              norm += ws[s];
            }
            inv = (norm > 0.0) ? 1.0 / norm : 1.0;
          }
          // (full stencil: the weights sum to 1 within 2 ulp, Normalize() changes B by <= 3e-16 relative)
          const int n0 = o[0] + 3 * o[1] + 9 * o[2];
#pragma unroll
          for (int s = 0; s < 8; s++) {
            const int n = n0 + ((s >> 2) & 1) + 3 * ((s >> 1) & 1) + 9 * (s & 1);
            B0 = fma(ws[s], sB[3 * n], B0);
            B1 = fma(ws[s], sB[3 * n + 1], B1);
            B2 = fma(ws[s], sB[3 * n + 2], B2);
          }
          const double sc = inv * sp.B_conv * invc;  // B *= B_conv (:2134); B /= LightSpeed (:2170)
          B0 *= sc, B1 *= sc, B2 *= sc;
        }
        v0 *= sp.length_conv, v1 *= sp.length_conv, v2 *= sp.length_conv;
        const double chargeQ = sp.charge[spec] * LocalParticleWeight;
        const double mass = sp.mass[spec] * LocalParticleWeight;
        const double QdT_over_2m = 0.5 * (chargeQ * sp.dtTotal / mass);
        const double s2 = QdT_over_2m * QdT_over_2m;
        const double P0 = -QdT_over_2m * B0, P1 = -QdT_over_2m * B1, P2 = -QdT_over_2m * B2;
        const double c0 = 1.0 / (1.0 + s2 * (B0 * B0 + B1 * B1 + B2 * B2));
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
        reinterpret_cast<double2 *>(row + OFF_A)[0] = make_double2(kk * al[0], kk * al[1]);
        reinterpret_cast<double2 *>(row + OFF_A)[1] = make_double2(kk * al[2], kk * al[3]);
        reinterpret_cast<double2 *>(row + OFF_A)[2] = make_double2(kk * al[4], kk * al[5]);
        reinterpret_cast<double2 *>(row + OFF_A)[3] = make_double2(kk * al[6], kk * al[7]);
        row[OFF_A + 8] = kk * al[8];
        const double qV = chargeQ * invV;  // Jg/CellVolume (:2367)
        reinterpret_cast<double2 *>(row + OFF_QV)[0] =
            make_double2(qV * (al[0] * v0 + al[1] * v1 + al[2] * v2), qV * (al[3] * v0 + al[4] * v1 + al[5] * v2));
        row[OFF_QV + 2] = qV * (al[6] * v0 + al[7] * v1 + al[8] * v2);

        const double vsqr = v0 * v0 + v1 * v1 + v2 * v2;
        energyThread += 0.5 * mass * vsqr;
        const double vabs = sqrt(vsqr) * sp.dt[0];
#pragma unroll
        for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++)
          if (s == spec) vmSpec[s] += vabs, cntSpec[s]++;
      }
      __syncthreads();
      // ---------------- phase 2: register-tile accumulation ----------------
      if (mmThread) {
#pragma unroll 2
        for (int q = slice; q < np; q += N_SLICES) {
          const double *row = sF + q * ROW;
          const double2 a01 = reinterpret_cast<const double2 *>(row + OFF_A)[0];
          const double2 a23 = reinterpret_cast<const double2 *>(row + OFF_A)[1];
          const double2 a45 = reinterpret_cast<const double2 *>(row + OFF_A)[2];
          const double2 a67 = reinterpret_cast<const double2 *>(row + OFF_A)[3];
          const double a8 = row[OFF_A + 8];
          const double u0 = row[wi0] * row[wj0];
          const double u1 = row[wi1] * row[wj1];
          acc[0] = fma(u0, a01.x, acc[0]);
          acc[1] = fma(u0, a01.y, acc[1]);
          acc[2] = fma(u0, a23.x, acc[2]);
          acc[3] = fma(u0, a23.y, acc[3]);
          acc[4] = fma(u0, a45.x, acc[4]);
          acc[5] = fma(u0, a45.y, acc[5]);
          acc[6] = fma(u0, a67.x, acc[6]);
          acc[7] = fma(u0, a67.y, acc[7]);
          acc[8] = fma(u0, a8, acc[8]);
          acc[9] = fma(u1, a01.x, acc[9]);
          acc[10] = fma(u1, a01.y, acc[10]);
          acc[11] = fma(u1, a23.x, acc[11]);
          acc[12] = fma(u1, a23.y, acc[12]);
          acc[13] = fma(u1, a45.x, acc[13]);
          acc[14] = fma(u1, a45.y, acc[14]);
          acc[15] = fma(u1, a67.x, acc[15]);
          acc[16] = fma(u1, a67.y, acc[16]);
          acc[17] = fma(u1, a8, acc[17]);
        }
      } else if (jThread) {
        for (int q = slice; q < np; q += N_JSL) {
          const double *row = sF + q * ROW;
          const double wq = row[tile];
          const double2 q01 = reinterpret_cast<const double2 *>(row + OFF_QV)[0];
          acc[0] = fma(wq, q01.x, acc[0]);
          acc[1] = fma(wq, q01.y, acc[1]);
          acc[2] = fma(wq, row[OFF_QV + 2], acc[2]);
        }
      }
    }

    // ---------------- reduce the slice partials through shared memory ----------------
    __syncthreads();  // phase 2 done reading sF
    if (mmThread) {
      double *r = sF + slice * RED_STRIDE + tile * 18;
#pragma unroll
      for (int i = 0; i < 18; i++) r[i] = acc[i];
    } else if (jThread) {
      double *r = sF + slice * RED_STRIDE + 324 + tile * 3;
      r[0] = acc[0], r[1] = acc[1], r[2] = acc[2];
    }
    __syncthreads();
    // ---------------- flush: thread o owns outputs o, o+128, o+256 ----------------
#pragma unroll
    for (int rr = 0; rr < 3; rr++) {
      const int o = t + rr * DEP_THREADS;
      if (o < 324) {
        double val = 0.0;
#pragma unroll
        for (int s = 0; s < N_SLICES; s++) val += sF[s * RED_STRIDE + o];
        const int pair = o / 9, k = o - 9 * pair;
        const int ci = cPairI[pair], cj = cPairJ[pair];
        atomicAdd(M + (size_t)sUid[ci] * 243 + 9 * cIndexMatrix[ci][cj] + k, val);
        if (ci != cj) atomicAdd(M + (size_t)sUid[cj] * 243 + 9 * cIndexMatrix[cj][ci] + k, val);
      } else if (o < N_OUT) {
        const double val = sF[o] + sF[RED_STRIDE + o];
        const int c = (o - 324) / 3, k = (o - 324) - 3 * c;
        atomicAdd(J + (size_t)sUid[c] * 3 + k, val);
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
      if (lane == 0) sRedV[warp][s] = vm, sRedC[warp][s] = c;
    }
    __syncthreads();
    if (t < sp.n) {
      double vm = 0.0;
      int c = 0;
      for (int w = 0; w < DEP_THREADS / 32; w++) vm += sRedV[w][t], c += sRedC[w][t];
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
  long long grid = 148LL * 4 * 8;
  if (grid > nCells) grid = nCells;
  deposit_kernel<<<(int)grid, DEP_THREADS, 0, s>>>(m, sp, p, cellStart, bCurTile, J, M, energy, cflBits);
  (*launches)++;
}

}  // namespace amps
