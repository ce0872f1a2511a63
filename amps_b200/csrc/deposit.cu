// deposit.cu -- ECSIM current + mass-matrix deposition (sm_100a, fp64, FMA contraction allowed:
// results are compared to the reference within 1e-10 relative, summation order differs anyway).
//
//   a10 ECSIM::ProcessCell          src/pic/pic_field_solver_ecsim.cpp:1881-2438
//   a11 ECSIM::UpdateJMassMatrix    src/pic/pic_field_solver_ecsim.cpp:3244-3995
//   a12 ProcessJMassMatrix          src/pic/pic_field_solver_ecsim.cpp:1383-1395 (implicit: all copies of a
//                                   shared / periodic corner are ONE unique corner on the device)
//
// Per cell the mass matrix is a small contraction  MM[36 pairs][9] = sum_p u_p (x) alpha_p,
// u_p[c,c'] = (q~ beta / V) W_c W_c'.  A CTA owns one cell at a time: phase 1 computes the per
// particle factors (B gather, alpha, W, q~ alpha v) once into shared memory, phase 2 gives every
// thread a fixed (pair,row) register tile that it accumulates over the cell's particles, and the
// finished tile is flushed with one fp64 RED per value to the unique-corner arrays.
#include "amps_dev.cuh"

namespace amps {

__constant__ int cIndexMatrix[8][8] = {{0, 2, 8, 6, 18, 20, 26, 24},  {1, 0, 6, 7, 19, 18, 24, 25},
                                       {4, 3, 0, 1, 22, 21, 18, 19},  {3, 5, 2, 0, 21, 23, 20, 18},
                                       {9, 11, 17, 15, 0, 2, 8, 6},   {10, 9, 15, 16, 1, 0, 6, 7},
                                       {13, 12, 9, 10, 4, 3, 0, 1},   {12, 14, 11, 9, 3, 5, 2, 0}};
// cell-corner order (0,0,0)(1,0,0)(1,1,0)(0,1,0)(0,0,1)(1,0,1)(1,1,1)(0,1,1)
__constant__ int cCornerOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};

constexpr int DEP_THREADS = 128;
constexpr int DEP_CHUNK = 128;  // particles staged per pass

__device__ __forceinline__ void atomicMaxPositiveDouble(unsigned long long *addr, double v) {
  // v >= 0 and not NaN: the bit patterns of non-negative doubles order like unsigned integers
  atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

__global__ void __launch_bounds__(DEP_THREADS) deposit_kernel(DevMesh m, DevSpecies sp, ParticleSoA p, const int *__restrict__ cellStart,
                                                             const double *__restrict__ bCurTile, double *__restrict__ J, double *__restrict__ M,
                                                             double *__restrict__ energyOut, unsigned long long *__restrict__ cflBits) {
  __shared__ double sW[8][DEP_CHUNK];
  __shared__ double sA[9][DEP_CHUNK];
  __shared__ double sVr[3][DEP_CHUNK];  // q~ * (alpha v)
  __shared__ double sK[DEP_CHUNK];      // q~ * beta / V
  __shared__ double sRed[DEP_THREADS / 32][1 + AMPS_GPU_MAX_SPECIES];
  __shared__ int sCnt[DEP_THREADS / 32][AMPS_GPU_MAX_SPECIES];

  const int t = threadIdx.x;
  const int C = m.cellsPerBlock;
  const long long nCells = (long long)m.nLeaves * C;

  // fixed role of this thread in phase 2
  int ic = 0, jc = 0, row = 0;
  const bool mmThread = t < 108, jThread = (t >= 108 && t < 116);
  if (mmThread) {
    const int pair = t / 3;
    row = t - 3 * pair;
    // pair -> (ic,jc), jc<=ic : ic(ic+1)/2 + jc
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= pair) i++;
    ic = i, jc = pair - i * (i + 1) / 2;
  } else if (jThread) {
    ic = t - 108;
  }

  double energyThread = 0.0;
  double cflThread[AMPS_GPU_MAX_SPECIES];
#pragma unroll
  for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++) cflThread[s] = 0.0;

  for (long long cell = blockIdx.x; cell < nCells; cell += gridDim.x) {
    const int begin = cellStart[cell], end = cellStart[cell + 1];
    if (begin == end) continue;  // ProcessCell returns false: nothing is flushed
    const int leaf = (int)(cell / C);
    const int cin = (int)(cell - (long long)leaf * C);
    const LeafGeo &lg = m.leaf[leaf];
    // periodic "ghost" (boundary) blocks are skipped, :3815-3825
    if (m.periodic && lg.face != 0) continue;

    const int kc = cin / (m.N[0] * m.N[1]);
    const int jc_ = (cin - kc * m.N[0] * m.N[1]) / m.N[0];
    const int ic_ = cin - kc * m.N[0] * m.N[1] - jc_ * m.N[0];

    double dx[3], dxc[3], span[3];
    double CellVolume = 1;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      dx[d] = (lg.xmax[d] - lg.xmin[d]) / m.N[d] * sp.length_conv;
      dxc[d] = (lg.xmax[d] - lg.xmin[d]) / m.N[d];
      span[d] = lg.xmax[d] - lg.xmin[d];
    }
#pragma unroll
    for (int d = 0; d < 3; d++) CellVolume *= dx[d];
    const double *bT = bCurTile + (size_t)leaf * m.bTileStride;
    const int BS0 = m.TN[0], BS1 = m.TN[0] * m.TN[1];

    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    double eCell = 0.0;
    double vmean[AMPS_GPU_MAX_SPECIES];
    int cnt[AMPS_GPU_MAX_SPECIES];
#pragma unroll
    for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++) vmean[s] = 0.0, cnt[s] = 0;

    for (int base = begin; base < end; base += DEP_CHUNK) {
      const int np = min(DEP_CHUNK, end - base);
      __syncthreads();  // previous chunk fully consumed
      // ---------------- phase 1: per-particle factors ----------------
      if (t < np) {
        const int ip = base + t;
        double x[3] = {p.x[0][ip], p.x[1][ip], p.x[2][ip]};
        double v[3] = {p.v[0][ip], p.v[1][ip], p.v[2][ip]};
        const int spec = p.spec[ip];
        const double LocalParticleWeight = sp.weight[spec] * p.w[ip];

        // B at the particle: cell-centred linear stencil on B_cur (:2100-2129)
        double B[3] = {0.0, 0.0, 0.0};
        {
          const double iLoc = (x[0] - lg.xmin[0]) / span[0] * m.N[0];
          const double jLoc = (x[1] - lg.xmin[1]) / span[1] * m.N[1];
          const double kLoc = (x[2] - lg.xmin[2]) / span[2] * m.N[2];
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
          double norm = 0.0;
#pragma unroll
          for (int s = 0; s < 8; s++)
            if (valid & (1u << s)) norm += w[s];
          const double inv = (norm > 0.0) ? 1.0 / norm : 1.0;
          const int nd0 = centerLocalNumber(m, i0, j0, k0);
#pragma unroll
          for (int s = 0; s < 8; s++) {
            if (valid & (1u << s)) {
              const int nd = nd0 + ((s >> 2) & 1) + ((s >> 1) & 1) * BS0 + (s & 1) * BS1;
              const double ws = w[s] * inv;
              B[0] += ws * __ldg(bT + 3 * nd);
              B[1] += ws * __ldg(bT + 3 * nd + 1);
              B[2] += ws * __ldg(bT + 3 * nd + 2);
            }
          }
        }
#pragma unroll
        for (int d = 0; d < 3; d++) {
          B[d] *= sp.B_conv;
          v[d] *= sp.length_conv;
        }
        const double chargeQ = sp.charge[spec] * LocalParticleWeight;
        const double mass = sp.mass[spec] * LocalParticleWeight;
        const double QdT_over_m = chargeQ * sp.dtTotal / mass;
        const double QdT_over_2m = 0.5 * QdT_over_m;
        const double QdT_over_2m_squared = QdT_over_2m * QdT_over_2m;
#pragma unroll
        for (int d = 0; d < 3; d++) B[d] /= sp.LightSpeed;

        double P[3], BB[3][3];
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
        double alpha[9];
        alpha[0] = c0 * (1.0 + BB[0][0]);
        alpha[1] = c0 * (-P[2] + BB[0][1]);
        alpha[2] = c0 * (P[1] + BB[0][2]);
        alpha[3] = c0 * (P[2] + BB[1][0]);
        alpha[4] = c0 * (1.0 + BB[1][1]);
        alpha[5] = c0 * (-P[0] + BB[1][2]);
        alpha[6] = c0 * (-P[1] + BB[2][0]);
        alpha[7] = c0 * (P[0] + BB[2][1]);
        alpha[8] = c0 * (1.0 + BB[2][2]);

        // un-normalised corner weights WeightPG (CornerBased::InitStencil with the table argument, :2200)
        double xl[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          double xx = x[d];
          if (fabs(xx - lg.xmax[d]) < 1e-10 * dxc[d]) xx = lg.xmax[d] - 1e-10 * dxc[d];
          double q = (xx - lg.xmin[d]) / dxc[d];
          q -= (int)q;
          xl[d] = q;
        }
        const double ax0 = 1.0 - xl[0], ax1 = xl[0], ay0 = 1.0 - xl[1], ay1 = xl[1], az0 = 1.0 - xl[2], az1 = xl[2];
        sW[0][t] = ax0 * ay0 * az0;
        sW[1][t] = ax1 * ay0 * az0;
        sW[2][t] = ax1 * ay1 * az0;
        sW[3][t] = ax0 * ay1 * az0;
        sW[4][t] = ax0 * ay0 * az1;
        sW[5][t] = ax1 * ay0 * az1;
        sW[6][t] = ax1 * ay1 * az1;
        sW[7][t] = ax0 * ay1 * az1;

        const double vsqr = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
        vmean[spec] += sqrt(vsqr) * sp.dt[0];
        cnt[spec]++;
        eCell += 0.5 * mass * vsqr;

#pragma unroll
        for (int q = 0; q < 9; q++) sA[q][t] = alpha[q];
#pragma unroll
        for (int d = 0; d < 3; d++) sVr[d][t] = chargeQ * (alpha[3 * d] * v[0] + alpha[3 * d + 1] * v[1] + alpha[3 * d + 2] * v[2]);
        sK[t] = chargeQ * QdT_over_2m / CellVolume;
      }
      __syncthreads();
      // ---------------- phase 2: register tile accumulation ----------------
      if (mmThread) {
#pragma unroll 4
        for (int q = 0; q < np; q++) {
          const double u = sW[jc][q] * (sK[q] * sW[ic][q]);
          acc0 += sA[3 * row][q] * u;
          acc1 += sA[3 * row + 1][q] * u;
          acc2 += sA[3 * row + 2][q] * u;
        }
      } else if (jThread) {
#pragma unroll 4
        for (int q = 0; q < np; q++) {
          const double wq = sW[ic][q];
          acc0 += wq * sVr[0][q];
          acc1 += wq * sVr[1][q];
          acc2 += wq * sVr[2][q];
        }
      }
    }

    // ---------------- flush ----------------
    const int *uidT = m.cornerUid + (size_t)leaf * m.nCornerLocal;
    if (mmThread) {
      const int ui = uidT[cornerLocalNumber(m, ic_ + cCornerOff[ic][0], jc_ + cCornerOff[ic][1], kc + cCornerOff[ic][2])];
      double *Mi = M + (size_t)ui * 243 + 9 * cIndexMatrix[ic][jc] + 3 * row;
      atomicAdd(Mi, acc0);
      atomicAdd(Mi + 1, acc1);
      atomicAdd(Mi + 2, acc2);
      if (ic != jc) {
        const int uj = uidT[cornerLocalNumber(m, ic_ + cCornerOff[jc][0], jc_ + cCornerOff[jc][1], kc + cCornerOff[jc][2])];
        double *Mj = M + (size_t)uj * 243 + 9 * cIndexMatrix[jc][ic] + 3 * row;
        atomicAdd(Mj, acc0);
        atomicAdd(Mj + 1, acc1);
        atomicAdd(Mj + 2, acc2);
      }
    } else if (jThread) {
      const int ui = uidT[cornerLocalNumber(m, ic_ + cCornerOff[ic][0], jc_ + cCornerOff[ic][1], kc + cCornerOff[ic][2])];
      double *Ji = J + (size_t)ui * 3;
      atomicAdd(Ji, acc0 / CellVolume);
      atomicAdd(Ji + 1, acc1 / CellVolume);
      atomicAdd(Ji + 2, acc2 / CellVolume);
    }

    // ---------------- per-cell diagnostics: energy (x8, reference quirk :3860) and cfl ----------------
    {
      double e = eCell;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
      if ((t & 31) == 0) sRed[t >> 5][0] = e;
      for (int s = 0; s < sp.n; s++) {
        double vm = vmean[s];
        int c = cnt[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          vm += __shfl_xor_sync(0xffffffffu, vm, o);
          c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if ((t & 31) == 0) sRed[t >> 5][1 + s] = vm, sCnt[t >> 5][s] = c;
      }
      __syncthreads();
      if (t == 0) {
        double es = 0.0;
        for (int w = 0; w < DEP_THREADS / 32; w++) es += sRed[w][0];
        energyThread += 8.0 * es;
        const double diag = sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
        for (int s = 0; s < sp.n; s++) {
          double vm = 0.0;
          int c = 0;
          for (int w = 0; w < DEP_THREADS / 32; w++) vm += sRed[w][1 + s], c += sCnt[w][s];
          if (c > 0) {  // 0/0 = NaN never wins the reference's '>' comparison
            const double cfl = vm / (c * diag);
            if (cfl > cflThread[s]) cflThread[s] = cfl;
          }
        }
      }
    }
  }

  if (t == 0) {
    if (energyThread != 0.0) atomicAdd(energyOut, energyThread);
    for (int s = 0; s < sp.n; s++)
      if (cflThread[s] > 0.0) atomicMaxPositiveDouble(&cflBits[s], cflThread[s]);
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
  long long grid = 148LL * 8;
  if (grid > nCells) grid = nCells;
  deposit_kernel<<<(int)grid, DEP_THREADS, 0, s>>>(m, sp, p, cellStart, bCurTile, J, M, energy, cflBits);
  (*launches)++;
}

}  // namespace amps
