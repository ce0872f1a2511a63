// deposit_gc.cu -- the guiding-centre species of PIC::GYROKINETIC in ECSIM::ProcessCell (sm_100a, fp64).
//
//   use_gc_species                src/pic/pic_field_solver_ecsim.cpp:2084-2085  (IsGuidingCenterSpecies, pic.h:5052)
//   mu b to the corners           :2205-2225      |v_normal|^2 in the energy / cfl diagnostics  :2228-2238
//   explicit current q v_eff      :2253-2257      no mass matrix                                :2310
//   J += curl(M) per cell         :1828-1875 (ECSIM_AddGuidingCenterMagnetizationCurrentToCorners), called :2376
//
// The full-orbit species stay with deposit_kernel (deposit.cu), which is launched with the charge of the guiding-centre species set
// to zero (they add nothing to J and M there) and without its fused diagnostics.  This kernel then walks the sorted store once more,
// ONE WARP PER CELL: the energy / cfl diagnostics of every species (v_normal included), and for the guiding-centre particles the
// corner weights, B at the particle, the explicit current and the magnetisation sums; the 24 + 24 cell sums are reduced over the
// warp and the cell's current (Jg / V + the closure) is added to J with 24 REDs.  v_normal is read by ParticleBuffer slot
// (amps_gpu_v_normal_upload): the device never changes it, so it does not travel with the sorted copies.
#include "amps_dev.cuh"

namespace amps {

__device__ __forceinline__ int gcox(int c) { return ((c + 1) >> 1) & 1; }  // cell-corner order (0,0,0)(1,0,0)(1,1,0)(0,1,0)(0,0,1)...
__device__ __forceinline__ int gcoy(int c) { return (c >> 1) & 1; }
__device__ __forceinline__ int gcoz(int c) { return (c >> 2) & 1; }

__global__ void __launch_bounds__(256) gc_deposit_kernel(DevMesh m, DevSpecies sp, unsigned gcMask, ParticleSoA p, const int *__restrict__ cellStart,
                                                        const double *__restrict__ bCurTile, const double *__restrict__ vnByPtr, long long nVn,
                                                        double *__restrict__ J, double *__restrict__ energyOut,
                                                        unsigned long long *__restrict__ cflBits) {
  const int lane = threadIdx.x & 31;
  const int warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
  const int C = m.cellsPerBlock;
  const bool cornerB = sp.bMode == AMPS_B_CORNER_BASED;
  double eAcc = 0.0;
  double cflMax = 0.0;  // lane s keeps species s
  for (long long idx = warpGlobal; idx < (long long)m.nDepReal * C; idx += nWarps) {
    const int leaf = m.depLeaf[idx / C], cin = (int)(idx % C), cell = leaf * C + cin;
    const int begin = cellStart[cell], end = cellStart[cell + 1];
    if (begin == end) continue;
    const LeafGeo &lg = m.leaf[leaf];
    const int kc = cin / (m.N[0] * m.N[1]), jc = (cin - kc * m.N[0] * m.N[1]) / m.N[0], ic = cin - kc * m.N[0] * m.N[1] - jc * m.N[0];
    const double *bT = bCurTile + (size_t)leaf * m.bTileStride;
    double Jg[24], Mc[24];
#pragma unroll
    for (int q = 0; q < 24; q++) Jg[q] = 0.0, Mc[q] = 0.0;
    double vm[AMPS_GPU_MAX_SPECIES];
    int cnt[AMPS_GPU_MAX_SPECIES];
#pragma unroll
    for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++) vm[s] = 0.0, cnt[s] = 0;
    bool hasGc = false;
    for (int ip = begin + lane; ip < end; ip += 32) {
      const int spec = p.spec[ip] & 0x3f;
      const bool gc = (gcMask >> spec) & 1u;
      const double LocalParticleWeight = sp.weight[spec] * p.w[ip];
      const double v0 = p.v[0][ip] * sp.length_conv, v1 = p.v[1][ip] * sp.length_conv, v2 = p.v[2][ip] * sp.length_conv;
      double vsqr = v0 * v0 + v1 * v1 + v2 * v2;
      if (gc) {
        const long long pt = p.ptr[ip];
        const double vperp = (vnByPtr != nullptr && pt >= 0 && pt < nVn) ? vnByPtr[pt] : 0.0;
        vsqr += vperp * vperp;
      }
      eAcc += 0.5 * (sp.mass[spec] * LocalParticleWeight) * vsqr;
      const double vabs = sqrt(vsqr) * sp.dt[0];
#pragma unroll
      for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++)
        if (s == spec) vm[s] += vabs, cnt[s]++;
      if (!gc) continue;
      // CornerBased::InitStencil (pic_interpolation_routines.cpp:1090-1098): local coordinates and the raw corner weights WeightPG
      double xl[3];
      {
        const double xx[3] = {p.x[0][ip], p.x[1][ip], p.x[2][ip]};
#pragma unroll
        for (int d = 0; d < 3; d++) {
          double xs = xx[d];
          if (fabs(xs - lg.xmax[d]) < 1e-10 * lg.dxc[d]) xs = lg.xmax[d] - 1e-10 * lg.dxc[d];
          double r = (xs - lg.xmin[d]) / lg.dxc[d];
          r -= (int)r;
          xl[d] = r;
        }
      }
      const double X[2] = {1.0 - xl[0], xl[0]}, Y[2] = {1.0 - xl[1], xl[1]}, Z[2] = {1.0 - xl[2], xl[2]};
      double W[8];
#pragma unroll
      for (int c = 0; c < 8; c++) W[c] = X[gcox(c)] * Y[gcoy(c)] * Z[gcoz(c)];
      // B at the particle on B_cur (:2100-2129), then B *= B_conv (:2134); the closure uses it before the /LightSpeed scaling
      double B0 = 0.0, B1 = 0.0, B2 = 0.0;
      if (cornerB) {
#pragma unroll
        for (int s = 0; s < 8; s++) {
          const int di = (s >> 2) & 1, dj = (s >> 1) & 1, dk = s & 1;
          const double w = X[di] * Y[dj] * Z[dk];
          const double *t = bT + 3 * cornerLocalNumber(m, ic + di, jc + dj, kc + dk);
          B0 += w * t[0], B1 += w * t[1], B2 += w * t[2];
        }
      } else {
        int o[3];
        double w[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          o[d] = (xl[d] < 0.5) ? 0 : 1;
          w[d] = xl[d] + 0.5 - (double)o[d];
        }
        const int i0 = ic - 1 + o[0], j0 = jc - 1 + o[1], k0 = kc - 1 + o[2];
        unsigned valid = 0xffu;
        if (!m.periodic && lg.face) {  // AddCell drops centres outside the global box (pic.h:7235-7245), the rest is re-normalised
          if ((lg.face & 1) && i0 < 0) valid &= 0xf0u;
          if ((lg.face & 2) && i0 + 1 >= m.N[0]) valid &= 0x0fu;
          if ((lg.face & 4) && j0 < 0) valid &= 0xccu;
          if ((lg.face & 8) && j0 + 1 >= m.N[1]) valid &= 0x33u;
          if ((lg.face & 16) && k0 < 0) valid &= 0xaau;
          if ((lg.face & 32) && k0 + 1 >= m.N[2]) valid &= 0x55u;
        }
        double norm = 0.0;
#pragma unroll
        for (int s = 0; s < 8; s++) {
          const int di = (s >> 2) & 1, dj = (s >> 1) & 1, dk = s & 1;
          if (!(valid & (1u << s))) continue;
          const double ws = (di ? w[0] : 1.0 - w[0]) * (dj ? w[1] : 1.0 - w[1]) * (dk ? w[2] : 1.0 - w[2]);
          const double *t = bT + 3 * centerLocalNumber(m, i0 + di, j0 + dj, k0 + dk);
          B0 += ws * t[0], B1 += ws * t[1], B2 += ws * t[2];
          norm += ws;
        }
        if (norm > 0.0) B0 /= norm, B1 /= norm, B2 /= norm;
      }
      B0 *= sp.B_conv, B1 *= sp.B_conv, B2 *= sp.B_conv;
      const double chargeQ = sp.charge[spec] * LocalParticleWeight;
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const double t = chargeQ * W[c];
        Jg[3 * c] += t * v0, Jg[3 * c + 1] += t * v1, Jg[3 * c + 2] += t * v2;
      }
      const double absB = sqrt(B0 * B0 + B1 * B1 + B2 * B2);
      if (absB > 0.0 && p.mu != nullptr) {
        const double mu_tot = p.mu[ip] * LocalParticleWeight;
        const double b0 = B0 / absB, b1 = B1 / absB, b2 = B2 / absB;
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const double w = mu_tot * W[c];
          Mc[3 * c] += w * b0, Mc[3 * c + 1] += w * b1, Mc[3 * c + 2] += w * b2;
        }
      }
      hasGc = true;
    }
    // cfl of the cell per species (:2355-2359); 0/0 = NaN never wins the reference's '>' comparison
#pragma unroll
    for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++) {
      double a = vm[s];
      int c = cnt[s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o), c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == s && c > 0) {
        const double cfl = a / (c * lg.diag);
        if (cfl > cflMax) cflMax = cfl;
      }
    }
    if (!__any_sync(0xffffffffu, hasGc)) continue;
#pragma unroll
    for (int q = 0; q < 24; q++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) Jg[q] += __shfl_xor_sync(0xffffffffu, Jg[q], o), Mc[q] += __shfl_xor_sync(0xffffffffu, Mc[q], o);
    }
    if (lane < 24) {
      const int c = lane / 3, d = lane - 3 * c;
      const double CellVolume = 1.0 / lg.invV;
      // curl of the trilinear reconstruction of M = MCornerSum / V, one-sided at corner c (:1845-1873)
      const int uc = gcox(c), vc = gcoy(c), wc = gcoz(c);
      double Jc[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int ua = gcox(a), va = gcoy(a), wa = gcoz(a);
        double dNdx = 0.0, dNdy = 0.0, dNdz = 0.0;
        if (va == vc && wa == wc) dNdx = (ua ? 1.0 : -1.0) / (lg.dxc[0] * sp.length_conv);
        if (ua == uc && wa == wc) dNdy = (va ? 1.0 : -1.0) / (lg.dxc[1] * sp.length_conv);
        if (ua == uc && va == vc) dNdz = (wa ? 1.0 : -1.0) / (lg.dxc[2] * sp.length_conv);
        const double M0 = Mc[3 * a] / CellVolume, M1 = Mc[3 * a + 1] / CellVolume, M2 = Mc[3 * a + 2] / CellVolume;
        Jc[0] += dNdy * M2 - dNdz * M1;
        Jc[1] += dNdz * M0 - dNdx * M2;
        Jc[2] += dNdx * M1 - dNdy * M0;
      }
      double jg = 0.0, jc3 = 0.0;
#pragma unroll
      for (int q = 0; q < 24; q++)
        if (q == lane) jg = Jg[q];
#pragma unroll
      for (int q = 0; q < 3; q++)
        if (q == d) jc3 = Jc[q];
      const int uid = m.cornerUid[(size_t)leaf * m.nCornerLocal + cornerLocalNumber(m, ic + uc, jc + vc, kc + wc)];
      atomicAdd(J + (size_t)uid * 3 + d, jg / CellVolume + jc3);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) eAcc += __shfl_xor_sync(0xffffffffu, eAcc, o);
  if (lane == 0 && eAcc != 0.0) atomicAdd(energyOut, 8.0 * eAcc);  // added once per corner in the reference's flush loops (:3860)
  if (lane < sp.n && cflMax > 0.0) atomicMax(&cflBits[lane], (unsigned long long)__double_as_longlong(cflMax));
}

void launch_gc_deposit(const DevMesh &m, const DevSpecies &sp, unsigned gcMask, ParticleSoA p, const int *cellStart, const double *bCurTile,
                       const double *vnByPtr, long long nVn, double *J, double *energy, unsigned long long *cflBits, int nSM, cudaStream_t s) {
  gc_deposit_kernel<<<nSM * 2, 256, 0, s>>>(m, sp, gcMask, p, cellStart, bCurTile, vnByPtr, nVn, J, energy, cflBits);
}

}  // namespace amps
