// deposit.cu -- ECSIM current + mass-matrix deposition (sm_100a, fp64, FMA contraction allowed:
// results are compared to the reference within 1e-10 relative, summation order differs anyway).
//
//   a10 ECSIM::ProcessCell          src/pic/pic_field_solver_ecsim.cpp:1881-2438
//   a11 ECSIM::UpdateJMassMatrix    src/pic/pic_field_solver_ecsim.cpp:3244-3995
//   a12 ProcessJMassMatrix          src/pic/pic_field_solver_ecsim.cpp:1383-1395 (implicit: all copies of a
//                                   shared / periodic corner are ONE unique corner on the device)
//
// Algebra.  The reference accumulates, per cell, MM[c][c'] += (W_c W_c') k alpha (36 corner pairs x 9)
// and J[c] += W_c q~ alpha v / V (8 x 3).  The trilinear weights factorise, W_c = X_cx Y_cy Z_cz, so the
// product W_c W_c' only depends on the per-dimension sums (cx+c'x, cy+c'y, cz+c'z) in {0,1,2}^3: there
// are 27 distinct "classes" u_cls = XX[px] YY[py] ZZ[pz] (XX = {X0 X0, X0 X1, X1 X1}).  With
//        T[cls][0..8]  = sum_p u_cls (k alpha)        T[cls][9..11] = sum_p u_cls (q~ alpha v / V)
// the mass matrix is MM[c][c'] = T[cls(c,c')][0..8] and, because sum_c' W_c' = 1,
// J[c] = sum_c' T[cls(c,c')][9..11].  That is 324 accumulators per cell and 324 DFMA per particle.
//
// T = U^T A is a dense contraction over the particles of the cell (U[p][27], A[p][12]), so phase 2 runs on the fp64 MMA path
// (mma.sync.m8n8k4.f64 -> SASS DMMA.884): measured on the B200 (tools/fp64_peak.cu) DMMA and DFMA share ONE pipe (33.6 / 37.2
// TFLOP/s alone, 35 mixed), so the MMA form buys no arithmetic, but a DMMA carries 256 FMAs per issue slot, needs one 8-byte
// shared-memory operand per lane per 256 FMAs and keeps a 8x8 tile of T in TWO registers per lane: the 108 accumulator registers
// of the DFMA register-tile version (242 registers, 8 warps/SM, whose latency-bound phase 1 could not overlap) become 32.
//
// The kernel is bound by the fp64 pipe, not by HBM (57 B/particle in, ~4 KB/cell out).  Mapping: ONE WARP PER CELL, no
// block-level synchronisation at all; the resident warps drift apart, so the latency-bound phase of one warp overlaps the
// DMMA-bound phase of the others.  Per 32-particle chunk of its cell a warp runs
//   phase 1  lane <-> particle: B gather (27 centres of the cell staged in shared memory), alpha, the 27 class weights and
//            the 12 columns -> 39 doubles per particle in the warp's shared-memory slab, laid out [group of 4 particles][element]
//            [particle in group] so that an MMA operand load of the warp is 32 consecutive doubles (conflict-free)
//   phase 2  per group of 4 particles (the K of m8n8k4): 4 A operands (classes 8t+g), 2 B operands (columns 8n+g), 8 DMMA into
//            the 4 x 2 tiles that cover T padded to 32 x 16 (rows >= 27 and columns >= 12 are never read)
// and per cell
//   spill    the C fragments go to T[32][16] in the slab (the K sum is complete inside the MMA: nothing to fold)
//   flush    one fp64 RED per value into J[nCorners][3], M[nCorners][243] (576 + 24 per cell).
// The particles of the next chunk - of this cell or, in its last chunk, of the warp's next cell - are requested one chunk ahead,
// and the "cell header" (the B_cur values around the cell + the leaf geometry, 96 doubles) of the warp's next cell one cell ahead
// with cp.async into a double-buffered per-warp slot.  With two resident warps per scheduler every exposed latency of the per-cell
// and per-chunk serial sections is paid in full, so the cell walk uses no integer division and no dependent global load.
//
// The energy / cfl diagnostics of UpdateJMassMatrix (:2228-2238, :2355-2359, :3860-3864) are folded into phase 1 for at most
// two species (kDiag), a separate streaming kernel (diag_kernel) otherwise.
#include <cstddef>
#include "amps_dev.cuh"

namespace amps {

constexpr int N_EL = 39;                   // doubles per particle: u[27] (class weights) - a[12] (k alpha[9], q~ alpha v/V [3])
constexpr int EL_A = 27;
constexpr int GRP = 4 * N_EL;              // doubles per group of 4 particles, [element][particle in group]; 2*GRP = 24 mod 32 words:
                                           // the 39 phase-1 stores of a half-warp (4 groups x 4 particles) hit 32 distinct banks
constexpr int CHUNK = 32;                  // particles per phase-1 pass (one per lane)
#ifndef AMPS_DEP_WARPS
#define AMPS_DEP_WARPS 4
#endif
#ifndef AMPS_DEP_CTAS
#define AMPS_DEP_CTAS 3
#endif
constexpr int DEP_WARPS = AMPS_DEP_WARPS, DEP_THREADS = 32 * DEP_WARPS, DEP_CTAS_PER_SM = AMPS_DEP_CTAS;
constexpr int T_LD = 24;                   // leading dimension of T[32][16] after the spill: 2 T_LD = 16 mod 32 words, so the 16-byte
                                           // fragment stores of a quarter-warp (rows g, g+1) fall on disjoint banks
constexpr int T_J = 32 * T_LD;             // the three current columns once more, compact: Jt[27][3] (odd stride: the eight
                                           // classes a lane of the J flush sums lie on different banks)
constexpr int SLAB = (CHUNK / 4) * GRP;    // doubles per warp: the particle elements of a chunk, later T and Jt
static_assert(SLAB >= T_J + 27 * 3, "T and Jt must fit in the slab");
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// per-warp "cell header" in shared memory, double-buffered: the 27 x 3 centre (or 8 x 3 corner) values of B_cur the cell's
// stencils can touch + the leaf geometry phase 1 needs.  The header of the warp's NEXT cell is fetched with 8-byte cp.async
// (SASS LDGSTS: no registers are held while the copies are in flight) during the accumulation of the current one.
static_assert(offsetof(LeafGeo, xmin) == 0 && offsetof(LeafGeo, xmax) == 24 && offsetof(LeafGeo, dxc) == 88 && offsetof(LeafGeo, invdxc) == 112 &&
                  offsetof(LeafGeo, invV) == 136 && offsetof(LeafGeo, diag) == 144,
              "stage_header copies LeafGeo by double index");
constexpr int DEC_MAX = 1024;  // largest block (cells) whose (i,j,k) decode table is kept in shared memory; larger blocks divide
constexpr int HDR_GEO = 82, HDR = 96;  // doubles: B [0,81), xmin[3] xmax[3] dxc[3] invdxc[3] invV diag at [82,96)
__device__ __forceinline__ void cp_async8(double *dstShared, const double *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dstShared)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void atomicMaxPositiveDouble(unsigned long long *addr, double v) {
  // v >= 0 and not NaN: the bit patterns of non-negative doubles order like unsigned integers
  atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}
// 1/x for x >= 1 to ~2 ulp: float seed + two Newton steps (the IEEE division costs ~4x more)
__device__ __forceinline__ double fast_rcp(double x) {
  double y = (double)__frcp_rn((float)x);
  y = y * fma(-x, y, 2.0);
  y = y * fma(-x, y, 2.0);
  return y;
}
// corner offsets of the cell-corner order (0,0,0)(1,0,0)(1,1,0)(0,1,0)(0,0,1)(1,0,1)(1,1,1)(0,1,1) as arithmetic
__device__ __forceinline__ int cox(int c) { return ((c + 1) >> 1) & 1; }
__device__ __forceinline__ int coy(int c) { return (c >> 1) & 1; }
__device__ __forceinline__ int coz(int c) { return (c >> 2) & 1; }
// class of the ordered corner pair (c,c'): per-dimension sums of the corner offsets
__device__ __forceinline__ int pair_class(int c, int d) { return (cox(c) + cox(d)) * 9 + (coy(c) + coy(d)) * 3 + (coz(c) + coz(d)); }
// IndexMatrix[c][c'] (:1377-1380): neighbour slot ii+3jj+9kk, per-dimension offset 0,-1,+1 -> 0,1,2
__device__ __forceinline__ int nb_slot(int delta) { return (3 * delta * delta + delta) >> 1; }
__device__ __forceinline__ int index_matrix(int c, int d) {
  return nb_slot(cox(d) - cox(c)) + 3 * nb_slot(coy(d) - coy(c)) + 9 * nb_slot(coz(d) - coz(c));
}

// kDiag: the energy / cfl diagnostics (diag_kernel below) are folded into phase 1 when there are at most two species
// kGather: the counting sort only produced the permutation; sorted particle i is p[perm[i]] and this kernel writes it to dst[i]
//          (the physical reordering costs no extra pass: its reads are the deposit's own reads)
template <bool kCornerB, bool kDiag, bool kGather>
__global__ void __launch_bounds__(DEP_THREADS, DEP_CTAS_PER_SM) deposit_kernel(DevMesh m, DevSpecies sp, ParticleSoA p,
                                                                              const int *__restrict__ cellStart,
                                                                              const double *__restrict__ bCurTile, double *__restrict__ J,
                                                                              double *__restrict__ M, double *__restrict__ energyOut,
                                                                              unsigned long long *__restrict__ cflBits,
                                                                              const int *__restrict__ perm, ParticleSoA dst, int dep0, int dep1,
                                                                              int ghostPass) {
  extern __shared__ __align__(16) double sRows[];  // [DEP_WARPS][SLAB]
  __shared__ double sHdr[DEP_WARPS][2][HDR];  // cell headers (B_cur around the cell + leaf geometry), current and next cell
  __shared__ unsigned sDec[DEC_MAX];          // cell number inside a block -> i | j << 10 | k << 20 (blocks of at most DEC_MAX cells)
  __shared__ int sBoff[81];                   // centre-B mode: offset of stencil entry e = 3 n + d from the cell's own centre
  // flush tables: output o = (c*8+c')*9+col -> corner c (3 bits) | offset inside M[corner] (8 bits) | T index (9 bits)
  __shared__ unsigned int sFlush[576];
  __shared__ unsigned short sJcls[64];  // T index (without column) of pair (c,c')
  __shared__ double sQdt2m[AMPS_GPU_MAX_SPECIES], sInvBeta[AMPS_GPU_MAX_SPECIES];
  if (threadIdx.x < AMPS_GPU_MAX_SPECIES) {
    const double b = (threadIdx.x < sp.n) ? 0.5 * (sp.charge[threadIdx.x] * sp.dtTotal / sp.mass[threadIdx.x]) : 0.0;
    sQdt2m[threadIdx.x] = b;
    sInvBeta[threadIdx.x] = (b != 0.0) ? 1.0 / b : 0.0;  // neutral species deposit nothing
  }
  if (m.cellsPerBlock <= DEC_MAX)
    for (int o = threadIdx.x; o < m.cellsPerBlock; o += DEP_THREADS) {
      const int k = o / (m.N[0] * m.N[1]), j = (o - k * m.N[0] * m.N[1]) / m.N[0], i = o - k * m.N[0] * m.N[1] - j * m.N[0];
      sDec[o] = (unsigned)i | ((unsigned)j << 10) | ((unsigned)k << 20);
    }
  for (int o = threadIdx.x; o < 576; o += DEP_THREADS) {
    const int c = o / 72, r = o - 72 * c, d = r / 9, col = r - 9 * d;
    sFlush[o] = (unsigned)c | ((unsigned)(9 * index_matrix(c, d) + col) << 3) | ((unsigned)(pair_class(c, d) * T_LD + col) << 11);
    if (o < 64) sJcls[o] = (unsigned short)(T_J + pair_class(o >> 3, o & 7) * 3);
    if (o < 81) {
      const int n = o / 3, d = o - 3 * n;
      const int di = n % 3 - 1, dj = (n / 3) % 3 - 1, dk = n / 9 - 1;
      sBoff[o] = 3 * (di + m.TN[0] * (dj + dk * m.TN[1])) + d;
    }
  }
  // the slab starts finite: the elements of lanes beyond the end of a cell's last chunk are multiplied by zero weights
  for (int o = threadIdx.x; o < DEP_WARPS * SLAB; o += DEP_THREADS) sRows[o] = 0.0;
  __syncthreads();

  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warpGlobal = blockIdx.x * DEP_WARPS + wib, nWarps = gridDim.x * DEP_WARPS;
  const int C = m.cellsPerBlock;
  // this launch deposits the cells of the leaves m.depLeaf[dep0 .. dep1).  The warps walk the cells of the depositing leaves
  // only (periodic "ghost" blocks are skipped, :3815-3825), so that a warp's next cell is known one cell ahead and its first
  // particles are requested while the current cell is still being accumulated.
  const int idx0 = dep0 * C, idx1 = dep1 * C;
  double *rows = sRows + (size_t)wib * SLAB;
  int hbuf = 0;         // header buffer of the current cell
  bool staged = false;  // the header of the warp's next cell is already in flight
  // cell number inside its block -> (i,j,k): a table lookup for the usual block sizes, no integer division per cell
  auto decode = [&](int ci, int &i, int &j, int &k) {
    if (C <= DEC_MAX) {
      const unsigned t = sDec[ci];
      i = t & 1023u, j = (t >> 10) & 1023u, k = t >> 20;
    } else {
      k = ci / (m.N[0] * m.N[1]);
      j = (ci - k * m.N[0] * m.N[1]) / m.N[0];
      i = ci - k * m.N[0] * m.N[1] - j * m.N[0];
    }
  };
  // fetch the header of cell ci of leaf lf into h (asynchronously; complete after cp_async_wait_all + __syncwarp)
  auto stage_header = [&](int lf, int ci, double *h) {
    int i, j, k;
    decode(ci, i, j, k);
    const double *bT = bCurTile + (size_t)lf * m.bTileStride;
    if (kCornerB) {
      // _PIC_FIELD_SOLVER_B_CORNER_BASED_: B_cur on the 8 corners of the cell, slot = 4*di + 2*dj + dk (:1932-1946)
      if (lane < 24) {
        const int n = lane / 3, d = lane - 3 * n;
        cp_async8(h + lane, bT + 3 * cornerLocalNumber(m, i + ((n >> 2) & 1), j + ((n >> 1) & 1), k + (n & 1)) + d);
      }
    } else {
      const double *b0 = bT + 3 * centerLocalNumber(m, i, j, k);
#pragma unroll
      for (int e = lane; e < 81; e += 32) cp_async8(h + e, b0 + sBoff[e]);
    }
    // xmin[3] xmax[3] are doubles 0..5 of LeafGeo, dxc[3] invdxc[3] invV diag doubles 11..18
    if (lane < 14) cp_async8(h + HDR_GEO + lane, reinterpret_cast<const double *>(m.leaf + lf) + (lane < 6 ? lane : lane + 5));
  };

  // m8n8k4 fragments: A[row g][k kk] = u[particle 4 ks + kk][class 8 t + g], B[k kk][col g] = a[particle][column 8 n + g],
  // C[row g][cols 2 kk, 2 kk + 1].  Rows >= 27 / columns >= 12 of the padded T read a clamped element: their sums are never used
  const int g = lane >> 2, kk = lane & 3;
  const int eA0 = 4 * g + kk, eA1 = 4 * (8 + g) + kk, eA2 = 4 * (16 + g) + kk, eA3 = 4 * min(24 + g, 26) + kk;
  const int eB0 = 4 * (EL_A + g) + kk, eB1 = 4 * (EL_A + min(8 + g, 11)) + kk;
  double *const el = rows + (lane >> 2) * GRP + (lane & 3);  // phase 1: element e of this lane's particle is el[4 e]
  const double invc = 1.0 / sp.LightSpeed;
  double eAcc = 0.0, cflMax = 0.0;  // kDiag: per-lane energy; lane 16 s keeps the cfl of species s

  // software prefetch: the particle of the NEXT chunk (of this cell, or the first chunk of the warp's next cell) is loaded
  // while phase 2 of the current one runs
  double nx0 = 0, nx1 = 0, nx2 = 0, nv0 = 0, nv1 = 0, nv2 = 0, nw = 0;
  int nspec = 0;
  int nsrc = 0, nsrc2 = 0;  // kGather: source slot of the prefetched particle / of the one a chunk later
  int nptr = 0;             // kGather: its ParticleBuffer slot (travels to the sorted copy)
  bool pre = false;         // the first chunk of the next cell is already in flight

  // the warp walks cells idx0 + warpGlobal + n nWarps of the deposit order; (rl, off) = (idx / C, idx % C) is advanced without
  // a division
  const int stepQ = nWarps / C, stepR = nWarps - stepQ * C;
  int idx = idx0 + warpGlobal;
  int rl = idx / C, off = idx - rl * C;
  int nleaf = 0, noff = 0, nBegin = 0, nEnd = 0;  // the warp's NEXT cell and its cell table entry (requested one cell ahead)
  if (idx < idx1) {
    nleaf = m.depLeaf[rl], noff = off;
    const int nc = nleaf * C + noff;
    nBegin = cellStart[nc], nEnd = cellStart[nc + 1];
  }
  while (idx < idx1) {
    const int leaf = nleaf, cin = noff, cell = leaf * C + cin, begin = nBegin, end = nEnd;
    idx += nWarps;
    rl += stepQ, off += stepR;
    if (off >= C) off -= C, rl++;
    if (idx < idx1) {
      nleaf = m.depLeaf[rl], noff = off;
      const int nc = nleaf * C + noff;
      nBegin = cellStart[nc], nEnd = cellStart[nc + 1];
    } else {
      nBegin = nEnd = 0;
    }
    if (begin == end) continue;  // ProcessCell returns false: nothing is flushed (never prefetched: pre is false)
    const LeafGeo &lg = m.leaf[leaf];
    const int face = lg.face;
    int ic, jc, kc;
    decode(cin, ic, jc, kc);

    if (!pre) {
      // first cell of the warp, or the previous one was empty: the first chunk's loads are issued before the B staging so
      // that the two latencies overlap
      if (begin + lane < end) {
        const int ip = kGather ? perm[begin + lane] : begin + lane;
        nsrc = ip;
        nx0 = p.x[0][ip], nx1 = p.x[1][ip], nx2 = p.x[2][ip];
        nv0 = p.v[0][ip], nv1 = p.v[1][ip], nv2 = p.v[2][ip];
        nw = p.w[ip], nspec = p.spec[ip];
        if (kGather) nptr = p.ptr[ip];
      }
      if (kGather && begin + CHUNK + lane < end) nsrc2 = perm[begin + CHUNK + lane];
    }
    pre = false;
    int uidLane = 0;
    if (lane < 8) uidLane = m.cornerUid[(size_t)leaf * m.nCornerLocal + cornerLocalNumber(m, ic + cox(lane), jc + coy(lane), kc + coz(lane))];
    // the cell header: already in flight (requested during the previous cell) or fetched now; then request the next one
    if (!staged) stage_header(leaf, cin, sHdr[wib][hbuf]);
    cp_async_wait_all();
    __syncwarp();  // header visible to every lane; the previous cell's totals are fully flushed
    const double *sB = sHdr[wib][hbuf];
    const double *sG = sB + HDR_GEO;
    staged = nBegin < nEnd;
    if (staged) stage_header(nleaf, noff, sHdr[wib][hbuf ^ 1]);
    hbuf ^= 1;
    const double invV = sG[12];

    double acc[16];  // C fragments of the 4 x 2 tiles: acc[2 (2 t + n) + {0,1}]
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = 0.0;
    double vm0 = 0.0, vm1 = 0.0;  // kDiag: sum |v| dt of species 0 / 1 seen by this lane
    int cnt01 = 0;                // counts: species 0 in the low half, species 1 in the high half

    for (int base = begin; base < end; base += CHUNK) {
      const int np = min(CHUNK, end - base);
      __syncwarp();  // sB visible; previous chunk consumed
      // ---------------- phase 1: lane <-> particle ----------------
      const double x0 = nx0, x1 = nx1, x2 = nx2, pw = nw;
      double v0 = nv0, v1 = nv1, v2 = nv2;
      const int spec = nspec & 0x3f;  // masked at use: the load stays in flight during phase 2
      if (kGather && lane < np) {
        // the sorted copy (the scatter of the counting sort, fused): slot base+lane of dst <- slot nsrc of p
        const int o = base + lane;
        dst.x[0][o] = x0, dst.x[1][o] = x1, dst.x[2][o] = x2;
        dst.v[0][o] = v0, dst.v[1][o] = v1, dst.v[2][o] = v2;
        dst.w[o] = pw, dst.spec[o] = (uint8_t)nspec, dst.key[o] = cell, dst.ptr[o] = nptr;
        if (p.mu) dst.mu[o] = p.mu[nsrc];
        if (p.vpar) dst.vpar[o] = p.vpar[nsrc];
      }
      if (base + CHUNK < end) {  // (warp-uniform) the next chunk of this cell
        if (base + CHUNK + lane < end) {
          const int ip = kGather ? nsrc2 : base + CHUNK + lane;
          nsrc = ip;
          nx0 = p.x[0][ip], nx1 = p.x[1][ip], nx2 = p.x[2][ip];
          nv0 = p.v[0][ip], nv1 = p.v[1][ip], nv2 = p.v[2][ip];
          nw = p.w[ip], nspec = p.spec[ip];
          if (kGather) nptr = p.ptr[ip];
        }
        if (kGather) {  // the permutation entry a chunk further: of this cell, else of the first chunk of the next cell
          if (base + 2 * CHUNK < end) {
            if (base + 2 * CHUNK + lane < end) nsrc2 = perm[base + 2 * CHUNK + lane];
          } else if (nBegin + lane < nEnd) {
            nsrc2 = perm[nBegin + lane];
          }
        }
      } else if (nBegin < nEnd) {  // last chunk of this cell: request the first chunk of the warp's next cell
        pre = true;
        if (nBegin + lane < nEnd) {
          const int ip = kGather ? ((end - begin > CHUNK) ? nsrc2 : perm[nBegin + lane]) : nBegin + lane;
          nsrc = ip;
          nx0 = p.x[0][ip], nx1 = p.x[1][ip], nx2 = p.x[2][ip];
          nv0 = p.v[0][ip], nv1 = p.v[1][ip], nv2 = p.v[2][ip];
          nw = p.w[ip], nspec = p.spec[ip];
          if (kGather) nptr = p.ptr[ip];
        }
        if (kGather && nBegin + CHUNK + lane < nEnd) nsrc2 = perm[nBegin + CHUNK + lane];
      }
      if (lane < np) {
        const double LocalParticleWeight = sp.weight[spec] * pw;
        // local coordinates: CornerBased::InitStencil (pic_interpolation_routines.cpp:1090-1098)
        double xl[3];
        {
          const double xx[3] = {x0, x1, x2};
#pragma unroll
          for (int d = 0; d < 3; d++) {
            double xs = xx[d];
            const double xmx = sG[3 + d], dxc = sG[6 + d];
            if (fabs(xs - xmx) < 1e-10 * dxc) xs = xmx - 1e-10 * dxc;
            double r = (xs - sG[d]) * sG[9 + d];
            r -= (int)r;
            xl[d] = r;
          }
        }
        {
          // the 27 class weights u = XX[px] YY[py] ZZ[pz]: per-dimension pair products of the un-normalised corner weights
          // WeightPG (:2200)
          const double X0 = 1.0 - xl[0], X1 = xl[0], Y0 = 1.0 - xl[1], Y1 = xl[1], Z0 = 1.0 - xl[2], Z1 = xl[2];
          const double xx[3] = {X0 * X0, X0 * X1, X1 * X1};
          const double yy[3] = {Y0 * Y0, Y0 * Y1, Y1 * Y1}, zz[3] = {Z0 * Z0, Z0 * Z1, Z1 * Z1};
          double yz[9];
#pragma unroll
          for (int i = 0; i < 9; i++) yz[i] = yy[i / 3] * zz[i % 3];
#pragma unroll
          for (int i = 0; i < 27; i++) el[4 * i] = xx[i / 9] * yz[i % 9];
        }
        // B at the particle: cell-centred trilinear stencil on B_cur (:2100-2129); relative to this cell the
        // stencil cells are -1/0 (particle in the lower half) or 0/+1 (upper half) per dimension
        double B0 = 0.0, B1 = 0.0, B2 = 0.0;
        if (kCornerB) {
          // CornerBased::InitStencil on B_cur (:2102-2129): the trilinear corner weights (they sum to 1 within 2 ulp,
          // Normalize() changes B by <= 3e-16 relative)
          const double X0 = 1.0 - xl[0], X1 = xl[0], Y0 = 1.0 - xl[1], Y1 = xl[1], Z0 = 1.0 - xl[2], Z1 = xl[2];
          const double b00 = X0 * Y0, b01 = X0 * Y1, b10 = X1 * Y0, b11 = X1 * Y1;
          const double ws[8] = {b00 * Z0, b00 * Z1, b01 * Z0, b01 * Z1, b10 * Z0, b10 * Z1, b11 * Z0, b11 * Z1};
#pragma unroll
          for (int s = 0; s < 8; s++) {
            B0 = fma(ws[s], sB[3 * s], B0);
            B1 = fma(ws[s], sB[3 * s + 1], B1);
            B2 = fma(ws[s], sB[3 * s + 2], B2);
          }
          const double sc = sp.B_conv * invc;
          B0 *= sc, B1 *= sc, B2 *= sc;
        } else {
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
          if (!m.periodic && face) {  // AddCell drops centres outside the global box (pic.h:7235-7245), the rest is re-normalised
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
        if (kDiag) {  // ParticleEnergyCell, vmean_cell (:2228-2238)
          const double vsqr = v0 * v0 + v1 * v1 + v2 * v2;
          eAcc = fma(0.5 * (sp.mass[spec] * LocalParticleWeight), vsqr, eAcc);
          const double vabs = sqrt(vsqr) * sp.dt[0];
          if (spec == 0) vm0 += vabs, cnt01 += 1;
          else vm1 += vabs, cnt01 += 0x10000;
        }
        const double chargeQ = sp.charge[spec] * LocalParticleWeight;
        // beta = q~ dt / 2 m~ : the statistical weight cancels (to 1 ulp), so it is a per-species constant
        const double beta = sQdt2m[spec];
        const double s2 = beta * beta;
        const double c0 = fast_rcp(1.0 + s2 * (B0 * B0 + B1 * B1 + B2 * B2));
        // k alpha with k = matrixConst = q~ beta / V (:2311) and alpha = c0 (I - beta [B]x + beta^2 B B^T) (:1056-1067):
        // symmetric part S_ij = kc (delta_ij + beta^2 B_i B_j), antisymmetric part from kc beta B
        const double kc = (chargeQ * beta * invV) * c0;
        const double cB0 = kc * B0, cB1 = kc * B1, cB2 = kc * B2;
        const double t0 = s2 * B0, t1 = s2 * B1, t2 = s2 * B2;
        const double S01 = t0 * cB1, S02 = t0 * cB2, S12 = t1 * cB2;
        const double A0 = beta * cB0, A1 = beta * cB1, A2 = beta * cB2;  // kc * (-P_i)
        const double a0 = fma(t0, cB0, kc), a1 = S01 + A2, a2 = S02 - A1;
        const double a3 = S01 - A2, a4 = fma(t1, cB1, kc), a5 = S12 + A0;
        const double a6 = S02 + A1, a7 = S12 - A0, a8 = fma(t2, cB2, kc);
        // Jg/CellVolume = q~/V alpha v = (k alpha v)/beta (:2367)
        const double ib = sInvBeta[spec];
        double *a = el + 4 * EL_A;
        a[0] = a0, a[4] = a1, a[8] = a2, a[12] = a3, a[16] = a4, a[20] = a5, a[24] = a6, a[28] = a7, a[32] = a8;
        a[36] = ib * (a0 * v0 + a1 * v1 + a2 * v2);
        a[40] = ib * (a3 * v0 + a4 * v1 + a5 * v2);
        a[44] = ib * (a6 * v0 + a7 * v1 + a8 * v2);
      } else if (lane < ((np + 3) & ~3)) {
        // the last group of the cell is not full: zero class weights (the columns left there by earlier chunks are finite)
#pragma unroll
        for (int i = 0; i < 27; i++) el[4 * i] = 0.0;
      }
      __syncwarp();
      // ---------------- phase 2: T += U^T A on the fp64 MMA path, 4 particles per step ----------------
      {
        const int nks = (np + 3) >> 2;
        // rows 27..31 and columns 12..15 of the padded T are never read: their lanes load a clamped element.  (Predicating those
        // loads off -- one 128-byte wavefront instead of two -- was measured SLOWER, 2.39 instead of 2.22 ms: the loop-carried
        // operand registers break the software pipeline of the loop.)
#pragma unroll 2
        for (int ks = 0; ks < nks; ks++) {
          const double *q = rows + ks * GRP;
          const double fa0 = q[eA0], fa1 = q[eA1], fa2 = q[eA2], fa3 = q[eA3];
          const double fb0 = q[eB0], fb1 = q[eB1];
          dmma884(acc[0], acc[1], fa0, fb0);
          dmma884(acc[2], acc[3], fa0, fb1);
          dmma884(acc[4], acc[5], fa1, fb0);
          dmma884(acc[6], acc[7], fa1, fb1);
          dmma884(acc[8], acc[9], fa2, fb0);
          dmma884(acc[10], acc[11], fa2, fb1);
          dmma884(acc[12], acc[13], fa3, fb0);
          dmma884(acc[14], acc[15], fa3, fb1);
        }
      }
    }

    if (kDiag) {  // cfl of the cell per species (:2355-2359): sum|v|dt / (count |dx|)
      // one butterfly for both species: after the first exchange the lower half-warp holds partial sums of species 0, the upper
      // half those of species 1
      const bool lo = lane < 16;
      double a = lo ? vm0 : vm1;
      a += __shfl_xor_sync(0xffffffffu, lo ? vm1 : vm0, 16);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      const int c = __reduce_add_sync(0xffffffffu, cnt01);
      const int cs = lo ? (c & 0xffff) : (c >> 16);
      // lane 0: species 0, lane 16: species 1; 0/0 = NaN never wins the reference's '>' comparison.  cfl > cflMax is tested as
      // a > cflMax (count |dx|) (all factors positive), so the division only runs when the maximum moves
      const double den = cs * sG[13];
      if ((lane & 15) == 0 && cs > 0 && a > cflMax * den) cflMax = a / den;
    }
    // ---- spill the C fragments: T[class 8 t + g][column 8 n + 2 kk + {0,1}] (the K sum is complete, nothing to fold) ----
    __syncwarp();  // phase 2 finished reading the slab
#pragma unroll
    for (int t = 0; t < 4; t++)
#pragma unroll
      for (int n = 0; n < 2; n++)
        *reinterpret_cast<double2 *>(rows + (8 * t + g) * T_LD + 8 * n + 2 * kk) = make_double2(acc[2 * (2 * t + n)], acc[2 * (2 * t + n) + 1]);
    // columns 9, 10, 11 (the current) live in the n = 1 fragments of the lanes kk = 0 (8, 9) and kk = 1 (10, 11)
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int cls = 8 * t + g;
      if (cls < 27) {
        if (kk == 0) rows[T_J + 3 * cls] = acc[2 * (2 * t + 1) + 1];
        if (kk == 1) rows[T_J + 3 * cls + 1] = acc[2 * (2 * t + 1)], rows[T_J + 3 * cls + 2] = acc[2 * (2 * t + 1) + 1];
      }
    }
    __syncwarp();
    // ---- flush the mass matrix: 64 ordered corner pairs x 9 (both (c,c') and (c',c) get the same block, :2411-2420)
#pragma unroll
    for (int o = lane; o < 576; o += 32) {
      const unsigned e = sFlush[o];
      const int ui = __shfl_sync(0xffffffffu, uidLane, e & 7u);
      atomicAdd(M + (size_t)ui * 243 + ((e >> 3) & 255u), rows[e >> 11]);
    }
    // ---- current: J[c] = sum_c' T[cls(c,c')][9..11] ----
    {
      const int c = (lane < 24) ? lane / 3 : 0, dcol = (lane < 24) ? lane - 3 * c : 0;
      const int ui = __shfl_sync(0xffffffffu, uidLane, c);
      if (lane < 24) {
        double val = 0.0;
#pragma unroll
        for (int d = 0; d < 8; d++) val += rows[sJcls[c * 8 + d] + dcol];
        atomicAdd(J + (size_t)ui * 3 + dcol, val);
      }
    }
  }
  if (kGather && m.periodic && ghostPass) {
    // periodic "ghost" (boundary) blocks deposit nothing and hold no particle after the wrap; should one be there (uploaded
    // outside the real domain and not moved yet), it still travels to the sorted copy
    for (int g = m.nDepReal + warpGlobal; g < m.nLeaves; g += nWarps) {
      const int leaf = m.depLeaf[g];
      const int b = cellStart[(size_t)leaf * C], e = cellStart[(size_t)(leaf + 1) * C];
      for (int ip = b + lane; ip < e; ip += 32) {
        const int sidx = perm[ip];
        for (int d = 0; d < 3; d++) dst.x[d][ip] = p.x[d][sidx], dst.v[d][ip] = p.v[d][sidx];
        dst.w[ip] = p.w[sidx], dst.spec[ip] = p.spec[sidx], dst.key[ip] = p.key[sidx], dst.ptr[ip] = p.ptr[sidx];
        if (p.mu) dst.mu[ip] = p.mu[sidx];
        if (p.vpar) dst.vpar[ip] = p.vpar[sidx];
      }
    }
  }
  if (kDiag) {  // energy is added once per corner in the reference's flush loops (x8, :3860)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) eAcc += __shfl_xor_sync(0xffffffffu, eAcc, o);
    if (lane == 0 && eAcc != 0.0) atomicAdd(energyOut, 8.0 * eAcc);
    if ((lane & 15) == 0 && (lane >> 4) < sp.n && cflMax > 0.0) atomicMaxPositiveDouble(&cflBits[lane >> 4], cflMax);
  }
}

// ------------------------------------------------------------------------------------------------
// diagnostics of UpdateJMassMatrix: particle energy (x8: the reference adds the cell energy once per
// corner, :3860) and per-species cfl = max over cells of  sum|v| dt / (count |dx|)  (:2237, :2357-2359)
// one warp per cell, grid-stride
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) diag_kernel(DevMesh m, DevSpecies sp, ParticleSoA p, const int *__restrict__ cellStart,
                                                  double *__restrict__ energyOut, unsigned long long *__restrict__ cflBits) {
  const int lane = threadIdx.x & 31;
  const int warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nWarps = (gridDim.x * blockDim.x) >> 5;
  const int nCells = m.nLeaves * m.cellsPerBlock;
  const int C = m.cellsPerBlock;
  double e = 0.0;
  double cflMax = 0.0;  // lane s keeps species s
  for (int cell = warpGlobal; cell < nCells; cell += nWarps) {
    const int begin = cellStart[cell], end = cellStart[cell + 1];
    if (begin == end) continue;
    const LeafGeo &lg = m.leaf[cell / C];
    if (m.periodic && lg.face != 0) continue;
    double vm[AMPS_GPU_MAX_SPECIES];
    int cnt[AMPS_GPU_MAX_SPECIES];
#pragma unroll
    for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++) vm[s] = 0.0, cnt[s] = 0;
    for (int ip = begin + lane; ip < end; ip += 32) {
      const double v0 = p.v[0][ip] * sp.length_conv, v1 = p.v[1][ip] * sp.length_conv, v2 = p.v[2][ip] * sp.length_conv;
      const int spec = p.spec[ip] & 0x3f;
      const double mass = sp.mass[spec] * (sp.weight[spec] * p.w[ip]);
      const double vsqr = v0 * v0 + v1 * v1 + v2 * v2;
      e += 0.5 * mass * vsqr;
      const double vabs = sqrt(vsqr) * sp.dt[0];
#pragma unroll
      for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++)
        if (s == spec) vm[s] += vabs, cnt[s]++;
    }
#pragma unroll
    for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++) {
      if (s < sp.n) {
        double a = vm[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        const int c = __reduce_add_sync(0xffffffffu, cnt[s]);
        if (lane == s && c > 0) {  // 0/0 = NaN never wins the reference's '>' comparison
          const double cfl = a / (c * lg.diag);
          if (cfl > cflMax) cflMax = cfl;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if (lane == 0 && e != 0.0) atomicAdd(energyOut, 8.0 * e);
  if (lane < sp.n && cflMax > 0.0) atomicMaxPositiveDouble(&cflBits[lane], cflMax);
}

// ------------------------------------------------------------------------------------------------
// f4 (first half): ECSIM::ComputeNetCharge  src/pic/pic_field_solver_ecsim.cpp:4690-4828
// rho_new on the unique centre nodes: every particle spreads q~ over the 8 centres of its cell-centred trilinear stencil.
// Like the reference (a local q_Center array per block) one CTA accumulates a block's (N+2)^3 centres in shared memory
// (fp64 shared atomics) and flushes them once, divided by the cell volume, with global REDs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) net_charge_kernel(DevMesh m, DevSpecies sp, ParticleSoA p, const int *__restrict__ cellStart, double chargeConv,
                                                        int slices, double *__restrict__ rho) {
  extern __shared__ double sQ[];  // [TN0*TN1*TN2] block-local centre numbering
  const int leaf = blockIdx.x / slices, slice = blockIdx.x - leaf * slices;
  const int C = m.cellsPerBlock;
  const LeafGeo &lg = m.leaf[leaf];
  if (m.periodic && lg.face != 0) return;  // boundary "ghost" block (:4711-4722)
  const int begin = cellStart[(size_t)leaf * C], end = cellStart[(size_t)(leaf + 1) * C];
  const long long len = (long long)end - begin;
  const int b = begin + (int)(len * slice / slices), e = begin + (int)(len * (slice + 1) / slices);
  if (b >= e) return;
  for (int i = threadIdx.x; i < m.nCenterLocal; i += blockDim.x) sQ[i] = 0.0;
  __syncthreads();
  const int BS0 = m.TN[0], BS1 = m.TN[0] * m.TN[1];
  for (int ip = b + threadIdx.x; ip < e; ip += blockDim.x) {
    const double x[3] = {p.x[0][ip], p.x[1][ip], p.x[2][ip]};
    const int spec = p.spec[ip] & 0x3f;
    const double chargeQ = (sp.charge[spec] * chargeConv) * (sp.weight[spec] * p.w[ip]);
    double loc[3], wd[3];
    int o[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      loc[d] = (x[d] - lg.xmin[d]) / (lg.xmax[d] - lg.xmin[d]) * m.N[d];
      o[d] = (loc[d] < 0.5) ? -1 : (int)(loc[d] - 0.50);
      wd[d] = loc[d] - (o[d] + 0.5);
    }
    double w[8];
    w[0] = (1.0 - wd[0]) * (1.0 - wd[1]) * (1.0 - wd[2]);
    w[1] = (1.0 - wd[0]) * (1.0 - wd[1]) * wd[2];
    w[2] = (1.0 - wd[0]) * wd[1] * (1.0 - wd[2]);
    w[3] = (1.0 - wd[0]) * wd[1] * wd[2];
    w[4] = wd[0] * (1.0 - wd[1]) * (1.0 - wd[2]);
    w[5] = wd[0] * (1.0 - wd[1]) * wd[2];
    w[6] = wd[0] * wd[1] * (1.0 - wd[2]);
    w[7] = wd[0] * wd[1] * wd[2];
    unsigned valid = 0xffu;
    if (!m.periodic && lg.face) {  // AddCell drops centres outside the global box; the rest is re-normalised (Length != 8)
      if ((lg.face & 1) && o[0] < 0) valid &= 0xf0u;
      if ((lg.face & 2) && o[0] + 1 >= m.N[0]) valid &= 0x0fu;
      if ((lg.face & 4) && o[1] < 0) valid &= 0xccu;
      if ((lg.face & 8) && o[1] + 1 >= m.N[1]) valid &= 0x33u;
      if ((lg.face & 16) && o[2] < 0) valid &= 0xaau;
      if ((lg.face & 32) && o[2] + 1 >= m.N[2]) valid &= 0x55u;
    }
    double inv = 1.0;
    if (valid != 0xffu) {
      double norm = 0.0;
#pragma unroll
      for (int s = 0; s < 8; s++)
        if (valid & (1u << s)) norm += w[s];
      if (norm > 0.0) inv = 1.0 / norm;
    }
    const int nd0 = centerLocalNumber(m, o[0], o[1], o[2]);
#pragma unroll
    for (int s = 0; s < 8; s++)
      if (valid & (1u << s)) atomicAdd(&sQ[nd0 + ((s >> 2) & 1) + ((s >> 1) & 1) * BS0 + (s & 1) * BS1], (w[s] * inv) * chargeQ);
  }
  __syncthreads();
  const double invVol = lg.invdxc[0] * lg.invdxc[1] * lg.invdxc[2];  // 1/CellVolume, CellVolume = prod dx (:4737-4739)
  const int *cuid = m.centerUid + (size_t)leaf * m.nCenterLocal;
  for (int i = threadIdx.x; i < m.nCenterLocal; i += blockDim.x) {
    const double q = sQ[i];
    const int u = cuid[i];
    if (q != 0.0 && u >= 0) atomicAdd(&rho[u], q * invVol);
  }
}

void launch_net_charge(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *cellStart, double chargeConv, double *rho, long long nUpper,
                       cudaStream_t s) {
  cudaMemsetAsync(rho, 0, sizeof(double) * (size_t)m.nCenters, s);
  long long perLeaf = nUpper / (m.nLeaves > 0 ? m.nLeaves : 1);
  int slices = (int)((perLeaf + 8191) / 8192);
  if (slices < 1) slices = 1;
  if (slices > 32) slices = 32;
  net_charge_kernel<<<m.nLeaves * slices, 256, sizeof(double) * m.nCenterLocal, s>>>(m, sp, p, cellStart, chargeConv, slices, rho);
}

void launch_deposit(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *cellStart, const double *bCurTile, double *J, double *M,
                    double *energy, unsigned long long *cflBits, int nSM, const int *perm, ParticleSoA dst, int dep0, int dep1, unsigned flags,
                    cudaStream_t s, long long *launches) {
  if (dep1 < 0 || dep1 > m.nDepReal) dep1 = m.nDepReal;
  if (flags & DEP_ZERO_JM) {  // SetCornerNodeAssociatedDataValue, :3266-3267; later ranges of the same deposit add to them
    cudaMemsetAsync(J, 0, sizeof(double) * 3 * (size_t)m.nCorners, s);
    cudaMemsetAsync(M, 0, sizeof(double) * 243 * (size_t)m.nCorners, s);
  }
  if (flags & DEP_ZERO_DIAG) {
    cudaMemsetAsync(energy, 0, sizeof(double), s);
    cudaMemsetAsync(cflBits, 0, sizeof(unsigned long long) * AMPS_GPU_MAX_SPECIES, s);
  }
  // DEP_SPARE_SMS: a few SMs stay free for the exchange kernels that run next to this launch
  const int grid = ((flags & DEP_SPARE_SMS) && nSM > 32 ? nSM - 8 : nSM) * DEP_CTAS_PER_SM;
  const size_t smem = sizeof(double) * DEP_WARPS * SLAB;
  const bool corner = sp.bMode == AMPS_B_CORNER_BASED, diag = sp.n <= 2 && !(flags & DEP_NO_DIAG), gather = perm != nullptr;
  const int ghostPass = (flags & DEP_GHOST_PASS) ? 1 : 0;
#define AMPS_DEP_LAUNCH(CB, DG, GA)                                                                                          \
  do {                                                                                                                       \
    static OncePerDevice once;                                                                                               \
    if (once.first()) cudaFuncSetAttribute(deposit_kernel<CB, DG, GA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    deposit_kernel<CB, DG, GA><<<grid, DEP_THREADS, smem, s>>>(m, sp, p, cellStart, bCurTile, J, M, energy, cflBits, perm, dst, dep0, dep1, ghostPass); \
  } while (0)
  if (corner) {
    if (diag) { if (gather) AMPS_DEP_LAUNCH(true, true, true); else AMPS_DEP_LAUNCH(true, true, false); }
    else { if (gather) AMPS_DEP_LAUNCH(true, false, true); else AMPS_DEP_LAUNCH(true, false, false); }
  } else {
    if (diag) { if (gather) AMPS_DEP_LAUNCH(false, true, true); else AMPS_DEP_LAUNCH(false, true, false); }
    else { if (gather) AMPS_DEP_LAUNCH(false, false, true); else AMPS_DEP_LAUNCH(false, false, false); }
  }
#undef AMPS_DEP_LAUNCH
  (*launches) += 1;
  if (!diag && (flags & DEP_FINAL) && !(flags & DEP_NO_DIAG)) {
    // more than two species: the diagnostics run as their own pass over the SORTED store (after the last range)
    diag_kernel<<<nSM * 4, 256, 0, s>>>(m, sp, gather ? dst : p, cellStart, energy, cflBits);
    (*launches) += 1;
  }
}

}  // namespace amps
