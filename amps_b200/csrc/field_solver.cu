// field_solver.cu -- the field half of the ECSIM step on the device, sm_100a, fp64.
//
//   ECSIM::TimeStep                 src/pic/pic_field_solver_ecsim.cpp:6004-6157
//   GetStencil (one matrix row)     src/pic/ecsim/get_stencil.cpp: identity + (theta c dt)^2 (grad div - laplace) + 4 pi theta dt M
//   UpdateRhs / SampleRhsScalar     src/pic/ecsim/update_rhs.cpp
//   UpdateMatrixElement             :6004-6020  value = parameter + 4 pi dt theta * (mass-matrix entry)
//   cLinearSystemCornerNode::MultiplyVector / Solve   srcInterface/LinearSystemCornerNode.h:2749-3067, :3212 (GMRES of the SWMF library)
//   UpdateB :5160, UpdateE :5909
//
// The reference stores an explicit sparse row per unknown (81 elements with pointers into the corner buffers) and walks it for every
// product.  Here the operator is matrix-free on the unique corner nodes: the constant part is 243 numbers (3 x 3 components x 27
// neighbour slots, the compact tables of InitDiscritizationStencil) in constant memory, the variable part IS the mass matrix the
// deposit left in HBM (M[corner][9 slot + 3 p + q], the reference's MassMatrixOffsetTable), so a product reads M once, coalesced
// (1944 B per corner), one warp per corner: 510 MB per product for the 64^3 box = HBM-bound, nothing is assembled.  The Krylov
// vectors (6.3 MB each) live in L2.  GMRES(m) with classical Gram-Schmidt: all inner products of an iteration in ONE pass
// (multi_dot_kernel), the update and the norm of the new direction in one more (orthogonalize_kernel), so an iteration is four
// launches and one 8-byte-per-vector read-back for the Givens rotations on the host.
#include "amps_dev.cuh"

namespace amps {

// K[9 slot + 3 p + q], the parameter part of row component p, column component q, neighbour slot: 243 doubles, copied to shared
// memory by every CTA (constant memory would serialise: the lanes of a warp read different entries)

// kRhs = false: y = A x.  kRhs = true: y = -(A - I) E - f J + theta c dt curl B   (UpdateRhs; x = E^n)
// nb[c][27] neighbour corner per slot (-1: none -> the row is the boundary row dE = 0), cc[c][8] the eight cells around the corner
// Lane mapping: entry e = 9 slot + 3 p + q of a corner's row is handled by lane (e mod 27) in trip e / 27, so a lane keeps ONE (p, q)
// for the whole row (e mod 9 = lane mod 9) and its neighbour slot is lane / 9 + 3 trip: one accumulator per lane, nine 216-byte
// coalesced loads of M per corner, and the three row sums fall out of four shuffles.  (27 of 32 lanes work.)
template <bool kRhs>
__global__ void __launch_bounds__(256, 4) ecsim_operator_kernel(int nCorners, const int *__restrict__ nb, const int *__restrict__ cc,
                                                               const double *__restrict__ Kc, const double *__restrict__ M,
                                                               const double *__restrict__ x, double f, const double *__restrict__ J,
                                                               const double *__restrict__ B, double c4x, double c4y, double c4z,
                                                               double *__restrict__ y) {
  __shared__ double sK[243];
  for (int e = threadIdx.x; e < 243; e += blockDim.x) sK[e] = Kc[e] - ((kRhs && (e == 0 || e == 4 || e == 8)) ? 1.0 : 0.0);  // slot 0, p == q
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
  const bool work = lane < 27;
  const int l27 = work ? lane : 0, q = l27 % 3, s0 = l27 / 9;
  for (int c = warp; c < nCorners; c += nWarps) {
    const int myNb = work ? nb[(size_t)c * 27 + lane] : 0;
    const bool boundary = __any_sync(0xffffffffu, myNb < 0);
    const double *Mc = M + (size_t)c * 243 + l27;
    double m[9], xv[9];
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const int n = max(__shfl_sync(0xffffffffu, myNb, s0 + 3 * t), 0);
      m[t] = Mc[27 * t];
      xv[t] = x[(size_t)3 * n + q];
    }
    double a = 0.0;
#pragma unroll
    for (int t = 0; t < 9; t++) a = fma(fma(f, m[t], sK[l27 + 27 * t]), xv[t], a);
    if (!work) a = 0.0;
    // lanes l, l + 9, l + 18 hold the three slot groups of one (p, q); then the three q of a p
    a += __shfl_down_sync(0xffffffffu, a, 9) + __shfl_down_sync(0xffffffffu, a, 18);
    a += __shfl_down_sync(0xffffffffu, a, 1) + __shfl_down_sync(0xffffffffu, a, 2);
    const double ap = __shfl_sync(0xffffffffu, a, 3 * (lane < 3 ? lane : 0));  // lane p (< 3) takes the sum of row component p
    double curl = 0.0;
    if (kRhs) {
      // curl B^n from the 2 x 2 face averages of the centre values (get_stencil.cpp: coeff4 = 0.25 theta c dt / dx): 16 signed
      // centre values per component; lane 8 comp + j takes the face-average point (a, b) = j >> 1 and the first / second difference
      double t = 0.0;
      if (lane < 24 && !boundary) {
        const int *cell = cc + (size_t)c * 8;
        auto Bv = [&](int aa, int b, int d, int comp) { return B[(size_t)3 * cell[(aa + 1) + 2 * (b + 1) + 4 * (d + 1)] + comp]; };
        const int comp = lane >> 3, j = lane & 7, aa = -((j >> 2) & 1), b = -((j >> 1) & 1), second = j & 1;
        if (comp == 0) t = second ? -c4z * (Bv(aa, b, 0, 1) - Bv(aa, b, -1, 1)) : c4y * (Bv(aa, 0, b, 2) - Bv(aa, -1, b, 2));
        if (comp == 1) t = second ? -c4x * (Bv(0, b, aa, 2) - Bv(-1, b, aa, 2)) : c4z * (Bv(aa, b, 0, 0) - Bv(aa, b, -1, 0));
        if (comp == 2) t = second ? -c4y * (Bv(b, 0, aa, 0) - Bv(b, -1, aa, 0)) : c4x * (Bv(0, b, aa, 1) - Bv(-1, b, aa, 1));
      }
      t += __shfl_xor_sync(0xffffffffu, t, 4);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      curl = __shfl_sync(0xffffffffu, t, 8 * (lane < 3 ? lane : 0));
    }
    if (lane >= 3) continue;
    if (boundary) {  // dE = 0 on a domain boundary (GetStencil: unit row, zero right-hand side)
      y[(size_t)3 * c + lane] = kRhs ? 0.0 : x[(size_t)3 * c + lane];
      continue;
    }
    y[(size_t)3 * c + lane] = kRhs ? (-ap - f * J[(size_t)3 * c + lane] + curl) : ap;
  }
}

// out[j] += V_j . w for j < nVec (V_j = V + j ld), out[nVec] += w . w; every block covers one contiguous chunk of rows, so w is read
// from HBM / L2 once and stays in L1 while the block walks the vectors
// mask (several ranks): 1 for the entries of the corners this rank is the primary owner of, 0 elsewhere, so that the all-reduced
// sums count every physical corner once
__global__ void __launch_bounds__(256) multi_dot_kernel(const double *__restrict__ V, size_t ld, int nVec, const double *__restrict__ w, int n,
                                                       double *__restrict__ out, const unsigned char *__restrict__ mask) {
  __shared__ double sPart[8];
  const int per = (n + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = min(n, r0 + per);
  for (int j = 0; j <= nVec; j++) {
    const double *v = (j < nVec) ? V + (size_t)j * ld : w;
    double s = 0.0;
    if (mask) {
      for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
        if (mask[i / 3]) s = fma(v[i], w[i], s);
    } else {
      for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) s = fma(v[i], w[i], s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sPart[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int q = 0; q < (int)(blockDim.x >> 5); q++) t += sPart[q];
      atomicAdd(out + j, t);
    }
    __syncthreads();
  }
}

// w -= sum_j h[j] V_j  (classical Gram-Schmidt with the inner products of multi_dot_kernel); norm2 += |w|^2 of the result
__global__ void __launch_bounds__(256) orthogonalize_kernel(const double *__restrict__ V, size_t ld, int nVec, const double *__restrict__ h,
                                                           double *__restrict__ w, int n, double *__restrict__ norm2,
                                                           const unsigned char *__restrict__ mask) {
  __shared__ double sPart[8];
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double a = w[i];
    for (int j = 0; j < nVec; j++) a = fma(-h[j], V[(size_t)j * ld + i], a);
    w[i] = a;
    if (!mask || mask[i / 3]) s = fma(a, a, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sPart[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); q++) t += sPart[q];
    atomicAdd(norm2, t);
  }
}

// out = alpha a + beta b (out may alias a or b); alpha, beta may come from device scalars: alpha *= 1/sqrt(*invSqrtOf) when given
__global__ void __launch_bounds__(256) axpby_kernel(int n, double alpha, const double *a, double beta, const double *b, const double *invSqrtOf,
                                                   double *out) {
  if (invSqrtOf) alpha *= rsqrt(*invSqrtOf);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = alpha * a[i] + (b ? beta * b[i] : 0.0);
}

// x += sum_j y[j] V_j
__global__ void __launch_bounds__(256) combine_kernel(const double *__restrict__ V, size_t ld, int nVec, const double *__restrict__ yv,
                                                     double *__restrict__ x, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double a = x[i];
    for (int j = 0; j < nVec; j++) a = fma(yv[j], V[(size_t)j * ld + i], a);
    x[i] = a;
  }
}

// UpdateB (:5160): B^{n+1} = B^n + sum over the four edge pairs of a face of -/+ 0.25 c dt / dx (E differences), E = E^{n+theta};
// zc[z][8] the corner nodes of cell z (ii + 2 jj + 4 kk)
__global__ void __launch_bounds__(256) update_B_kernel(int nCenters, const int *__restrict__ zc, const double *__restrict__ Eh,
                                                      const double *__restrict__ Bn, double c4x, double c4y, double c4z, double *__restrict__ Bout) {
  for (int z = blockIdx.x * blockDim.x + threadIdx.x; z < nCenters; z += gridDim.x * blockDim.x) {
    const int *cn = zc + (size_t)z * 8;
    if (cn[0] < 0 || cn[7] < 0) {  // not a cell of this rank's domain (ghost centre of an open boundary)
      for (int d = 0; d < 3; d++) Bout[(size_t)3 * z + d] = Bn[(size_t)3 * z + d];
      continue;
    }
    double E[2][2][2][3];
    for (int kk = 0; kk < 2; kk++)
      for (int jj = 0; jj < 2; jj++)
        for (int ii = 0; ii < 2; ii++)
          for (int d = 0; d < 3; d++) E[ii][jj][kk][d] = Eh[(size_t)3 * cn[ii + 2 * jj + 4 * kk] + d];
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int a = 0; a < 2; a++)
      for (int b = 0; b < 2; b++) {
        t0 += -c4y * (E[a][1][b][2] - E[a][0][b][2]) + c4z * (E[a][b][1][1] - E[a][b][0][1]);
        t1 += -c4z * (E[a][b][1][0] - E[a][b][0][0]) + c4x * (E[1][a][b][2] - E[0][a][b][2]);
        t2 += -c4x * (E[1][a][b][1] - E[0][a][b][1]) + c4y * (E[a][1][b][0] - E[a][0][b][0]);
      }
    Bout[(size_t)3 * z] = Bn[(size_t)3 * z] + t0, Bout[(size_t)3 * z + 1] = Bn[(size_t)3 * z + 1] + t1, Bout[(size_t)3 * z + 2] = Bn[(size_t)3 * z + 2] + t2;
  }
}

static inline int grid_rows(long long n, int perThread = 1) {
  long long g = (n + 255LL * perThread) / (256LL * perThread);
  if (g < 1) g = 1;
  if (g > 148 * 8) g = 148 * 8;
  return (int)g;
}

void launch_ecsim_operator(bool rhs, int nCorners, const int *nb, const int *cc, const double *Kc, const double *M, const double *x, double f,
                           const double *J, const double *B, const double c4[3], double *y, cudaStream_t s) {
  const int grid = 148 * 8;  // 8 warps per CTA, one corner per warp and trip
  if (rhs) ecsim_operator_kernel<true><<<grid, 256, 0, s>>>(nCorners, nb, cc, Kc, M, x, f, J, B, c4[0], c4[1], c4[2], y);
  else ecsim_operator_kernel<false><<<grid, 256, 0, s>>>(nCorners, nb, cc, Kc, M, x, f, nullptr, nullptr, 0.0, 0.0, 0.0, y);
}
void launch_multi_dot(const double *V, size_t ld, int nVec, const double *w, int n, double *out, const unsigned char *mask, cudaStream_t s,
                      bool zero) {
  if (zero) cudaMemsetAsync(out, 0, sizeof(double) * (nVec + 1), s);
  multi_dot_kernel<<<148 * 4, 256, 0, s>>>(V, ld, nVec, w, n, out, mask);
}
void launch_orthogonalize(const double *V, size_t ld, int nVec, const double *h, double *w, int n, double *norm2, const unsigned char *mask,
                          cudaStream_t s, bool zero) {
  if (zero) cudaMemsetAsync(norm2, 0, sizeof(double), s);
  orthogonalize_kernel<<<grid_rows(n), 256, 0, s>>>(V, ld, nVec, h, w, n, norm2, mask);
}

// field halo (several ranks): buf[3 i + d] <- vec[3 uid[i] + d] and back
__global__ void __launch_bounds__(256) halo_pack_kernel(const int *__restrict__ uid, int n, const double *__restrict__ vec, double *__restrict__ buf) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 3 * n; e += gridDim.x * blockDim.x) buf[e] = vec[(size_t)3 * uid[e / 3] + e % 3];
}
__global__ void __launch_bounds__(256) halo_unpack_kernel(const int *__restrict__ uid, int n, const double *__restrict__ buf, double *__restrict__ vec) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 3 * n; e += gridDim.x * blockDim.x) vec[(size_t)3 * uid[e / 3] + e % 3] = buf[e];
}
void launch_halo_pack(const int *uid, int n, const double *vec, double *buf, cudaStream_t s) {
  if (n > 0) halo_pack_kernel<<<grid_rows(3LL * n), 256, 0, s>>>(uid, n, vec, buf);
}
void launch_halo_unpack(const int *uid, int n, const double *buf, double *vec, cudaStream_t s) {
  if (n > 0) halo_unpack_kernel<<<grid_rows(3LL * n), 256, 0, s>>>(uid, n, buf, vec);
}
void launch_axpby(int n, double alpha, const double *a, double beta, const double *b, const double *invSqrtOf, double *out, cudaStream_t s) {
  axpby_kernel<<<grid_rows(n), 256, 0, s>>>(n, alpha, a, beta, b, invSqrtOf, out);
}
void launch_combine(const double *V, size_t ld, int nVec, const double *y, double *x, int n, cudaStream_t s) {
  combine_kernel<<<grid_rows(n), 256, 0, s>>>(V, ld, nVec, y, x, n);
}
void launch_update_B(int nCenters, const int *zc, const double *Eh, const double *Bn, const double c4[3], double *Bout, cudaStream_t s) {
  update_B_kernel<<<grid_rows(nCenters), 256, 0, s>>>(nCenters, zc, Eh, Bn, c4[0], c4[1], c4[2], Bout);
}

}  // namespace amps
