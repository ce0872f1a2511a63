// sort.cu -- on-device counting sort of the particle SoA by (block,cell) key (sm_100a).
//
// Replaces the reference's per-cell doubly linked lists: the temp->first list swap at the end of
// MoveParticles (src/pic/pic_mover.cpp:1056-1088) and CreateParticleTable
// (src/pic/pic_pbuffer.cpp:1160-1310, K3 in SURVEY 2.5: count per cell, host prefix sum, fill).
//   1. histogram of keys        (normally produced by the mover itself)
//   2. exclusive scan -> cellStart[nCells+1]   (3-phase, all on device)
//   3. scatter of every SoA component into the ping-pong copy; deleted particles (key<0) drop out
// All of it is integer/byte work bound by HBM; no host synchronisation.
#include "amps_dev.cuh"

namespace amps {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// one atomic per distinct key in the warp (the input is nearly sorted: typically 1-3 keys per warp)
__device__ __forceinline__ int warp_aggregated_slot(int *__restrict__ counter, int key) {
  const unsigned mask = __match_any_sync(__activemask(), key);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(mask) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(&counter[key], __popc(mask));
  base = __shfl_sync(mask, base, leader);
  return base + __popc(mask & ((1u << lane) - 1u));
}

__global__ void histogram_kernel(const int *__restrict__ key, const int *__restrict__ nSrc, int *__restrict__ cellCount) {
  const int n = *nSrc;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int k = key[i];
    if (k >= 0) atomicAdd(&cellCount[k], 1);
  }
}

__device__ __forceinline__ int block_exclusive_scan(int v, int *sWarp, int &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) sWarp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < (blockDim.x >> 5)) ? sWarp[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    sWarp[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) sWarp[32] = wi;
  }
  __syncthreads();
  total = sWarp[32];
  const int r = incl - v + sWarp[warp];
  __syncthreads();
  return r;
}

// phase 1: per-tile sums
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const int *__restrict__ in, long long n, int *__restrict__ tileSum) {
  __shared__ int sWarp[33];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q++)
    if (base + q < n) s += in[base + q];
  int total;
  block_exclusive_scan(s, sWarp, total);
  if (threadIdx.x == 0) tileSum[blockIdx.x] = total;
}
// phase 2: one CTA scans the tile sums in place (exclusive) and writes the grand total
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_offsets_kernel(int *__restrict__ tileSum, int nTiles, int *__restrict__ totalOut) {
  __shared__ int sWarp[33];
  int carry = 0;
  for (int base = 0; base < nTiles; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const int v = (i < nTiles) ? tileSum[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, sWarp, total);
    if (i < nTiles) tileSum[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *totalOut = carry;
}
// phase 3: final exclusive scan; also writes out[n] = total
__global__ void __launch_bounds__(SCAN_THREADS) scan_write_kernel(const int *__restrict__ in, long long n, const int *__restrict__ tileOff,
                                                                 const int *__restrict__ totalIn, int *__restrict__ out) {
  __shared__ int sWarp[33];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q++) {
    v[q] = (base + q < n) ? in[base + q] : 0;
    s += v[q];
  }
  int total;
  int ex = block_exclusive_scan(s, sWarp, total) + tileOff[blockIdx.x];
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q++) {
    if (base + q < n) out[base + q] = ex;
    ex += v[q];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *totalIn;
}

struct ParticleRegs {
  double x0, x1, x2, v0, v1, v2, w;
  int k, pt;
  uint8_t sp;
};
__device__ __forceinline__ void load_particle(const ParticleSoA &src, int i, ParticleRegs &r) {
  r.x0 = src.x[0][i], r.x1 = src.x[1][i], r.x2 = src.x[2][i];
  r.v0 = src.v[0][i], r.v1 = src.v[1][i], r.v2 = src.v[2][i];
  r.w = src.w[i], r.sp = src.spec[i], r.pt = src.ptr[i];
}
__device__ __forceinline__ void store_particle(const ParticleSoA &dst, int pos, const ParticleRegs &r) {
  dst.x[0][pos] = r.x0, dst.x[1][pos] = r.x1, dst.x[2][pos] = r.x2;
  dst.v[0][pos] = r.v0, dst.v[1][pos] = r.v1, dst.v[2][pos] = r.v2;
  dst.w[pos] = r.w, dst.spec[pos] = r.sp, dst.key[pos] = r.k, dst.ptr[pos] = r.pt;
}

// two particles per thread and iteration: twice the loads in flight per warp (the kernel is latency bound)
__global__ void __launch_bounds__(256, 5) scatter_kernel(ParticleSoA src, ParticleSoA dst, const int *__restrict__ nSrc, const int *__restrict__ cellStart,
                                                     int *__restrict__ cellFill) {
  const int n = *nSrc;
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 2 * stride) {
    const int j = i + stride;
    ParticleRegs a, b;
    a.k = src.key[i];
    b.k = (j < n) ? src.key[j] : -1;
    if (a.k >= 0) load_particle(src, i, a);
    if (b.k >= 0) load_particle(src, j, b);
    if (a.k >= 0) {
      const int pos = cellStart[a.k] + warp_aggregated_slot(cellFill, a.k);
      store_particle(dst, pos, a);
      if (src.mu) dst.mu[pos] = src.mu[i];
      if (src.vpar) dst.vpar[pos] = src.vpar[i];
    }
    if (b.k >= 0) {
      const int pos = cellStart[b.k] + warp_aggregated_slot(cellFill, b.k);
      store_particle(dst, pos, b);
      if (src.mu) dst.mu[pos] = src.mu[j];
      if (src.vpar) dst.vpar[pos] = src.vpar[j];
    }
  }
}

// permutation only: perm[position in the (block,cell)-sorted order] = current slot.  8 B per particle instead of the 130 B of
// the full scatter; the deposit that follows gathers through perm and writes the sorted copy as a by-product.
// A thread takes four consecutive slots (one 16-byte load); the four rounds of warp-aggregated atomics are issued before any of
// their results is used, so that one trip of a warp has one DRAM and one L2 round trip for 128 keys (the kernel was latency bound
// at one key per trip and thread: 99 us for 3.4e7 keys).  Empty slots (key < 0) form their own match group, which has no atomic.
constexpr int PERM_VEC = 4;
constexpr int PERM_CTAS = 6;  // CTAs per SM: 48 warps, each with the keys of its next trip already in flight
                              // (finishing a trip one trip later, behind the next atomics, measured no faster: 64 registers, 4 CTAs)
__device__ __forceinline__ void perm_load_keys(const int *__restrict__ key, long long i0, int n, int (&k)[PERM_VEC]) {
  if (i0 + PERM_VEC <= n) {
    const int4 q = *reinterpret_cast<const int4 *>(key + i0);
    k[0] = q.x, k[1] = q.y, k[2] = q.z, k[3] = q.w;
  } else {
#pragma unroll
    for (int q = 0; q < PERM_VEC; q++) k[q] = (i0 + q < n) ? key[i0 + q] : -1;
  }
}
__global__ void __launch_bounds__(256, PERM_CTAS) perm_kernel(const int *__restrict__ key, const int *__restrict__ nSrc, const int *__restrict__ cellStart,
                                                             int *__restrict__ cellFill, int *__restrict__ perm) {
  const int n = *nSrc;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x * PERM_VEC;
  // the trip count is uniform over the warp (every lane takes part in the match / shuffle rounds).  The profile of the version
  // without it showed the two exposed latencies of a trip, the key load (35 % of the stall samples at the first MATCH) and the
  // atomics (28 % at the first SHFL): the keys of the NEXT trip are requested before the current one is processed.
  long long w0 = ((long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * PERM_VEC;
  int kn[PERM_VEC];
  if (w0 < n) perm_load_keys(key, w0 + lane * PERM_VEC, n, kn);
  for (; w0 < n; w0 += stride) {
    const long long i0 = w0 + lane * PERM_VEC;
    int k[PERM_VEC];
#pragma unroll
    for (int q = 0; q < PERM_VEC; q++) k[q] = kn[q];
    if (w0 + stride < n) perm_load_keys(key, i0 + stride, n, kn);
    unsigned mask[PERM_VEC];
    int base[PERM_VEC], start[PERM_VEC];
#pragma unroll
    for (int q = 0; q < PERM_VEC; q++) {
      mask[q] = __match_any_sync(0xffffffffu, k[q]);
      start[q] = k[q] >= 0 ? __ldg(cellStart + k[q]) : 0;
      base[q] = 0;
      if (k[q] >= 0 && lane == __ffs(mask[q]) - 1) base[q] = atomicAdd(&cellFill[k[q]], __popc(mask[q]));
    }
#pragma unroll
    for (int q = 0; q < PERM_VEC; q++) {
      base[q] = __shfl_sync(0xffffffffu, base[q], __ffs(mask[q]) - 1);
      if (k[q] >= 0) perm[start[q] + base[q] + __popc(mask[q] & ((1u << lane) - 1u))] = (int)(i0 + q);
    }
  }
}

size_t sort_scan_tmp_bytes(long long nCells) {
  const long long nTiles = (nCells + SCAN_TILE - 1) / SCAN_TILE;
  return (size_t)(nTiles + 2) * sizeof(int);
}

// perm != nullptr: only the permutation is produced (dst is not written)
void launch_sort(const DevMesh &m, ParticleSoA src, ParticleSoA dst, const int *nSrc, int *cellCount, int *cellStart, int *cellFill, int *nDst,
                 long long nUpper, bool countValid, void *scanTmp, int *perm, cudaStream_t s, long long *launches) {
  const long long nCells = (long long)m.nLeaves * m.cellsPerBlock;
  const int nTiles = (int)((nCells + SCAN_TILE - 1) / SCAN_TILE);
  int *tileSum = reinterpret_cast<int *>(scanTmp);
  const int pgrid = (int)((nUpper + 255) / 256 > 148 * 16 ? 148 * 16 : (nUpper + 255) / 256 < 1 ? 1 : (nUpper + 255) / 256);
  if (!countValid) {
    cudaMemsetAsync(cellCount, 0, sizeof(int) * nCells, s);
    histogram_kernel<<<pgrid, 256, 0, s>>>(src.key, nSrc, cellCount);
    (*launches)++;
  }
  scan_tile_sums_kernel<<<nTiles, SCAN_THREADS, 0, s>>>(cellCount, nCells, tileSum);
  scan_tile_offsets_kernel<<<1, SCAN_THREADS, 0, s>>>(tileSum, nTiles, nDst);
  scan_write_kernel<<<nTiles, SCAN_THREADS, 0, s>>>(cellCount, nCells, tileSum, nDst, cellStart);
  cudaMemsetAsync(cellFill, 0, sizeof(int) * nCells, s);
  if (perm) {
    const long long want = (nUpper + 256 * PERM_VEC - 1) / (256 * PERM_VEC);
    perm_kernel<<<(int)(want > 148 * PERM_CTAS ? 148 * PERM_CTAS : want < 1 ? 1 : want), 256, 0, s>>>(src.key, nSrc, cellStart, cellFill, perm);
  }
  else scatter_kernel<<<pgrid, 256, 0, s>>>(src, dst, nSrc, cellStart, cellFill);
  (*launches) += 4;
}

}  // namespace amps
