// exchange.cu -- device side of the cross-rank exchanges (sm_100a), one rank per GPU.
//
//   a17 PIC::Parallel::ExchangeParticleData   src/pic/pic_parallel.cpp:50-488
//       The reference walks the per-cell lists of the boundary-layer blocks and sends {cell descriptor,
//       particle bytes}; here the mover has already written the new (block,cell) key of every particle, so a
//       leaver is any particle whose block is owned by another rank: it is appended to that rank's send
//       buffer (8 doubles: x,v,w and the GLOBAL cell id + species; 9 with the magnetic moment), removed from the local histogram and
//       marked deleted.  Arrivals are appended behind the resident particles with their key translated to the
//       local block numbering and counted into the histogram; the counting sort that follows files them.
//   a12 SyncMassMatrix / ProcessJMassMatrix    src/pic/ecsim/halo_sync.cpp:79-124, pic_field_solver_ecsim.cpp:1383
//       corners that two ranks deposit into exchange their partial J[3], M[243] and add.
// The transfers themselves are NCCL send/recv over NVLink issued by amps_gpu.cu on the context's stream.
#include "amps_dev.cuh"

namespace amps {

// one candidate: append it to its owner's send buffer when the owner is another rank
// peerRecv == nullptr: the record goes to region `dest` of the local send buffer (NCCL send/recv moves it);
// peerRecv != nullptr: it is written straight into region `me` of rank dest's receive buffer over NVLink (peer memory)
__device__ __forceinline__ void pack_one_leaver(const ParticleSoA &p, int i, int k, const int *__restrict__ leafOwner, const int *__restrict__ leafGlobal,
                                                int C, int me, double *__restrict__ sendBuf, long long capPerPeer, int *__restrict__ sendCount,
                                                int *__restrict__ cellCount, int *__restrict__ errFlag, double *const *__restrict__ peerRecv = nullptr) {
  const int leaf = k / C;
  const int dest = leafOwner[leaf];
  if (dest == me) return;
  const int slot = atomicAdd(&sendCount[dest], 1);
  if (slot >= capPerPeer) {
    atomicOr(errFlag, 1);
    return;  // the particle stays (in a foreign block); the host reports AMPS_GPU_ERR_CAPACITY
  }
  const int L = migration_record_len(p);
  double *r = peerRecv ? peerRecv[dest] + ((size_t)me * capPerPeer + slot) * L : sendBuf + ((size_t)dest * capPerPeer + slot) * L;
  const long long gkey = (long long)leafGlobal[leaf] * C + (k - leaf * C);
  r[0] = p.x[0][i], r[1] = p.x[1][i], r[2] = p.x[2][i];
  r[3] = p.v[0][i], r[4] = p.v[1][i], r[5] = p.v[2][i];
  r[6] = p.w[i];
  r[7] = __longlong_as_double((gkey << 8) | (long long)p.spec[i]);
  if (p.mu) r[8] = p.mu[i];
  if (p.vpar) r[L - 1] = p.vpar[i];
  atomicSub(&cellCount[k], 1);
  p.key[i] = -1;
}

__global__ void __launch_bounds__(256) pack_leavers_kernel(ParticleSoA p, const int *__restrict__ nSlots, const int *__restrict__ leafOwner,
                                                          const int *__restrict__ leafGlobal, int C, int me, double *__restrict__ sendBuf,
                                                          long long capPerPeer, int *__restrict__ sendCount, int *__restrict__ cellCount,
                                                          int *__restrict__ errFlag, double *const *__restrict__ peerRecv) {
  const int n = *nSlots;
  // leavers are rare: the scan reads four keys per 16-byte load (the key array is a 256-byte aligned allocation)
  const int n4 = (n + 3) >> 2;
  const int4 *key4 = reinterpret_cast<const int4 *>(p.key);
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += gridDim.x * blockDim.x) {
    const int4 kk = key4[q];
    const int ks[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
    const int i = 4 * q + j;
    const int k = ks[j];
    if (i >= n || k < 0) continue;
    pack_one_leaver(p, i, k, leafOwner, leafGlobal, C, me, sendBuf, capPerPeer, sendCount, cellCount, errFlag, peerRecv);
    }
  }
  if (peerRecv) __threadfence_system();  // the records must have reached the peers before the counts that announce them leave
}

// peer-memory variant: the arrivals of rank src lie in region src of this rank's receive buffer, their number in
// allCounts[src * R + me] (all-gathered on the device): no count ever visits the host
__global__ void __launch_bounds__(256) unpack_arrivals_peer_kernel(const double *__restrict__ recvBuf, const int *__restrict__ allCounts, int R,
                                                                  long long capPerPeer, ParticleSoA p, const int *__restrict__ nSlots,
                                                                  const int *__restrict__ g2l, const int *__restrict__ leafOwner, int C, int me,
                                                                  long long capacity, int *__restrict__ cellCount, int *__restrict__ errFlag) {
  const long long base = *nSlots;
  const int L = migration_record_len(p);
  long long off = 0;
  for (int src = 0; src < R; src++) {
    if (src == me) continue;
    int cnt = allCounts[(size_t)src * R + me];
    if (cnt > capPerPeer) {  // the sender overflowed its region (it flagged itself too)
      if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(errFlag, 1);
      cnt = (int)capPerPeer;
    }
    const double *reg = recvBuf + (size_t)src * capPerPeer * L;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
      const double *r = reg + (size_t)j * L;
      const long long meta = __double_as_longlong(r[7]);
      const long long gkey = meta >> 8;
      const int gleaf = (int)(gkey / C);
      const int leaf = g2l[gleaf];
      const long long i = base + off + j;
      if (leaf < 0 || leafOwner[leaf] != me || i >= capacity) {
        atomicOr(errFlag, 2);
        if (i < capacity) p.key[i] = -1;
        continue;
      }
      const int k = leaf * C + (int)(gkey - (long long)gleaf * C);
      p.x[0][i] = r[0], p.x[1][i] = r[1], p.x[2][i] = r[2];
      p.v[0][i] = r[3], p.v[1][i] = r[4], p.v[2][i] = r[5];
      p.w[i] = r[6];
      if (p.mu) p.mu[i] = r[8];
      if (p.vpar) p.vpar[i] = r[L - 1];
      p.spec[i] = (uint8_t)(meta & 0xff);
      p.key[i] = k;
      p.ptr[i] = -1;
      atomicAdd(&cellCount[k], 1);
    }
    off += cnt;
  }
}
__global__ void bump_count_peer_kernel(int *nSlots, const int *__restrict__ allCounts, int R, int me, long long capPerPeer, long long capacity,
                                       long long *__restrict__ sentRecv) {
  long long recv = 0, sent = 0;
  for (int r = 0; r < R; r++) {
    if (r == me) continue;
    const long long a = allCounts[(size_t)r * R + me], b = allCounts[(size_t)me * R + r];
    recv += a > capPerPeer ? capPerPeer : a;
    sent += b > capPerPeer ? capPerPeer : b;
  }
  const long long v = (long long)*nSlots + recv;
  *nSlots = (int)(v > capacity ? capacity : v);
  if (sentRecv) sentRecv[0] = sent, sentRecv[1] = recv;
}

__global__ void __launch_bounds__(256) unpack_arrivals_kernel(const double *__restrict__ recvBuf, int nRecv, ParticleSoA p, int *__restrict__ nSlots,
                                                             const int *__restrict__ g2l, const int *__restrict__ leafOwner, int C, int me,
                                                             long long capacity, int *__restrict__ cellCount, int *__restrict__ errFlag) {
  const int base = *nSlots;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nRecv; j += gridDim.x * blockDim.x) {
    const int L = migration_record_len(p);
    const double *r = recvBuf + (size_t)j * L;
    const long long meta = __double_as_longlong(r[7]);
    const long long gkey = meta >> 8;
    const int gleaf = (int)(gkey / C);
    const int leaf = g2l[gleaf];
    const long long i = (long long)base + j;
    if (leaf < 0 || leafOwner[leaf] != me || i >= capacity) {
      atomicOr(errFlag, 2);
      if (i < capacity) p.key[i] = -1;
      continue;
    }
    const int k = leaf * C + (int)(gkey - (long long)gleaf * C);
    p.x[0][i] = r[0], p.x[1][i] = r[1], p.x[2][i] = r[2];
    p.v[0][i] = r[3], p.v[1][i] = r[4], p.v[2][i] = r[5];
    p.w[i] = r[6];
    if (p.mu) p.mu[i] = r[8];
    if (p.vpar) p.vpar[i] = r[L - 1];
    p.spec[i] = (uint8_t)(meta & 0xff);
    p.key[i] = k;
    p.ptr[i] = -1;  // no ParticleBuffer slot on this rank yet (GetNewParticle on download)
    atomicAdd(&cellCount[k], 1);
  }
}
__global__ void bump_count_kernel(int *nSlots, int nRecv, long long capacity) {
  long long v = (long long)*nSlots + nRecv;
  *nSlots = (int)(v > capacity ? capacity : v);
}

__global__ void __launch_bounds__(256) pack_corners_kernel(const int *__restrict__ uids, int n, const double *__restrict__ J,
                                                          const double *__restrict__ M, double *__restrict__ buf) {
  const long long total = (long long)n * 246;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e / 246), q = (int)(e - (long long)c * 246);
    const int u = uids[c];
    buf[e] = (q < 3) ? J[(size_t)u * 3 + q] : M[(size_t)u * 243 + (q - 3)];
  }
}
__global__ void __launch_bounds__(256) add_corners_kernel(const int *__restrict__ uids, int n, double *__restrict__ J, double *__restrict__ M,
                                                         const double *__restrict__ buf) {
  const long long total = (long long)n * 246;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e / 246), q = (int)(e - (long long)c * 246);
    const int u = uids[c];
    if (q < 3) J[(size_t)u * 3 + q] += buf[e];
    else M[(size_t)u * 243 + (q - 3)] += buf[e];
  }
}

// The mass matrix is symmetric by construction: ProcessCell adds the SAME 3x3 block to corner c under neighbour c' and to corner
// c' under neighbour c (pic_field_solver_ecsim.cpp:2411-2420), so M[c][slot(d)] == M[c+d][slot(-d)].  The packed row of a corner
// keeps J[3] and the 14 neighbour slots of AMPS_GPU_JM_PACKED_SLOTS (self + one of each +-d pair): 129 of the 246 doubles.
struct PackedSlots {
  int s[14];
};
__global__ void __launch_bounds__(256) pack_jm_half_kernel(int uid0, int n, const double *__restrict__ J, const double *__restrict__ M,
                                                          double *__restrict__ out, PackedSlots slots) {
  const long long total = (long long)n * 129;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e / 129), q = (int)(e - (long long)c * 129);
    const size_t u = (size_t)uid0 + c;
    double v;
    if (q < 3) v = J[u * 3 + q];
    else {
      const int b = (q - 3) / 9, r = (q - 3) - 9 * b;
      v = M[u * 243 + 9 * slots.s[b] + r];
    }
    out[u * 129 + q] = v;
  }
}

// all peers in one launch: the lists of the peers are concatenated in rank order like the buffers; a corner shared with several
// ranks appears once per peer, hence the atomics
__global__ void __launch_bounds__(256) add_corners_atomic_kernel(const int *__restrict__ uids, int n, double *__restrict__ J, double *__restrict__ M,
                                                                const double *__restrict__ buf) {
  const long long total = (long long)n * 246;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e / 246), q = (int)(e - (long long)c * 246);
    const int u = uids[c];
    if (q < 3) atomicAdd(&J[(size_t)u * 3 + q], buf[e]);
    else atomicAdd(&M[(size_t)u * 243 + (q - 3)], buf[e]);
  }
}

static inline int grid_for(long long n) {
  long long g = (n + 255) / 256;
  if (g < 1) g = 1;
  if (g > 148 * 16) g = 148 * 16;
  return (int)g;
}

void launch_pack_leavers(const DevMesh &m, ParticleSoA p, const int *nSlots, long long nUpper, const int *leafOwner, const int *leafGlobal, int me,
                         double *sendBuf, long long capPerPeer, int *sendCount, int *cellCount, int *errFlag, double *const *peerRecv, cudaStream_t s) {
  pack_leavers_kernel<<<grid_for((nUpper + 3) / 4), 256, 0, s>>>(p, nSlots, leafOwner, leafGlobal, m.cellsPerBlock, me, sendBuf, capPerPeer, sendCount, cellCount,
                                                      errFlag, peerRecv);
}
void launch_unpack_arrivals_peer(const DevMesh &m, const double *recvBuf, const int *allCounts, int R, long long capPerPeer, ParticleSoA p, int *nSlots,
                                 const int *g2l, const int *leafOwner, int me, long long capacity, int *cellCount, int *errFlag, long long *sentRecv,
                                 cudaStream_t s) {
  unpack_arrivals_peer_kernel<<<148 * 4, 256, 0, s>>>(recvBuf, allCounts, R, capPerPeer, p, nSlots, g2l, leafOwner, m.cellsPerBlock, me, capacity,
                                                      cellCount, errFlag);
  bump_count_peer_kernel<<<1, 1, 0, s>>>(nSlots, allCounts, R, me, capPerPeer, capacity, sentRecv);
}
void launch_unpack_arrivals(const DevMesh &m, const double *recvBuf, int nRecv, ParticleSoA p, int *nSlots, const int *g2l, const int *leafOwner, int me,
                            long long capacity, int *cellCount, int *errFlag, cudaStream_t s) {
  if (nRecv <= 0) return;
  unpack_arrivals_kernel<<<grid_for(nRecv), 256, 0, s>>>(recvBuf, nRecv, p, nSlots, g2l, leafOwner, m.cellsPerBlock, me, capacity, cellCount, errFlag);
  bump_count_kernel<<<1, 1, 0, s>>>(nSlots, nRecv, capacity);
}
void launch_pack_jm_half(int uid0, int n, const double *J, const double *M, double *out, const int *slots14, cudaStream_t s) {
  PackedSlots ps;
  for (int i = 0; i < 14; i++) ps.s[i] = slots14[i];
  if (n > 0) pack_jm_half_kernel<<<grid_for((long long)n * 129), 256, 0, s>>>(uid0, n, J, M, out, ps);
}
void launch_pack_corners(const int *uids, int n, const double *J, const double *M, double *buf, cudaStream_t s) {
  if (n > 0) pack_corners_kernel<<<grid_for((long long)n * 246), 256, 0, s>>>(uids, n, J, M, buf);
}
void launch_add_corners_atomic(const int *uids, int n, double *J, double *M, const double *buf, cudaStream_t s) {
  if (n > 0) add_corners_atomic_kernel<<<grid_for((long long)n * 246), 256, 0, s>>>(uids, n, J, M, buf);
}
void launch_add_corners(const int *uids, int n, double *J, double *M, const double *buf, cudaStream_t s) {
  if (n > 0) add_corners_kernel<<<grid_for((long long)n * 246), 256, 0, s>>>(uids, n, J, M, buf);
}

}  // namespace amps
