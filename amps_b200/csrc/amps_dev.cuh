// amps_dev.cuh -- device-side data model shared by the kernels of libamps_gpu.so (sm_100a).
//
// HBM layout (all library-owned):
//   particles   SoA, sorted by key = leaf*cellsPerBlock + cell :  x[3][cap] v[3][cap] w[cap] (f64)
//               spec[cap] (u8: bits 0-5 species, bit 6 InitFlag as in the reference record)  key[cap] (i32, -1 = deleted)  ptr[cap] (i32, ParticleBuffer slot);
//               two copies (ping-pong for the counting sort)
//   cell table  cellStart[nCells+1] (i32) : particle range of each cell
//   mesh        flattened cTreeNodeAMR arrays + per-leaf LeafGeo + unique node tables
//   fields      unique-node arrays E_half[nCorners][3], B_prev/B_cur[nCenters][3] and per-leaf
//               staged tiles (SetBlock_E / SetBlock_B result) padded to 16 B for TMA bulk copies
//   J, M        J[nCorners][3], M[nCorners][243]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/amps_gpu.h"

namespace amps {

struct LeafGeo {
  double xmin[3], xmax[3];
  int imin[3];
  int isize;
  int level;
  int flags;  // AMPS_NODE_*
  int real;   // paired real leaf of a periodic ghost leaf, else -1
  int face;   // bit f: no neighbour across face f
  int node;
  int neib;   // minNeibRefinmentLevel (low 16 bits, signed) | maxNeibRefinmentLevel << 16   (meshAMRgeneric.h:829)
  // derived per-leaf constants of the deposit (host computed at mesh upload)
  double dxc[3];     // (xmax-xmin)/N                       CornerBased::InitStencil :1086-1088
  double invdxc[3];  // 1/dxc
  double invV;       // 1/CellVolume, CellVolume = prod(dxc*length_conv)   ProcessCell :1959-1961
  double diag;       // sqrt(sum (dxc*length_conv)^2)       ProcessCell :2358
};

struct DevMesh {
  int N[3], g[3], TN[3];
  int nRoot[3];
  int L;  // max_refinement_level
  int nNodes, nLeaves, nCorners, nCenters;
  int cellsPerBlock, nCornerLocal, nCenterLocal;
  int eTileStride, bTileStride;  // doubles per leaf tile (padded to even => 16 B multiples)
  int periodic;
  double xGlobalMin[3], xGlobalMax[3], dxMaxRef[3], dxRoot[3], eps;
  const int *child;     // [nNodes][8]
  const int *imin;      // [nNodes][3]
  const int *isize;     // [nNodes]
  const int *nodeLeaf;  // [nNodes]
  const int *nodeFlags; // [nNodes]
  const int *nodeLevel; // [nNodes]
  const double *nxmin;  // [nNodes][3]
  const double *nxmax;  // [nNodes][3]
  const int *rootNode;  // [nRoot0*nRoot1*nRoot2]
  const LeafGeo *leaf;  // [nLeaves]
  const int *cornerUid; // [nLeaves][nCornerLocal]
  const int *centerUid; // [nLeaves][nCenterLocal]
  // deposit order: the nDepReal leaves that deposit (on several ranks: those that touch a shared corner first), then the
  // periodic "ghost" leaves that do not
  const int *depLeaf;  // [nLeaves]
  int nDepReal;
  // structure cache of the coupler's AMR stencil (cplr_stencil.cuh; built by launch_build_cplr_cache at the first test-particle
  // move on a refined mesh, nullptr before / when switched off)
  const int *neib26;           // [nLeaves][27] node across (sx,sy,sz) in {-1,0,+1}^3 of the leaf's block (cs_neib), slot sx+1+3(sy+1)+9(sz+1)
  const int *mbSlot;           // [nLeaves] table of the leaf's block as the coarse block of a multi-block stencil, -1 = none
  const unsigned char *mbTab;  // [tables][(N0+2)(N1+2)(N2+2)][MB_ENTRY]
};

struct DevSpecies {
  int n;
  int timeStepMode;
  int bMode;
  int boundaryMode;
  int globalStencilFull;  // the reference's global StencilTable holds an 8-cell stencil (ComputeNetCharge ran): full B stencils are not
                          // normalised any more (pic_interpolation_routines.cpp:903); amps_gpu_global_stencil_set
  int pad0;
  double charge[AMPS_GPU_MAX_SPECIES], mass[AMPS_GPU_MAX_SPECIES], weight[AMPS_GPU_MAX_SPECIES], dt[AMPS_GPU_MAX_SPECIES];
  double dtTotal, B_conv, length_conv, LightSpeed;
};

struct ParticleSoA {
  double *x[3];
  double *v[3];
  double *w;
  uint8_t *spec;
  int *key;
  int *ptr;
  double *mu;  // magnetic moment (guiding-centre movers), nullptr unless cfg.carry_magnetic_moment
  double *vpar;  // v_parallel (gyrokinetic movers), nullptr unless cfg.carry_v_parallel
};

// counters written by the mover (device copy of amps_gpu_move_stats + error word)
struct DevMoveStats {
  unsigned long long n_moved, n_cross_cell, n_cross_block, n_left_domain, n_not_in_use, n_periodic_wrap, n_error;
  unsigned long long n_sub_steps;
};

// ---- meshAMRgeneric.h:74-75 ----
__device__ __forceinline__ int cornerLocalNumber(const DevMesh &m, int i, int j, int k) {
  return i + m.g[0] + (1 + m.TN[0]) * (j + m.g[1] + (k + m.g[2]) * (1 + m.TN[1]));
}
__device__ __forceinline__ int centerLocalNumber(const DevMesh &m, int i, int j, int k) {
  return i + m.g[0] + m.TN[0] * (j + m.g[1] + (k + m.g[2]) * m.TN[1]);
}

// "done once per device" flags of the launch helpers (function attributes are per device; a process may drive several)
struct OncePerDevice {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// launch_deposit flags
enum : unsigned {
  DEP_ZERO_JM = 1,     // zero J, M first
  DEP_ZERO_DIAG = 2,   // zero the energy / cfl accumulators first
  DEP_GHOST_PASS = 4,  // (gather) also copy the particles of the periodic ghost leaves into the sorted store
  DEP_FINAL = 8,       // last range of a deposit: run the >2-species diagnostics pass
  DEP_SPARE_SMS = 16,  // leave a few SMs to concurrently running exchange kernels
  DEP_NO_DIAG = 32,    // neither the fused nor the separate energy / cfl pass (guiding-centre species: launch_gc_deposit does them)
  DEP_ALL = DEP_ZERO_JM | DEP_ZERO_DIAG | DEP_GHOST_PASS | DEP_FINAL
};

// mover_tp.cu: fills neib26 [nLeaves][27] and the tables of the nTab leaves tabLeaf[] (MB_ENTRY bytes per dual cell)
size_t cplr_cache_table_bytes(const DevMesh &m);
void launch_build_cplr_cache(const DevMesh &m, int *neib26, const int *tabLeaf, int nTab, unsigned char *tab, cudaStream_t s);

// deposit_gc.cu: the guiding-centre species of cfg.gc_species_mask + the energy / cfl diagnostics of all species, on the sorted store
void launch_gc_deposit(const DevMesh &m, const DevSpecies &sp, unsigned gcMask, ParticleSoA p, const int *cellStart, const double *bCurTile,
                       const double *vnByPtr, long long nVn, double *J, double *energy, unsigned long long *cflBits, int nSM, cudaStream_t s);

// launch helpers (defined per TU that needs them)
// field_solver.cu
void launch_ecsim_operator(bool rhs, int nCorners, const int *nb, const int *cc, const double *Kc, const double *M, const double *x, double f,
                           const double *J, const double *B, const double c4[3], double *y, cudaStream_t s);
// zero = false: the caller cleared the accumulators (the Arnoldi loop clears all its columns with one memset)
void launch_multi_dot(const double *V, size_t ld, int nVec, const double *w, int n, double *out, const unsigned char *mask, cudaStream_t s,
                      bool zero = true);
void launch_orthogonalize(const double *V, size_t ld, int nVec, const double *h, double *w, int n, double *norm2, const unsigned char *mask,
                          cudaStream_t s, bool zero = true);
void launch_halo_pack(const int *uid, int n, const double *vec, double *buf, cudaStream_t s);
void launch_halo_unpack(const int *uid, int n, const double *buf, double *vec, cudaStream_t s);
void launch_axpby(int n, double alpha, const double *a, double beta, const double *b, const double *invSqrtOf, double *out, cudaStream_t s);
void launch_combine(const double *V, size_t ld, int nVec, const double *y, double *x, int n, cudaStream_t s);
void launch_update_B(int nCenters, const int *zc, const double *Eh, const double *Bn, const double c4[3], double *Bout, cudaStream_t s);
void launch_stage_tiles(const DevMesh &m, bool cornerB, const double *E_half, const double *B_prev, const double *B_cur, double *eTile, double *bPrevTile,
                        double *bCurTile, cudaStream_t s);
void launch_move_lapenta(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *cellStart, const double *eTile, const double *bTile,
                         int *cellCount, DevMoveStats *stats, int slices, amps_gpu_exit_record *exitBuf, unsigned long long *exitCount,
                         long long exitCap, const unsigned char *redoMask, const int *redoLeafList, const int *nRedoLeaves, cudaStream_t s);
// leafRedo[nLeaves] (flagged particles per block), redoLeafList[nLeaves] + nRedoLeaves (blocks with any), all zeroed by the caller
void launch_move_lapenta_fast(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *cellStart, const double *eTile, const double *bTile,
                              int *cellCount, DevMoveStats *stats, int slices, unsigned char *redoMask, int *leafRedo, int *redoLeafList,
                              int *nRedoLeaves, cudaStream_t s);
void launch_sort(const DevMesh &m, ParticleSoA src, ParticleSoA dst, const int *nSrc, int *cellCount, int *cellStart, int *cellFill, int *nDst,
                 long long capacity, bool countValid, void *scanTmp, int *perm, cudaStream_t s, long long *launches);
// perm != nullptr: p is the UNSORTED store, particle i of the sorted order is p[perm[i]]; the kernel also writes the sorted copy to dst
void launch_deposit(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *cellStart, const double *bCurTile, double *J, double *M,
                    double *energy, unsigned long long *cflBits, int nSM, const int *perm, ParticleSoA dst, int dep0, int dep1, unsigned flags,
                    cudaStream_t s, long long *launches);  // cells [cell0, cell1); cell1 < 0 = all; J, M and the diagnostics are zeroed when cell0 == 0
void launch_move_gyrokinetic(const DevMesh &m, const DevSpecies &sp, int order, int interp, double rSphere, ParticleSoA p, const int *nSlots,
                             long long nUpper, const double *bgTile, const double *gradBTile, const double *uE, const double *uB,
                             const double *uGradB, int *cellCount, DevMoveStats *stats, cudaStream_t s);
void launch_sample_cells(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *cellStart, double *sample, unsigned long long *nSampled,
                         int nSM, cudaStream_t s);
void launch_pack_jm_half(int uid0, int n, const double *J, const double *M, double *out, const int *slots14, cudaStream_t s);
void launch_species_moments(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *cellStart, double *spec, int nSM, cudaStream_t s);
void launch_correct_particle_location(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *nSlots, long long nUpper, const double *phi,
                                      const double *spec, const unsigned *neibMask, double qom0, int *cellCount, unsigned long long *counters,
                                      cudaStream_t s);
void launch_net_charge(const DevMesh &m, const DevSpecies &sp, ParticleSoA p, const int *cellStart, double chargeConv, double *rho, long long nUpper,
                       cudaStream_t s);
size_t sort_scan_tmp_bytes(long long nCells);
// migration record: 8 doubles (x,v,w,meta) + mu when the particles carry it
constexpr int AMPS_MIGRATION_RECORD_MAX = 10;  // doubles: the send / receive regions are sized for it
__host__ __device__ inline int migration_record_len(const ParticleSoA &p) { return 8 + (p.mu ? 1 : 0) + (p.vpar ? 1 : 0); }
void launch_pack_leavers(const DevMesh &m, ParticleSoA p, const int *nSlots, long long nUpper, const int *leafOwner, const int *leafGlobal, int me,
                         double *sendBuf, long long capPerPeer, int *sendCount, int *cellCount, int *errFlag, double *const *peerRecv, cudaStream_t s);
void launch_unpack_arrivals_peer(const DevMesh &m, const double *recvBuf, const int *allCounts, int R, long long capPerPeer, ParticleSoA p, int *nSlots,
                                 const int *g2l, const int *leafOwner, int me, long long capacity, int *cellCount, int *errFlag, long long *sentRecv,
                                 cudaStream_t s);
void launch_unpack_arrivals(const DevMesh &m, const double *recvBuf, int nRecv, ParticleSoA p, int *nSlots, const int *g2l, const int *leafOwner, int me,
                            long long capacity, int *cellCount, int *errFlag, cudaStream_t s);
void launch_pack_corners(const int *uids, int n, const double *J, const double *M, double *buf, cudaStream_t s);
void launch_add_corners_atomic(const int *uids, int n, double *J, double *M, const double *buf, cudaStream_t s);
void launch_add_corners(const int *uids, int n, double *J, double *M, const double *buf, cudaStream_t s);
void launch_stage_background(const DevMesh &m, const double *E, const double *B, double *tile, cudaStream_t s);
void launch_move_relativistic_boris(const DevMesh &m, const DevSpecies &sp, int interp, int backward, double c, double rSphere, long long exitCap,
                                    ParticleSoA p, const int *nSlots, long long nUpper, const double *bgTile, const double *uE, const double *uB,
                                    int *cellCount, DevMoveStats *stats, amps_gpu_exit_record *exitBuf, unsigned long long *exitCount,
                                    cudaStream_t s);
void launch_move_boris(const DevMesh &m, const DevSpecies &sp, bool markidis, int interp, int backward, double c, double rSphere, long long exitCap,
                       double gravityGM, ParticleSoA p, const int *nSlots, long long nUpper, const double *bgTile, const double *uE, const double *uB, int *cellCount,
                       DevMoveStats *stats, amps_gpu_exit_record *exitBuf, unsigned long long *exitCount, cudaStream_t s);
void launch_stage_center_table(const DevMesh &m, int nVar, const double *var, double *tile, cudaStream_t s);
void launch_gc_magnetic_moment_init(const DevMesh &m, const DevSpecies &sp, int interp, ParticleSoA p, const int *nSlots, long long nUpper,
                                    const double *bgTile, const double *uE, const double *uB, DevMoveStats *stats, cudaStream_t s, const double *ecsimE = nullptr, const double *ecsimB = nullptr);
void launch_move_guiding_center(const DevMesh &m, const DevSpecies &sp, int order, int interp, int idealMhd, double rSphere, long long exitCap,
                                ParticleSoA p, const int *nSlots, long long nUpper, const double *bgTile, const double *gradBTile, const double *uE,
                                const double *uB, const double *uGradB, int *cellCount, DevMoveStats *stats, amps_gpu_exit_record *exitBuf,
                                unsigned long long *exitCount, cudaStream_t s, const double *ecsimE = nullptr, const double *ecsimB = nullptr);
void launch_magnetic_moment_init(const DevMesh &m, const DevSpecies &sp, int interp, double c, ParticleSoA p, const int *nSlots, long long nUpper,
                                 const double *bgTile, const double *uE, const double *uB, DevMoveStats *stats, cudaStream_t s);
void launch_magnetic_moment_set(ParticleSoA p, double *target, const int *nSlots, long long nUpper, const double *muByPtr, long long nMu,
                                cudaStream_t s);
void launch_move_relativistic_gca(const DevMesh &m, const DevSpecies &sp, int interp, double c, double rSphere, long long exitCap, ParticleSoA p,
                                  const int *nSlots, long long nUpper, const double *bgTile, const double *gcaTile, const double *uE, const double *uB,
                                  const double *uVar, int *cellCount, DevMoveStats *stats, amps_gpu_exit_record *exitBuf,
                                  unsigned long long *exitCount, cudaStream_t s);
void launch_division_selftest(const double *a, const double *b, int n, unsigned long long *out, cudaStream_t s);

}  // namespace amps
