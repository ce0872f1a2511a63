// amps_gpu.cu -- host side of the C ABI declared in include/amps_gpu.h: context, device memory,
// uploads/downloads and kernel sequencing.  No CPU fallback anywhere: without a CUDA device
// amps_gpu_init() fails with AMPS_GPU_ERR_NO_DEVICE.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <map>
#include <vector>

#include "amps_dev.cuh"

using namespace amps;

struct amps_gpu_ctx {
  amps_gpu_config cfg;
  std::string err;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  int nSM = 148;

  DevMesh dm;
  DevSpecies sp;
  bool meshReady = false, fieldsReady = false;
  std::vector<void *> meshAllocs;

  // fields
  double *d_Ehalf = nullptr, *d_Bprev = nullptr, *d_Bcur = nullptr;
  double *d_eTile = nullptr, *d_bPrevTile = nullptr, *d_bCurTile = nullptr;

  // field solve (row f1): E^n on the unique corners, node adjacency, Krylov space
  double *d_E = nullptr;
  int *d_fNb = nullptr, *d_fCc = nullptr, *d_fZc = nullptr;
  bool fieldSolverReady = false, eReady = false;
  double dxc0[3] = {1.0, 1.0, 1.0};  // cell size of leaf 0 (single-level meshes: of every leaf)
  double *d_krylov = nullptr;   // [(restart + 1) + 2][3 nCorners]: V_0.., w, x
  int krylovVectors = 0;
  int lastFieldIters = 0;       // iterations of the previous field step: that many Arnoldi steps (less one) run before the host looks
  double *d_hcol = nullptr, *h_hcol = nullptr;  // inner products of one iteration (device / pinned host), + the norm
  double *d_ycoef = nullptr;
  // several ranks: halo lists of the field solve (per peer: corners to send / receive, centres to send / receive, concatenated in
  // rank order) and the primary-owner mask of the inner products
  struct FieldHalo {
    std::vector<long long> cSendOff, cRecvOff, zSendOff, zRecvOff;  // [nRanks + 1] offsets into the concatenated lists
    int *d_cSend = nullptr, *d_cRecv = nullptr, *d_zSend = nullptr, *d_zRecv = nullptr;
    double *d_sendBuf = nullptr, *d_recvBuf = nullptr;
    long long bufEntries = 0;
    std::vector<std::vector<int>> h[4];  // per kind, per peer (host copies until the lists are concatenated)
    bool dirty = true, set = false;
  } fh;
  unsigned char *d_primary = nullptr;  // [nCorners]
  double *d_fieldK = nullptr;   // operator constants K[243] (+ the right-hand side kept for restarts behind the Krylov vectors)
  double fieldKtheta = -1.0;    // theta the constants were built for
  bool fieldWarmValid = false;  // the solution of the previous field step is still in the workspace
  int fieldWarmN = 0;

  // particles
  ParticleSoA buf[2];
  int cur = 0;
  int *d_n = nullptr;  // [2]
  long long nUpper = 0;  // host-side bound of the used slots of buf[cur] (grid sizes); re-tightened after every sort (h_nSorted)
  int *h_nSorted = nullptr;         // pinned: particle count written by the last sort (slots [0, count) are exactly the residents)
  cudaEvent_t evSorted = nullptr;   // the copy into h_nSorted has completed
  bool nSortedPending = false;
  std::vector<int> h_leafOwnerHost;   // [nLeaves] owner rank (several ranks only)
  std::vector<int32_t> h_uploadedSlots;  // ParticleBuffer slots handed over by the last upload / slot assignment (download bookkeeping)
  bool sorted = false, countValid = false;
  int *d_cellCount = nullptr, *d_cellStart = nullptr, *d_cellFill = nullptr;
  void *d_scanTmp = nullptr;
  long long nCells = 0;

  // deposit
  double *d_J = nullptr, *d_M = nullptr, *d_energy = nullptr;
  unsigned long long *d_cfl = nullptr;
  DevMoveStats *d_stats = nullptr;

  // coupler table of the test-particle movers + exit records
  double *d_bgE = nullptr, *d_bgB = nullptr, *d_bgTile = nullptr;
  double *d_gcaVar = nullptr, *d_gcaTile = nullptr;  // relativistic GCA: 15 drift variables per centre node
  bool gcaReady = false;
  bool meshRefined = false;  // some leaf is below level 0
  std::vector<int> h_leafNeib;   // [nLeaves] LeafGeo::neib (which blocks get a table of the coupler's stencil cache)
  std::vector<int> h_leafLevel;
  double *d_vnByPtr = nullptr;   // PB::GetVNormal by ParticleBuffer slot (guiding-centre species of the deposit)
  long long nVn = 0;
  bool cplrCacheTried = false;   // build_cplr_cache ran for this mesh (it may have declined: table too large / switched off)
  int *d_neib26 = nullptr, *d_mbSlot = nullptr, *d_mbLeaf = nullptr;
  unsigned char *d_mbTab = nullptr;
  unsigned char *d_redoMask = nullptr;  // particles the fast mover left to the exact kernel
  int *d_perm = nullptr;      // sorted position -> slot (fused sort + deposit inside amps_gpu_step)
  double *d_rho = nullptr;    // ComputeNetCharge: rho_new on the unique centre nodes
  double *d_pack = nullptr;                   // packed J + half mass matrix [nCorners][129] (amps_gpu_step_JM_packed)
  double *d_sample = nullptr;                 // PIC::Sampling collecting buffer [nCells][n_species][13]
  unsigned long long *d_nSampled = nullptr;   // particles sampled per species
  double *d_spec = nullptr;   // species moments on the unique corners [nCorners][n_species][10]
  double *d_phi = nullptr;    // div-E correction potential on the unique centre nodes
  unsigned *d_neibMask = nullptr;            // [nLeaves] bit (sx+1)+3(sy+1)+9(sz+1): no (in use) neighbour block in that direction
  unsigned long long *d_cplCount = nullptr;  // CorrectParticleLocation: displaced, deleted, errors
  bool specReady = false, phiReady = false;
  // amps_gpu_step_JM: the deposit runs in cell ranges; the corners whose last contributing cell lies in range k form the
  // uid runs dlRuns[dlRunStart[k] .. dlRunStart[k+1]) and are copied to the host while range k+1 is deposited
  struct DlRun { int uid0, n; };
  std::vector<int> dlCellEnd;      // [nChunks] end cell of each range
  std::vector<int> dlRunStart;     // [nChunks+1]
  std::vector<DlRun> dlRuns;
  cudaStream_t copyStream = nullptr;
  std::vector<cudaEvent_t> dlEvents;
  int *d_leafRedo = nullptr;  // [nLeaves] counts, then [nLeaves] list of flagged blocks, then 1 counter
  double *d_gradBVar = nullptr, *d_gradBTile = nullptr;  // guiding centre: grad B, 9 values per centre node
  bool gradBReady = false;
  bool backgroundReady = false;
  amps_gpu_exit_record *d_exitBuf = nullptr;
  unsigned long long *d_exitCount = nullptr;

  // domain decomposition / NCCL
  int rank = 0, nRanks = 1;
  int *d_leafOwner = nullptr, *d_leafGlobal = nullptr, *d_g2l = nullptr;
  ncclComm_t comm = nullptr;
  double *d_sendBuf = nullptr, *d_recvBuf = nullptr;
  long long capPerPeer = 0;
  int *d_sendCount = nullptr, *d_allCounts = nullptr, *d_errFlag = nullptr;
  // peer-memory migration: the receive buffers of the other ranks mapped through CUDA IPC; the leavers are written there directly
  bool peerMigrate = false;
  std::vector<void *> h_peerMapped;   // per rank: mapped base of its receive buffer (nullptr for this rank)
  double **d_peerRecv = nullptr;      // the same table on the device
  int *h_errLazy = nullptr;           // pinned: error flag of the last exchange, read back behind the next sort
  int pendingErr = 0;
  long long *d_sentRecv = nullptr;    // {sent, received} of the last peer exchange (device)
  std::vector<int *> d_sharedUid;      // per peer
  std::vector<long long> nShared;      // per peer
  double *d_cornerSend = nullptr, *d_cornerRecv = nullptr;
  long long cornerBufDoubles = 0;

  // phase profiling (CUDA events on ctx->stream around move / sort / deposit)
  bool profile = false;
  std::vector<cudaEvent_t> evPool;
  size_t evUsed = 0;
  std::vector<std::pair<int, std::pair<size_t, size_t>>> evSpans;  // phase, (begin,end) event index
  // deposit order (DevMesh::depLeaf) and the overlap of the shared-corner exchange with the deposit of the interior leaves
  std::vector<int> h_leafCornerUid;           // [nLeaves][nCornerLocal] host copy (to find the leaves that touch shared corners)
  std::vector<char> h_leafGhost;              // [nLeaves] periodic ghost block: deposits nothing
  std::vector<int> h_realBefore;              // [nLeaves+1] depositing leaves with a smaller index (ascending order only)
  std::vector<std::vector<int>> h_sharedUid;  // per peer
  int *d_sharedUidAll = nullptr;              // the peers' lists concatenated in rank order (one pack / one add launch)
  long long nSharedAll = 0;
  bool sharedAllDirty = true;
  int *d_depLeaf = nullptr;
  int nDepBoundary = 0;                       // leading entries of depLeaf that touch a shared corner (0: order is ascending)
  bool depDirty = false;                      // shared corners changed since depLeaf was built
  cudaStream_t commStream = nullptr;
  cudaEvent_t evBoundary = nullptr, evRecv = nullptr;
  bool overlapJM = false;                     // the deposit just enqueued recorded evBoundary: the exchange may start there
  cudaEvent_t evCounts = nullptr;  // migration: the counts have reached the host
  bool jmZeroed = false;           // J, M were zeroed ahead of the deposit that follows
  std::map<int, double> subMs;                                      // debug: sub-phase id (>= 16) -> ms
  bool subPhases = getenv("AMPS_GPU_DEBUG_SUBPHASES") != nullptr;
  double phaseMs[AMPS_GPU_N_PHASES] = {0, 0, 0, 0};
  long long phaseCount[AMPS_GPU_N_PHASES] = {0, 0, 0, 0};
};

static size_t prof_mark(amps_gpu_ctx *ctx) {
  if (ctx->evUsed == ctx->evPool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->evPool.push_back(e);
  }
  cudaEventRecord(ctx->evPool[ctx->evUsed], ctx->stream);
  return ctx->evUsed++;
}
struct ProfScope {
  amps_gpu_ctx *ctx;
  int phase;
  size_t b;
  ProfScope(amps_gpu_ctx *c, int ph) : ctx(c), phase(ph), b(0) {
    if (ctx->profile) b = prof_mark(ctx);
  }
  ~ProfScope() {
    if (ctx->profile) {
      size_t e = prof_mark(ctx);
      ctx->evSpans.push_back({phase, {b, e}});
    }
  }
};

// debug (AMPS_GPU_DEBUG_SUBPHASES): spans inside the exchange phase, printed by amps_gpu_profile
struct Sub {
  amps_gpu_ctx *c;
  int id;
  size_t b = 0;
  bool on;
  Sub(amps_gpu_ctx *cc, int i) : c(cc), id(i), on(cc->profile && cc->subPhases) {
    if (on) b = prof_mark(c);
  }
  void end() {
    if (on) c->evSpans.push_back({id, {b, prof_mark(c)}}), on = false;
  }
};

#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) {                                                                           \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"; \
      return AMPS_GPU_ERR_CUDA;                                                                        \
    }                                                                                                  \
  } while (0)

#define FAIL(code, msg) \
  do {                  \
    ctx->err = (msg);   \
    return (code);      \
  } while (0)

template <class T>
static int dev_alloc(amps_gpu_ctx *ctx, T **p, size_t n) {
  CK(cudaMalloc((void **)p, n * sizeof(T) + 16));
  return AMPS_GPU_OK;
}

static int alloc_particles(amps_gpu_ctx *ctx, ParticleSoA &b, long long cap) {
  int rc;
  for (int d = 0; d < 3; d++) {
    if ((rc = dev_alloc(ctx, &b.x[d], cap))) return rc;
    if ((rc = dev_alloc(ctx, &b.v[d], cap))) return rc;
  }
  if ((rc = dev_alloc(ctx, &b.w, cap))) return rc;
  if ((rc = dev_alloc(ctx, &b.spec, cap))) return rc;
  if ((rc = dev_alloc(ctx, &b.key, cap))) return rc;
  if ((rc = dev_alloc(ctx, &b.ptr, cap))) return rc;
  b.mu = nullptr;
  if (ctx->cfg.carry_magnetic_moment) {
    if ((rc = dev_alloc(ctx, &b.mu, cap))) return rc;
    CK(cudaMemsetAsync(b.mu, 0, sizeof(double) * cap, ctx->stream));
  }
  b.vpar = nullptr;
  if (ctx->cfg.carry_v_parallel) {
    if ((rc = dev_alloc(ctx, &b.vpar, cap))) return rc;
    CK(cudaMemsetAsync(b.vpar, 0, sizeof(double) * cap, ctx->stream));
  }
  return AMPS_GPU_OK;
}
static void free_particles(ParticleSoA &b) {
  for (int d = 0; d < 3; d++) cudaFree(b.x[d]), cudaFree(b.v[d]);
  cudaFree(b.w), cudaFree(b.spec), cudaFree(b.key), cudaFree(b.ptr), cudaFree(b.mu), cudaFree(b.vpar);
}

template <class T>
static int upload_array(amps_gpu_ctx *ctx, const T **dst, const T *src, size_t n) {
  T *d = nullptr;
  CK(cudaMalloc((void **)&d, (n ? n : 1) * sizeof(T)));
  ctx->meshAllocs.push_back(d);
  if (n) CK(cudaMemcpyAsync(d, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  *dst = d;
  return AMPS_GPU_OK;
}

// ---- NCCL resolved at run time: the library has no link-time dependency on libnccl, and inside a process that
//      already loaded one (torch's bundled copy) the same instance is reused ----
struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
static NcclApi &nccl_api() {
  static NcclApi a;
  if (a.h || a.ok) return a;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.h) break;
  }
  if (!a.h) return a;
#define LOADSYM(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.h, sym))
  LOADSYM(GetUniqueId, "ncclGetUniqueId");
  LOADSYM(CommInitRank, "ncclCommInitRank");
  LOADSYM(CommDestroy, "ncclCommDestroy");
  LOADSYM(GroupStart, "ncclGroupStart");
  LOADSYM(GroupEnd, "ncclGroupEnd");
  LOADSYM(Send, "ncclSend");
  LOADSYM(Recv, "ncclRecv");
  LOADSYM(AllGather, "ncclAllGather");
  LOADSYM(AllReduce, "ncclAllReduce");
  LOADSYM(GetErrorString, "ncclGetErrorString");
#undef LOADSYM
  a.ok = a.GetUniqueId && a.CommInitRank && a.GroupStart && a.GroupEnd && a.Send && a.Recv && a.AllGather && a.AllReduce;
  return a;
}
#define NCK(call)                                                                                            \
  do {                                                                                                       \
    ncclResult_t r_ = (call);                                                                                \
    if (r_ != ncclSuccess) {                                                                                 \
      ctx->err = std::string(#call) + ": " + (nccl_api().GetErrorString ? nccl_api().GetErrorString(r_) : "nccl error"); \
      return AMPS_GPU_ERR_CUDA;                                                                              \
    }                                                                                                        \
  } while (0)

extern "C" {

int amps_gpu_init(const amps_gpu_config *cfg, amps_gpu_ctx **out) {
  if (!cfg || !out) return AMPS_GPU_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    fprintf(stderr, "amps_gpu_init: no CUDA device visible; this library has no CPU fallback\n");
    return AMPS_GPU_ERR_NO_DEVICE;
  }
  if (cfg->device < 0 || cfg->device >= ndev) return AMPS_GPU_ERR_ARG;
  if (cfg->n_species < 1 || cfg->n_species > AMPS_GPU_MAX_SPECIES) return AMPS_GPU_ERR_ARG;
  if (cfg->capacity < 1 || cfg->capacity > 2147483000LL) return AMPS_GPU_ERR_ARG;
  for (int d = 0; d < 3; d++)
    if (cfg->block_cells[d] < 1 || cfg->ghost_cells[d] < 1) return AMPS_GPU_ERR_ARG;
  if (cfg->b_mode != AMPS_B_CENTER_BASED && cfg->b_mode != AMPS_B_CORNER_BASED) return AMPS_GPU_ERR_ARG;
  amps_gpu_ctx *ctx = new amps_gpu_ctx();
  ctx->cfg = *cfg;
  memset(&ctx->dm, 0, sizeof(ctx->dm));
  memset(&ctx->sp, 0, sizeof(ctx->sp));
  memset(ctx->buf, 0, sizeof(ctx->buf));
  *out = ctx;  // returned even on failure so that the caller can read last_error
  CK(cudaSetDevice(cfg->device));
  CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CK(cudaDeviceGetAttribute(&ctx->nSM, cudaDevAttrMultiProcessorCount, cfg->device));

  DevSpecies &sp = ctx->sp;
  sp.n = cfg->n_species;
  sp.timeStepMode = cfg->time_step_mode;
  sp.bMode = cfg->b_mode;
  sp.boundaryMode = cfg->boundary_mode;
  for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++) {
    sp.charge[s] = cfg->charge[s], sp.mass[s] = cfg->mass[s], sp.weight[s] = cfg->species_weight[s], sp.dt[s] = cfg->time_step[s];
  }
  sp.dtTotal = cfg->ecsim_dt_total, sp.B_conv = cfg->ecsim_B_conv, sp.length_conv = cfg->ecsim_length_conv, sp.LightSpeed = cfg->ecsim_light_speed;

  int rc;
  for (int b = 0; b < 2; b++)
    if ((rc = alloc_particles(ctx, ctx->buf[b], cfg->capacity))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_n, 2))) return rc;
  CK(cudaMemset(ctx->d_n, 0, 2 * sizeof(int)));
  if ((rc = dev_alloc(ctx, &ctx->d_energy, 1))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_cfl, AMPS_GPU_MAX_SPECIES))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_stats, 1))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_exitCount, 1))) return rc;
  CK(cudaMemset(ctx->d_exitCount, 0, sizeof(unsigned long long)));
  if (cfg->exit_record_capacity > 0)
    if ((rc = dev_alloc(ctx, &ctx->d_exitBuf, (size_t)cfg->exit_record_capacity))) return rc;
  return AMPS_GPU_OK;
}

int amps_gpu_finalize(amps_gpu_ctx *ctx) {
  if (!ctx) return AMPS_GPU_OK;
  cudaSetDevice(ctx->cfg.device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (void *p : ctx->meshAllocs) cudaFree(p);
  for (cudaEvent_t e : ctx->evPool) cudaEventDestroy(e);
  if (ctx->comm && nccl_api().CommDestroy) nccl_api().CommDestroy(ctx->comm);
  cudaFree(ctx->d_gcaVar), cudaFree(ctx->d_gcaTile), cudaFree(ctx->d_gradBVar), cudaFree(ctx->d_gradBTile);
  cudaFree(ctx->d_redoMask), cudaFree(ctx->d_leafRedo), cudaFree(ctx->d_perm), cudaFree(ctx->d_rho);
  for (cudaEvent_t e : ctx->dlEvents) cudaEventDestroy(e);
  if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
  if (ctx->evCounts) cudaEventDestroy(ctx->evCounts);
  cudaFree(ctx->d_E), cudaFree(ctx->d_fNb), cudaFree(ctx->d_fCc), cudaFree(ctx->d_fZc), cudaFree(ctx->d_krylov), cudaFree(ctx->d_hcol), cudaFree(ctx->d_ycoef), cudaFree(ctx->d_fieldK);
  if (ctx->h_hcol) cudaFreeHost(ctx->h_hcol);
  for (void *q : ctx->h_peerMapped)
    if (q) cudaIpcCloseMemHandle(q);
  cudaFree(ctx->d_peerRecv), cudaFree(ctx->d_sentRecv);
  cudaFree(ctx->d_vnByPtr), cudaFree(ctx->d_neib26), cudaFree(ctx->d_mbSlot), cudaFree(ctx->d_mbLeaf), cudaFree(ctx->d_mbTab);
  if (ctx->h_errLazy) cudaFreeHost(ctx->h_errLazy);
  if (ctx->evSorted) cudaEventDestroy(ctx->evSorted);
  if (ctx->h_nSorted) cudaFreeHost(ctx->h_nSorted);
  cudaFree(ctx->d_spec), cudaFree(ctx->d_phi), cudaFree(ctx->d_cplCount), cudaFree(ctx->d_sample), cudaFree(ctx->d_nSampled), cudaFree(ctx->d_pack);
  if (ctx->evBoundary) cudaEventDestroy(ctx->evBoundary);
  if (ctx->evRecv) cudaEventDestroy(ctx->evRecv);
  if (ctx->commStream) cudaStreamDestroy(ctx->commStream);
  cudaFree(ctx->d_bgE), cudaFree(ctx->d_bgB), cudaFree(ctx->d_bgTile), cudaFree(ctx->d_exitBuf), cudaFree(ctx->d_exitCount);
  cudaFree(ctx->d_sendBuf), cudaFree(ctx->d_recvBuf), cudaFree(ctx->d_sendCount), cudaFree(ctx->d_allCounts), cudaFree(ctx->d_errFlag);
  for (int *p : ctx->d_sharedUid) cudaFree(p);
  cudaFree(ctx->d_cornerSend), cudaFree(ctx->d_cornerRecv), cudaFree(ctx->d_sharedUidAll);
  cudaFree(ctx->d_Ehalf), cudaFree(ctx->d_Bprev), cudaFree(ctx->d_Bcur);
  cudaFree(ctx->d_eTile), cudaFree(ctx->d_bPrevTile), cudaFree(ctx->d_bCurTile);
  for (int b = 0; b < 2; b++) free_particles(ctx->buf[b]);
  cudaFree(ctx->d_n), cudaFree(ctx->d_cellCount), cudaFree(ctx->d_cellStart), cudaFree(ctx->d_cellFill), cudaFree(ctx->d_scanTmp);
  cudaFree(ctx->d_J), cudaFree(ctx->d_M), cudaFree(ctx->d_energy), cudaFree(ctx->d_cfl), cudaFree(ctx->d_stats);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return AMPS_GPU_OK;
}

const char *amps_gpu_last_error(const amps_gpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t amps_gpu_launch_count(const amps_gpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int amps_gpu_last_move_redo(amps_gpu_ctx *ctx, int64_t *n) {
  if (!ctx || !n) return AMPS_GPU_ERR_ARG;
  *n = 0;
  if (!ctx->d_leafRedo || ctx->cfg.exact_arithmetic) return AMPS_GPU_OK;
  CK(cudaSetDevice(ctx->cfg.device));
  std::vector<int> h((size_t)ctx->dm.nLeaves);
  CK(cudaMemcpyAsync(h.data(), ctx->d_leafRedo, sizeof(int) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int v : h) *n += v;
  return AMPS_GPU_OK;
}
void *amps_gpu_stream(amps_gpu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int amps_gpu_synchronize(amps_gpu_ctx *ctx) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  CK(cudaStreamSynchronize(ctx->stream));
  return AMPS_GPU_OK;
}

// frees what amps_gpu_mesh_upload and the per-mesh uploads allocated (see there)
static int release_mesh(amps_gpu_ctx *ctx) {
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->copyStream) CK(cudaStreamSynchronize(ctx->copyStream));
  if (ctx->commStream) CK(cudaStreamSynchronize(ctx->commStream));
  for (void *p : ctx->meshAllocs) cudaFree(p);
  ctx->meshAllocs.clear();
  auto drop = [](auto *&p) {
    cudaFree(p);
    p = nullptr;
  };
  drop(ctx->d_gcaVar), drop(ctx->d_gcaTile), drop(ctx->d_gradBVar), drop(ctx->d_gradBTile);
  drop(ctx->d_leafRedo), drop(ctx->d_rho), drop(ctx->d_spec), drop(ctx->d_phi), drop(ctx->d_sample), drop(ctx->d_nSampled), drop(ctx->d_pack);
  drop(ctx->d_bgE), drop(ctx->d_bgB), drop(ctx->d_bgTile);
  drop(ctx->d_neib26), drop(ctx->d_mbSlot), drop(ctx->d_mbLeaf), drop(ctx->d_mbTab);
  ctx->cplrCacheTried = false;
  // the field solve is sized by the mesh as well: a new epoch needs amps_gpu_field_solver_init (and the halo / primary lists) again
  drop(ctx->d_E), drop(ctx->d_fNb), drop(ctx->d_fCc), drop(ctx->d_fZc), drop(ctx->d_krylov), drop(ctx->d_primary);
  ctx->fieldSolverReady = ctx->eReady = false, ctx->fieldWarmValid = false, ctx->krylovVectors = 0, ctx->lastFieldIters = 0;
  drop(ctx->d_sendBuf), drop(ctx->d_recvBuf), drop(ctx->d_sendCount), drop(ctx->d_allCounts), drop(ctx->d_errFlag);
  for (int *&p : ctx->d_sharedUid) drop(p);
  ctx->d_sharedUid.clear(), ctx->nShared.clear(), ctx->h_sharedUid.clear();
  drop(ctx->d_sharedUidAll);
  ctx->nSharedAll = 0, ctx->sharedAllDirty = true;
  drop(ctx->d_cornerSend), drop(ctx->d_cornerRecv);
  ctx->cornerBufDoubles = 0;
  drop(ctx->d_Ehalf), drop(ctx->d_Bprev), drop(ctx->d_Bcur), drop(ctx->d_eTile), drop(ctx->d_bPrevTile), drop(ctx->d_bCurTile);
  drop(ctx->d_cellCount), drop(ctx->d_cellStart), drop(ctx->d_cellFill), drop(ctx->d_scanTmp), drop(ctx->d_J), drop(ctx->d_M);
  ctx->d_leafOwner = ctx->d_leafGlobal = ctx->d_g2l = nullptr;  // (were in meshAllocs)
  ctx->d_depLeaf = nullptr, ctx->d_neibMask = nullptr;
  ctx->h_leafCornerUid.clear(), ctx->h_leafGhost.clear(), ctx->h_realBefore.clear();
  ctx->dlCellEnd.clear(), ctx->dlRunStart.clear(), ctx->dlRuns.clear();
  ctx->meshReady = ctx->fieldsReady = ctx->gcaReady = ctx->gradBReady = ctx->backgroundReady = ctx->specReady = ctx->phiReady = false;
  ctx->meshRefined = false;
  ctx->nDepBoundary = 0, ctx->depDirty = false, ctx->overlapJM = false, ctx->jmZeroed = false;
  CK(cudaMemsetAsync(ctx->d_n, 0, 2 * sizeof(int), ctx->stream));
  ctx->cur = 0, ctx->nUpper = 0, ctx->sorted = false, ctx->countValid = false;
  memset(&ctx->dm, 0, sizeof(ctx->dm));
  return AMPS_GPU_OK;
}

int amps_gpu_mesh_upload(amps_gpu_ctx *ctx, const amps_gpu_mesh *mesh) {
  if (!ctx || !mesh) return AMPS_GPU_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  if (ctx->meshReady) {
    // a new mesh epoch (the reference rebuilds BlockTable when nMeshModificationCounter changes, pic_mesh.cpp:1658): everything
    // sized by the old mesh goes, and so do the resident particles and fields - their (block,cell) keys and node ids named the
    // old blocks; the host uploads them again.  Communicator, capacity and species tables stay.
    int rcRelease;
    if ((rcRelease = release_mesh(ctx))) return rcRelease;
  }
  if (mesh->n_nodes < 1 || mesh->n_leaves < 1 || mesh->n_corners < 1 || mesh->n_centers < 1) FAIL(AMPS_GPU_ERR_ARG, "empty mesh");
  DevMesh &m = ctx->dm;
  for (int d = 0; d < 3; d++) {
    m.N[d] = ctx->cfg.block_cells[d], m.g[d] = ctx->cfg.ghost_cells[d], m.TN[d] = m.N[d] + 2 * m.g[d];
    m.nRoot[d] = mesh->n_root[d];
    m.xGlobalMin[d] = mesh->x_global_min[d], m.xGlobalMax[d] = mesh->x_global_max[d];
    m.dxMaxRef[d] = mesh->dx_max_refinement[d], m.dxRoot[d] = mesh->dx_root_block[d];
  }
  m.L = mesh->max_refinement_level;
  m.eps = mesh->eps;
  m.nNodes = mesh->n_nodes, m.nLeaves = mesh->n_leaves, m.nCorners = mesh->n_corners, m.nCenters = mesh->n_centers;
  m.cellsPerBlock = m.N[0] * m.N[1] * m.N[2];
  m.nCornerLocal = (m.TN[0] + 1) * (m.TN[1] + 1) * (m.TN[2] + 1);
  m.nCenterLocal = m.TN[0] * m.TN[1] * m.TN[2];
  m.eTileStride = (3 * m.nCornerLocal + 1) & ~1;
  m.bTileStride = (ctx->cfg.b_mode == AMPS_B_CORNER_BASED) ? m.eTileStride : ((3 * m.nCenterLocal + 1) & ~1);
  m.periodic = ctx->cfg.periodic;
  const long long nCells = (long long)m.nLeaves * m.cellsPerBlock;
  if (nCells > 2147483000LL) FAIL(AMPS_GPU_ERR_ARG, "too many cells for 32-bit keys");
  ctx->nCells = nCells;

  // per-leaf geometry
  std::vector<LeafGeo> lg(m.nLeaves);
  ctx->h_leafNeib.assign((size_t)m.nLeaves, 0), ctx->h_leafLevel.assign((size_t)m.nLeaves, 0);
  ctx->cplrCacheTried = false;
  for (int l = 0; l < m.nLeaves; l++) {
    const int n = mesh->leaf_node[l];
    if (n < 0 || n >= m.nNodes) FAIL(AMPS_GPU_ERR_ARG, "leaf_node out of range");
    LeafGeo &g = lg[l];
    for (int d = 0; d < 3; d++) {
      g.xmin[d] = mesh->node_xmin[3 * n + d], g.xmax[d] = mesh->node_xmax[3 * n + d];
      g.imin[d] = mesh->node_imin[3 * n + d];
    }
    g.isize = mesh->node_isize[n];
    g.level = mesh->node_level[n];
    if (g.level > 0) ctx->meshRefined = true;
    g.flags = mesh->node_flags[n];
    g.real = mesh->leaf_real[l];
    g.face = mesh->leaf_face_boundary[l];
    g.node = n;
    {
      // SetNeibRefinmentLevelLimits (meshAMRgeneric.h:1018-1048): levels of the 24 face, 8 corner and 24 edge neighbours
      // found by lattice probes (neibNodeFace/Corner/Edge, :505-725)
      int mn = -1, mx = -1;
      auto probe = [&](int ix0, int ix1, int ix2) {
        if (ix0 < 0 || ix1 < 0 || ix2 < 0) return;
        const int r0 = ix0 >> m.L, r1 = ix1 >> m.L, r2 = ix2 >> m.L;
        if (r0 >= m.nRoot[0] || r1 >= m.nRoot[1] || r2 >= m.nRoot[2]) return;
        int q = mesh->root_node[r0 + m.nRoot[0] * (r1 + m.nRoot[1] * r2)];
        while (true) {
          const int h = mesh->node_isize[q] / 2;
          const int i = (ix0 - mesh->node_imin[3 * q] < h) ? 0 : 1, j = (ix1 - mesh->node_imin[3 * q + 1] < h) ? 0 : 1,
                    k = (ix2 - mesh->node_imin[3 * q + 2] < h) ? 0 : 1;
          const int t = mesh->node_child[8 * q + i + 2 * (j + 2 * k)];
          if (t < 0) break;
          q = t;
        }
        const int lv = mesh->node_level[q];
        if (mn == -1 || mn > lv) mn = lv;
        if (mx < lv) mx = lv;
      };
      const int S = g.isize, h = (S > 1) ? S / 2 : 0;
      // offsets along one direction: -1 (beyond the low side), S (beyond the high side), 0 and S/2 (the two halves inside)
      const int out[2] = {-1, S}, in[2] = {0, h};
      for (int dn = 0; dn < 3; dn++) {  // faces: normal dn, 2 sides x 2 x 2 sub-faces
        const int t0 = (dn + 1) % 3, t1 = (dn + 2) % 3;
        for (int sd = 0; sd < 2; sd++)
          for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++) {
              int ix[3];
              ix[dn] = g.imin[dn] + out[sd], ix[t0] = g.imin[t0] + in[a], ix[t1] = g.imin[t1] + in[b];
              probe(ix[0], ix[1], ix[2]);
            }
      }
      for (int c = 0; c < 8; c++) probe(g.imin[0] + out[c & 1], g.imin[1] + out[(c >> 1) & 1], g.imin[2] + out[(c >> 2) & 1]);
      for (int de = 0; de < 3; de++) {  // edges along de: 4 transverse corners x 2 segments
        const int t0 = (de + 1) % 3, t1 = (de + 2) % 3;
        for (int a = 0; a < 2; a++)
          for (int b = 0; b < 2; b++)
            for (int sg = 0; sg < 2; sg++) {
              int ix[3];
              ix[de] = g.imin[de] + in[sg], ix[t0] = g.imin[t0] + out[a], ix[t1] = g.imin[t1] + out[b];
              probe(ix[0], ix[1], ix[2]);
            }
      }
      g.neib = (mn & 0xffff) | ((mx & 0xffff) << 16);
      ctx->h_leafNeib[l] = g.neib, ctx->h_leafLevel[l] = g.level;
    }
    {
      double vol = 1, d2 = 0;
      for (int d = 0; d < 3; d++) {
        g.dxc[d] = (g.xmax[d] - g.xmin[d]) / m.N[d];
        if (l == 0) ctx->dxc0[d] = g.dxc[d];
        g.invdxc[d] = 1.0 / g.dxc[d];
        const double dxl = g.dxc[d] * ctx->cfg.ecsim_length_conv;
        vol *= dxl;
        d2 += dxl * dxl;
      }
      g.invV = 1.0 / vol;
      g.diag = sqrt(d2);
    }
  }
  // ---- download schedule of amps_gpu_step_JM: for every corner the last cell range that deposits into it ----
  {
    // ranges of leaves: whole z-layers of root blocks on an unrefined mesh (leaves and corner ids are both z-major there, so
    // the corners finished by a range are one contiguous run), equal shares of the leaf list otherwise
    std::vector<int> leafEnd;
    const int perLayer = m.nRoot[0] * m.nRoot[1];
    if (!ctx->meshRefined && m.nLeaves == perLayer * m.nRoot[2] && m.nRoot[2] >= 4) {
      const int group = (m.nRoot[2] + 9) / 10;  // at most ~10 ranges
      for (int z = group; z < m.nRoot[2]; z += group) leafEnd.push_back(z * perLayer);
      leafEnd.push_back(m.nLeaves);
    } else {
      const int nEq = (m.nLeaves >= 64) ? 8 : 1;
      for (int k = 0; k < nEq; k++) leafEnd.push_back((int)((long long)m.nLeaves * (k + 1) / nEq));
    }
    const int nChunks = (int)leafEnd.size();
    std::vector<int> lastChunk((size_t)m.nCorners, -1);
    ctx->dlCellEnd.assign(nChunks, 0);
    for (int k = 0; k < nChunks; k++) {
      const int l0 = k ? leafEnd[k - 1] : 0, l1 = leafEnd[k];
      ctx->dlCellEnd[k] = l1 * m.cellsPerBlock;
      for (int l = l0; l < l1; l++) {
        if (ctx->cfg.periodic && mesh->leaf_face_boundary[l] != 0) continue;  // periodic ghost blocks deposit nothing
        const int *cu = mesh->leaf_corner_uid + (size_t)l * m.nCornerLocal;
        for (int kk = 0; kk <= m.N[2]; kk++)
          for (int jj = 0; jj <= m.N[1]; jj++)
            for (int ii = 0; ii <= m.N[0]; ii++) {
              const int u = cu[ii + m.g[0] + (m.TN[0] + 1) * (jj + m.g[1] + (kk + m.g[2]) * (m.TN[1] + 1))];
              if (u >= 0) lastChunk[u] = k;  // ranges are visited in order: the last writer wins
            }
      }
    }
    ctx->dlRunStart.assign(nChunks + 1, 0);
    ctx->dlRuns.clear();
    for (int k = 0; k < nChunks; k++) {
      ctx->dlRunStart[k] = (int)ctx->dlRuns.size();
      for (int u = 0; u < m.nCorners;) {
        // corners no cell deposits into (ghost-layer nodes) stay zero: they travel with range 0
        const int ck = lastChunk[u] < 0 ? 0 : lastChunk[u];
        if (ck != k) {
          u++;
          continue;
        }
        int v = u + 1;
        while (v < m.nCorners && (lastChunk[v] < 0 ? 0 : lastChunk[v]) == k) v++;
        // runs closer than 2048 corners are merged: a few corners travel twice (their final copy comes later on the
        // same stream), but a range costs two API calls instead of hundreds
        const int first = ctx->dlRunStart[k];
        if ((int)ctx->dlRuns.size() > first && u - (ctx->dlRuns.back().uid0 + ctx->dlRuns.back().n) < 256)
          ctx->dlRuns.back().n = v - ctx->dlRuns.back().uid0;
        else
          ctx->dlRuns.push_back({u, v - u});
        u = v;
      }
    }
    ctx->dlRunStart[nChunks] = (int)ctx->dlRuns.size();
    if (ctx->dlRuns.size() > 4096) {  // scattered ownership (e.g. unordered AMR leaves): one range, one run
      ctx->dlCellEnd.assign(1, m.nLeaves * m.cellsPerBlock);
      ctx->dlRunStart = {0, 1};
      ctx->dlRuns = {{0, m.nCorners}};
    }
  }
  int rc;
  const int nRootTot = m.nRoot[0] * m.nRoot[1] * m.nRoot[2];
  if ((rc = upload_array(ctx, &m.child, mesh->node_child, (size_t)8 * m.nNodes))) return rc;
  if ((rc = upload_array(ctx, &m.imin, mesh->node_imin, (size_t)3 * m.nNodes))) return rc;
  if ((rc = upload_array(ctx, &m.isize, mesh->node_isize, (size_t)m.nNodes))) return rc;
  if ((rc = upload_array(ctx, &m.nodeLeaf, mesh->node_leaf, (size_t)m.nNodes))) return rc;
  if ((rc = upload_array(ctx, &m.nodeFlags, mesh->node_flags, (size_t)m.nNodes))) return rc;
  if ((rc = upload_array(ctx, &m.nodeLevel, mesh->node_level, (size_t)m.nNodes))) return rc;
  if ((rc = upload_array(ctx, &m.nxmin, mesh->node_xmin, (size_t)3 * m.nNodes))) return rc;
  if ((rc = upload_array(ctx, &m.nxmax, mesh->node_xmax, (size_t)3 * m.nNodes))) return rc;
  if ((rc = upload_array(ctx, &m.rootNode, mesh->root_node, (size_t)nRootTot))) return rc;
  if ((rc = upload_array(ctx, &m.leaf, lg.data(), (size_t)m.nLeaves))) return rc;
  if ((rc = upload_array(ctx, &m.cornerUid, mesh->leaf_corner_uid, (size_t)m.nLeaves * m.nCornerLocal))) return rc;
  if ((rc = upload_array(ctx, &m.centerUid, mesh->leaf_center_uid, (size_t)m.nLeaves * m.nCenterLocal))) return rc;
  {
    ctx->h_leafGhost.assign((size_t)m.nLeaves, 0);
    ctx->h_realBefore.assign((size_t)m.nLeaves + 1, 0);
    std::vector<int> depLeaf, ghostLeaf;
    for (int l = 0; l < m.nLeaves; l++) {
      const bool ghost = ctx->cfg.periodic && mesh->leaf_face_boundary[l] != 0;
      ctx->h_leafGhost[l] = ghost;
      (ghost ? ghostLeaf : depLeaf).push_back(l);
      ctx->h_realBefore[l + 1] = (int)depLeaf.size();
    }
    m.nDepReal = (int)depLeaf.size();
    depLeaf.insert(depLeaf.end(), ghostLeaf.begin(), ghostLeaf.end());
    if ((rc = upload_array(ctx, &m.depLeaf, depLeaf.data(), depLeaf.size()))) return rc;
    ctx->d_depLeaf = const_cast<int *>(m.depLeaf);
    ctx->nDepBoundary = 0, ctx->depDirty = false;
    // isBoundaryCell (pic_field_solver_ecsim.cpp:6963-6999) asks for the face / edge / corner neighbour blocks: on a single-level
    // mesh they are the adjacent root blocks (refined meshes: CorrectParticleLocation is rejected like ComputeNetCharge)
    std::vector<unsigned> neibMask((size_t)m.nLeaves, 0u);
    if (!ctx->meshRefined)
      for (int l = 0; l < m.nLeaves; l++) {
        const int n = mesh->leaf_node[l];
        int r[3];
        for (int d = 0; d < 3; d++) r[d] = mesh->node_imin[3 * n + d] >> m.L;
        for (int q = 0; q < 27; q++) {
          if (q == 13) continue;
          const int s[3] = {q % 3 - 1, (q / 3) % 3 - 1, q / 9 - 1};
          bool bad = false;
          int rr[3];
          for (int d = 0; d < 3; d++) {
            rr[d] = r[d] + s[d];
            if (rr[d] < 0 || rr[d] >= m.nRoot[d]) bad = true;
          }
          if (!bad) {
            const int nb = mesh->root_node[rr[0] + m.nRoot[0] * (rr[1] + m.nRoot[1] * rr[2])];
            bad = nb < 0 || !(mesh->node_flags[nb] & AMPS_NODE_USED);
          }
          if (bad) neibMask[l] |= 1u << q;
        }
      }
    const unsigned *dmask;
    if ((rc = upload_array(ctx, &dmask, neibMask.data(), neibMask.size()))) return rc;
    ctx->d_neibMask = const_cast<unsigned *>(dmask);
    if (mesh->n_ranks > 1) ctx->h_leafCornerUid.assign(mesh->leaf_corner_uid, mesh->leaf_corner_uid + (size_t)m.nLeaves * m.nCornerLocal);
    CK(cudaStreamSynchronize(ctx->stream));  // depLeaf is a local
  }
  ctx->rank = (mesh->n_ranks > 1) ? mesh->this_rank : 0;
  ctx->nRanks = (mesh->n_ranks > 1) ? mesh->n_ranks : 1;
  if (ctx->nRanks > 1) {
    if (!mesh->leaf_owner || !mesh->leaf_global_id || !mesh->global_leaf_to_local || mesh->n_global_leaves < 1)
      FAIL(AMPS_GPU_ERR_ARG, "n_ranks > 1 needs leaf_owner, leaf_global_id and global_leaf_to_local");
    const int *t1, *t2, *t3;
    if ((rc = upload_array(ctx, &t1, mesh->leaf_owner, (size_t)m.nLeaves))) return rc;
    if ((rc = upload_array(ctx, &t2, mesh->leaf_global_id, (size_t)m.nLeaves))) return rc;
    if ((rc = upload_array(ctx, &t3, mesh->global_leaf_to_local, (size_t)mesh->n_global_leaves))) return rc;
    ctx->d_leafOwner = const_cast<int *>(t1), ctx->d_leafGlobal = const_cast<int *>(t2), ctx->d_g2l = const_cast<int *>(t3);
    ctx->h_leafOwnerHost.assign(mesh->leaf_owner, mesh->leaf_owner + m.nLeaves);
    ctx->capPerPeer = ctx->cfg.capacity / 32 > 65536 ? ctx->cfg.capacity / 32 : 65536;
    // sized for the longest record (x, v, w, key|species + magnetic moment + v_parallel)
    if ((rc = dev_alloc(ctx, &ctx->d_sendBuf, (size_t)ctx->nRanks * ctx->capPerPeer * AMPS_MIGRATION_RECORD_MAX))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_recvBuf, (size_t)ctx->nRanks * ctx->capPerPeer * AMPS_MIGRATION_RECORD_MAX))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_sendCount, (size_t)ctx->nRanks))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_allCounts, (size_t)ctx->nRanks * ctx->nRanks))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_errFlag, 1))) return rc;
    CK(cudaMemsetAsync(ctx->d_errFlag, 0, sizeof(int), ctx->stream));
    ctx->d_sharedUid.assign(ctx->nRanks, nullptr);
    ctx->nShared.assign(ctx->nRanks, 0);
    ctx->h_sharedUid.assign(ctx->nRanks, std::vector<int>());
  }
  CK(cudaStreamSynchronize(ctx->stream));  // lg is a local

  if ((rc = dev_alloc(ctx, &ctx->d_Ehalf, (size_t)3 * m.nCorners))) return rc;
  const size_t nBNodes = (ctx->cfg.b_mode == AMPS_B_CORNER_BASED) ? m.nCorners : m.nCenters;
  if ((rc = dev_alloc(ctx, &ctx->d_Bprev, (size_t)3 * nBNodes))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_Bcur, (size_t)3 * nBNodes))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_eTile, (size_t)m.nLeaves * m.eTileStride))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_bPrevTile, (size_t)m.nLeaves * m.bTileStride))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_bCurTile, (size_t)m.nLeaves * m.bTileStride))) return rc;
  CK(cudaMemsetAsync(ctx->d_eTile, 0, sizeof(double) * (size_t)m.nLeaves * m.eTileStride, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_bPrevTile, 0, sizeof(double) * (size_t)m.nLeaves * m.bTileStride, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_bCurTile, 0, sizeof(double) * (size_t)m.nLeaves * m.bTileStride, ctx->stream));
  if ((rc = dev_alloc(ctx, &ctx->d_cellCount, (size_t)nCells + 1))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_cellStart, (size_t)nCells + 1))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_cellFill, (size_t)nCells + 1))) return rc;
  CK(cudaMemsetAsync(ctx->d_cellStart, 0, sizeof(int) * ((size_t)nCells + 1), ctx->stream));
  CK(cudaMalloc(&ctx->d_scanTmp, sort_scan_tmp_bytes(nCells)));
  if ((rc = dev_alloc(ctx, &ctx->d_J, (size_t)3 * m.nCorners))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_M, (size_t)243 * m.nCorners))) return rc;
  ctx->meshReady = true;
  return AMPS_GPU_OK;
}

int amps_gpu_fields_upload(amps_gpu_ctx *ctx, const double *E_half, const double *B_prev, const double *B_cur) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "fields_upload before mesh_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  const DevMesh &m = ctx->dm;
  if (E_half) CK(cudaMemcpyAsync(ctx->d_Ehalf, E_half, sizeof(double) * 3 * (size_t)m.nCorners, cudaMemcpyHostToDevice, ctx->stream));
  const size_t nBNodes = (ctx->cfg.b_mode == AMPS_B_CORNER_BASED) ? m.nCorners : m.nCenters;
  if (B_prev) CK(cudaMemcpyAsync(ctx->d_Bprev, B_prev, sizeof(double) * 3 * nBNodes, cudaMemcpyHostToDevice, ctx->stream));
  if (B_cur) CK(cudaMemcpyAsync(ctx->d_Bcur, B_cur, sizeof(double) * 3 * nBNodes, cudaMemcpyHostToDevice, ctx->stream));
  launch_stage_tiles(m, ctx->cfg.b_mode == AMPS_B_CORNER_BASED, E_half ? ctx->d_Ehalf : nullptr, B_prev ? ctx->d_Bprev : nullptr, B_cur ? ctx->d_Bcur : nullptr, ctx->d_eTile,
                     ctx->d_bPrevTile, ctx->d_bCurTile, ctx->stream);
  ctx->launches++;
  CK(cudaGetLastError());
  ctx->fieldsReady = true;
  return AMPS_GPU_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// f1: the field half of the ECSIM step (ECSIM::TimeStep, pic_field_solver_ecsim.cpp:6004-6157) on the unique nodes; field_solver.cu
// ---------------------------------------------------------------------------------------------------------------------------
int amps_gpu_field_solver_init(amps_gpu_ctx *ctx, const int32_t *corner_nb, const int32_t *corner_cells, const int32_t *center_corners) {
  if (!ctx || !corner_nb || !corner_cells || !center_corners) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "field_solver_init before mesh_upload");
  if (ctx->meshRefined) FAIL(AMPS_GPU_ERR_STATE, "the device field solve covers single-level meshes (the compact 27-node rows of GetStencil)");
  if (ctx->cfg.b_mode != AMPS_B_CENTER_BASED) FAIL(AMPS_GPU_ERR_STATE, "the device field solve updates the centre-based B (UpdateB)");
  if (ctx->cfg.ecsim_B_conv != 1.0) FAIL(AMPS_GPU_ERR_STATE, "the device field solve assumes normalised units (E_conv = B_conv = 1)");
  CK(cudaSetDevice(ctx->cfg.device));
  const DevMesh &m = ctx->dm;
  int rc;
  cudaFree(ctx->d_fNb), cudaFree(ctx->d_fCc), cudaFree(ctx->d_fZc);
  if ((rc = dev_alloc(ctx, &ctx->d_fNb, (size_t)27 * m.nCorners))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_fCc, (size_t)8 * m.nCorners))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_fZc, (size_t)8 * m.nCenters))) return rc;
  CK(cudaMemcpy(ctx->d_fNb, corner_nb, sizeof(int) * 27 * (size_t)m.nCorners, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->d_fCc, corner_cells, sizeof(int) * 8 * (size_t)m.nCorners, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->d_fZc, center_corners, sizeof(int) * 8 * (size_t)m.nCenters, cudaMemcpyHostToDevice));
  if (!ctx->d_E) {
    if ((rc = dev_alloc(ctx, &ctx->d_E, (size_t)3 * m.nCorners))) return rc;
    CK(cudaMemset(ctx->d_E, 0, sizeof(double) * 3 * (size_t)m.nCorners));
  }
  ctx->fieldSolverReady = true;
  return AMPS_GPU_OK;
}

// Several ranks: who sends which node values to whom (the host derives the lists from the global node keys, amps_b200/mesh.py
// field_halo_lists; in AMPS: from the corner / centre nodes of DomainBoundaryLayerNodesList).  corner_send / corner_recv: local unique
// corner ids whose E-type values this rank sends to / receives from `peer` after every operator product (the sender is the corner's
// primary rank); center_send / center_recv: the same for B after UpdateB (the sender owns the cell).  Both sides list a pair's
// nodes in the same order.  primary[n_corners]: 1 where this rank counts the corner in the inner products.
int amps_gpu_field_halo_set(amps_gpu_ctx *ctx, int peer, const int32_t *corner_send, int64_t n_cs, const int32_t *corner_recv, int64_t n_cr,
                            const int32_t *center_send, int64_t n_zs, const int32_t *center_recv, int64_t n_zr) {
  if (!ctx || peer < 0 || peer >= ctx->nRanks || peer == ctx->rank) return AMPS_GPU_ERR_ARG;
  const int32_t *src[4] = {corner_send, corner_recv, center_send, center_recv};
  const int64_t cnt[4] = {n_cs, n_cr, n_zs, n_zr};
  for (int k = 0; k < 4; k++) {
    if (cnt[k] < 0 || (cnt[k] > 0 && !src[k])) return AMPS_GPU_ERR_ARG;
    ctx->fh.h[k].resize((size_t)ctx->nRanks);
    ctx->fh.h[k][peer].assign(src[k], src[k] + cnt[k]);
    const int lim = (k < 2) ? ctx->dm.nCorners : ctx->dm.nCenters;
    for (int v : ctx->fh.h[k][peer])
      if (v < 0 || v >= lim) FAIL(AMPS_GPU_ERR_ARG, "field halo node id out of range");
  }
  ctx->fh.dirty = true, ctx->fh.set = true;
  return AMPS_GPU_OK;
}
int amps_gpu_field_primary_set(amps_gpu_ctx *ctx, const uint8_t *primary) {
  if (!ctx || !primary) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "field_primary_set before mesh_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  int rc;
  if (!ctx->d_primary && (rc = dev_alloc(ctx, &ctx->d_primary, (size_t)ctx->dm.nCorners))) return rc;
  CK(cudaMemcpy(ctx->d_primary, primary, (size_t)ctx->dm.nCorners, cudaMemcpyHostToDevice));
  return AMPS_GPU_OK;
}
static int field_halo_build(amps_gpu_ctx *ctx) {
  auto &fh = ctx->fh;
  const int R = ctx->nRanks;
  int **dst[4] = {&fh.d_cSend, &fh.d_cRecv, &fh.d_zSend, &fh.d_zRecv};
  std::vector<long long> *off[4] = {&fh.cSendOff, &fh.cRecvOff, &fh.zSendOff, &fh.zRecvOff};
  long long maxTot = 0;
  for (int k = 0; k < 4; k++) {
    std::vector<int> all;
    off[k]->assign((size_t)R + 1, 0);
    fh.h[k].resize((size_t)R);
    for (int r = 0; r < R; r++) {
      (*off[k])[r] = (long long)all.size();
      all.insert(all.end(), fh.h[k][r].begin(), fh.h[k][r].end());
    }
    (*off[k])[R] = (long long)all.size();
    cudaFree(*dst[k]);
    *dst[k] = nullptr;
    if (!all.empty()) {
      CK(cudaMalloc(dst[k], all.size() * sizeof(int)));
      CK(cudaMemcpy(*dst[k], all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    maxTot = std::max<long long>(maxTot, (long long)all.size());
  }
  if (3 * maxTot > fh.bufEntries) {
    cudaFree(fh.d_sendBuf), cudaFree(fh.d_recvBuf);
    fh.bufEntries = 3 * maxTot;
    CK(cudaMalloc(&fh.d_sendBuf, sizeof(double) * (size_t)fh.bufEntries));
    CK(cudaMalloc(&fh.d_recvBuf, sizeof(double) * (size_t)fh.bufEntries));
  }
  fh.dirty = false;
  return AMPS_GPU_OK;
}
// refresh the copies other ranks hold of this rank's nodes: vec[n][3] on the corners (centers = false) or the centres
static int field_halo_exchange(amps_gpu_ctx *ctx, double *vec, bool centers) {
  if (ctx->nRanks <= 1) return AMPS_GPU_OK;
  auto &fh = ctx->fh;
  NcclApi &a = nccl_api();
  cudaStream_t s = ctx->stream;
  const std::vector<long long> &so = centers ? fh.zSendOff : fh.cSendOff, &ro = centers ? fh.zRecvOff : fh.cRecvOff;
  const int *sIds = centers ? fh.d_zSend : fh.d_cSend, *rIds = centers ? fh.d_zRecv : fh.d_cRecv;
  const int R = ctx->nRanks;
  launch_halo_pack(sIds, (int)so[R], vec, fh.d_sendBuf, s);
  NCK(a.GroupStart());
  for (int r = 0; r < R; r++) {
    if (r == ctx->rank) continue;
    const long long ns = so[r + 1] - so[r], nr = ro[r + 1] - ro[r];
    if (ns > 0) NCK(a.Send(fh.d_sendBuf + 3 * so[r], (size_t)(3 * ns), ncclDouble, r, ctx->comm, s));
    if (nr > 0) NCK(a.Recv(fh.d_recvBuf + 3 * ro[r], (size_t)(3 * nr), ncclDouble, r, ctx->comm, s));
  }
  NCK(a.GroupEnd());
  launch_halo_unpack(rIds, (int)ro[R], fh.d_recvBuf, vec, s);
  ctx->launches += 2;
  return AMPS_GPU_OK;
}

// E^n on the unique corners (E at the half step goes through amps_gpu_fields_upload)
int amps_gpu_E_upload(amps_gpu_ctx *ctx, const double *E_cur) {
  if (!ctx || !E_cur) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "E_upload before amps_gpu_mesh_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  if (!ctx->d_E) {  // without the device field solve: the guiding-centre movers of cfg.gc_fields_ecsim read it
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->d_E, (size_t)3 * ctx->dm.nCorners))) return rc;
  }
  CK(cudaMemcpyAsync(ctx->d_E, E_cur, sizeof(double) * 3 * (size_t)ctx->dm.nCorners, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->eReady = true;
  return AMPS_GPU_OK;
}

int amps_gpu_fields_download(amps_gpu_ctx *ctx, double *E_cur, double *E_half, double *B_cur) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "fields_download before mesh_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  const DevMesh &m = ctx->dm;
  const size_t nB = (ctx->cfg.b_mode == AMPS_B_CORNER_BASED) ? m.nCorners : m.nCenters;
  if (E_cur) {
    if (!ctx->d_E) FAIL(AMPS_GPU_ERR_STATE, "no E^n on the device (amps_gpu_field_solver_init)");
    CK(cudaMemcpyAsync(E_cur, ctx->d_E, sizeof(double) * 3 * (size_t)m.nCorners, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (E_half) CK(cudaMemcpyAsync(E_half, ctx->d_Ehalf, sizeof(double) * 3 * (size_t)m.nCorners, cudaMemcpyDeviceToHost, ctx->stream));
  if (B_cur) CK(cudaMemcpyAsync(B_cur, ctx->d_Bcur, sizeof(double) * 3 * nB, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AMPS_GPU_OK;
}

// One field step from the state on the device: E^n (amps_gpu_E_upload or the previous step), B^n (= B_cur of fields_upload or
// the previous step), J and M of the last deposit.  UpdateRhs -> GMRES(restart) from x0 = 0 on the relative residual (what the
// reference asks linear_solver_wrapper for, LinearSystemCornerNode.h:3282) -> E^{n+theta} = E^n + x (ProcessFinalSolution) ->
// UpdateB (B_prev <- B^n, B_cur <- B^{n+1}) -> UpdateE (E^n <- E^{n+1}); the tiles of the movers and of the deposit are staged
// again, so the particle step that follows sees E^{n+theta}, B^n and B^{n+1} exactly like PIC::TimeStep orders them.
int amps_gpu_field_step(amps_gpu_ctx *ctx, double theta, double tol, int max_iter, int restart, int warm_start, int *iterations,
                        double *rel_residual) {
  if (!ctx || theta <= 0.0 || tol <= 0.0 || max_iter < 1) return AMPS_GPU_ERR_ARG;
  if (!ctx->fieldSolverReady) FAIL(AMPS_GPU_ERR_STATE, "field_step before amps_gpu_field_solver_init");
  if (!ctx->fieldsReady) FAIL(AMPS_GPU_ERR_STATE, "field_step before amps_gpu_fields_upload (B^n)");
  CK(cudaSetDevice(ctx->cfg.device));
  const DevMesh &m = ctx->dm;
  cudaStream_t s = ctx->stream;
  const int n = 3 * m.nCorners;
  const bool multi = ctx->nRanks > 1;
  if (multi) {
    if (!ctx->comm) FAIL(AMPS_GPU_ERR_STATE, "field_step on several ranks before amps_gpu_comm_init");
    if (!ctx->fh.set || !ctx->d_primary) FAIL(AMPS_GPU_ERR_STATE, "field_step on several ranks needs amps_gpu_field_halo_set and amps_gpu_field_primary_set");
    int rch;
    if (ctx->fh.dirty && (rch = field_halo_build(ctx))) return rch;
  }
  const unsigned char *mask = multi ? ctx->d_primary : nullptr;
  // sums over all ranks of k doubles at p (device), in place.  (NCCL is only touched on several ranks: loading libnccl.so.2 into a
  // one-rank process would shadow the copy a host such as torch brings along when it is imported later.)
  auto allsum = [&](double *p, int k) -> int {
    if (multi) NCK(nccl_api().AllReduce(p, p, (size_t)k, ncclDouble, ncclSum, ctx->comm, s));
    return AMPS_GPU_OK;
  };
  if (restart < 1) restart = 30;
  if (restart > max_iter) restart = max_iter;
  int rc;
  if (ctx->krylovVectors < restart + 4) {
    cudaFree(ctx->d_krylov);
    ctx->d_krylov = nullptr;
    if ((rc = dev_alloc(ctx, &ctx->d_krylov, (size_t)(restart + 4) * n))) return rc;
    ctx->krylovVectors = restart + 4;
    ctx->fieldWarmValid = false;
    cudaFree(ctx->d_hcol), cudaFree(ctx->d_ycoef);
    if (ctx->h_hcol) cudaFreeHost(ctx->h_hcol);
    if ((rc = dev_alloc(ctx, &ctx->d_hcol, (size_t)(restart + 1) * (restart + 4)))) return rc;  // one column per Arnoldi step
    if ((rc = dev_alloc(ctx, &ctx->d_ycoef, (size_t)restart + 4))) return rc;
    CK(cudaMallocHost(&ctx->h_hcol, sizeof(double) * (restart + 2) * (restart + 4)));  // + one column for the coefficients y
  }
  // metric and time factors of GetStencil: coeff = theta c dt / dx, the operator constants K[9 slot + 3 p + q]
  double dx[3], coeff[3], c4rhs[3], c4b[3];
  const double cdt = ctx->cfg.ecsim_light_speed * ctx->cfg.ecsim_dt_total;
  for (int d = 0; d < 3; d++) {
    dx[d] = ctx->dxc0[d] * ctx->cfg.ecsim_length_conv;
    coeff[d] = cdt / dx[d] * theta;
    c4rhs[d] = 0.25 * coeff[d];
    c4b[d] = 0.25 * (cdt / dx[d]);
  }
  if (!ctx->d_fieldK && (rc = dev_alloc(ctx, &ctx->d_fieldK, (size_t)243))) return rc;
  if (ctx->fieldKtheta != theta) {
    static const double D2[3] = {1.0, -2.0, 1.0}, AV[3] = {0.25, 0.5, 0.25}, D1[3] = {-0.5, 0.0, 0.5};
    auto gd = [&](int p, int q, const int o[3]) {  // GradDivStencil[p][q] (== LaplacianStencil[p] for p == q), InitDiscritizationStencil
      double v = 1.0;
      for (int d = 0; d < 3; d++) {
        const double *t = (p == q) ? (d == p ? D2 : AV) : ((d == p || d == q) ? D1 : AV);
        v *= t[o[d] + 1];
      }
      return v;
    };
    double K[243];
    for (int a = -1; a <= 1; a++)
      for (int b = -1; b <= 1; b++)
        for (int c = -1; c <= 1; c++) {
          const int o[3] = {a, b, c};
          auto code = [](int d) { return (3 * d * d + d) >> 1; };
          const int slot = code(a) + 3 * code(b) + 9 * code(c);
          for (int p = 0; p < 3; p++)
            for (int q = 0; q < 3; q++) {
              double k = coeff[p] * coeff[q] * gd(p, q, o);
              if (p == q) {
                for (int e = 0; e < 3; e++) k -= coeff[e] * coeff[e] * gd(e, e, o);
                if (slot == 0) k += 1.0;
              }
              K[9 * slot + 3 * p + q] = k;
            }
        }
    CK(cudaMemcpyAsync(ctx->d_fieldK, K, sizeof(K), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));  // K is a local
    ctx->fieldKtheta = theta;
  }
  const double f = 4.0 * 3.14159265358979323846 * ctx->cfg.ecsim_dt_total * theta;
  double *V = ctx->d_krylov, *w = V + (size_t)(restart + 1) * n, *x = w + n, *rhsKeep = x + n;
  const double *Kc = ctx->d_fieldK;
  const size_t ld = (size_t)n;
  // right-hand side into w, then V_0 = r0 / |r0| (x0 = 0 -> r0 = rhs)
  launch_ecsim_operator(true, m.nCorners, ctx->d_fNb, ctx->d_fCc, Kc, ctx->d_M, ctx->d_E, f, ctx->d_J, ctx->d_Bcur, c4rhs, w, s);
  // several ranks: every vector of the iteration is kept consistent on the copies other ranks hold of a corner (ghost layers, shared
  // corners): the primary rank's value goes out after every product; linear combinations with all-reduced coefficients keep it so
  if ((rc = field_halo_exchange(ctx, w, false))) return rc;
  // x0 = 0 like the reference's SetInitialGuess (:6566), or the increment of the previous step (warm_start): the fields change
  // little from step to step, so most of the residual is gone before the first iteration.  Either way the iteration stops on
  // |r| <= tol |rhs|, which for x0 = 0 is the reference's relative residual.
  const bool warm = warm_start && ctx->fieldWarmValid && ctx->fieldWarmN == n;
  if (!warm) CK(cudaMemsetAsync(x, 0, sizeof(double) * n, s));
  ctx->launches++;
  std::vector<double> H((size_t)(restart + 1) * restart), cs(restart), sn(restart), g(restart + 1), y(restart);
  double r0norm = -1.0, rel = 1.0;
  int iters = 0;
  bool first = !warm;
  if (warm || max_iter > restart) CK(cudaMemcpyAsync(rhsKeep, w, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
  if (warm) {  // |rhs| is the reference of the stopping test
    launch_multi_dot(V, ld, 0, w, n, ctx->d_hcol, mask, s);
    if ((rc = allsum(ctx->d_hcol, 1))) return rc;
    CK(cudaMemcpyAsync(ctx->h_hcol, ctx->d_hcol, sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    r0norm = sqrt(ctx->h_hcol[0]);
  }
  // (after a restart the residual needs the right-hand side again: kept in the last spare vector)
  while (iters < max_iter) {
    if (!first) {
      // r = rhs - A x
      launch_ecsim_operator(false, m.nCorners, ctx->d_fNb, ctx->d_fCc, Kc, ctx->d_M, x, f, nullptr, nullptr, c4rhs, w, s);
      if ((rc = field_halo_exchange(ctx, w, false))) return rc;
      launch_axpby(n, 1.0, rhsKeep, -1.0, w, nullptr, w, s);
      ctx->launches += 2;
    }
    first = false;
    launch_multi_dot(V, ld, 0, w, n, ctx->d_hcol, mask, s);  // |w|^2
    if ((rc = allsum(ctx->d_hcol, 1))) return rc;
    CK(cudaMemcpyAsync(ctx->h_hcol, ctx->d_hcol, sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const double beta = sqrt(ctx->h_hcol[0]);
    if (r0norm < 0.0) r0norm = beta;
    rel = r0norm > 0.0 ? beta / r0norm : 0.0;
    if (beta == 0.0 || rel <= tol) break;
    launch_axpby(n, 1.0 / beta, w, 0.0, nullptr, nullptr, V, s);
    ctx->launches += 2;
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    // The host only needs the Hessenberg columns to decide when to stop.  The previous field step took lastFieldIters iterations:
    // that many Arnoldi steps less one are queued without a host round trip (every step has its own column on the device), the
    // columns come back together and the Givens rotations catch up; from there on the host looks after every step.  A step that
    // converges earlier than its predecessor has run a few products in vain; the solution uses the columns up to convergence only.
    const int hs = restart + 4;  // stride of a column in d_hcol / h_hcol
    const int blind = ctx->lastFieldIters - 1 - iters;
    int j = 0, done = 0;  // done: columns the host has rotated
    bool converged = false;
    CK(cudaMemsetAsync(ctx->d_hcol, 0, sizeof(double) * (size_t)(restart + 1) * hs, s));  // the accumulators of every column at once
    for (; j < restart && iters < max_iter && !converged;) {
      double *vj = V + (size_t)j * ld, *vn = V + (size_t)(j + 1) * ld;
      double *hc = ctx->d_hcol + (size_t)j * hs;
      launch_ecsim_operator(false, m.nCorners, ctx->d_fNb, ctx->d_fCc, Kc, ctx->d_M, vj, f, nullptr, nullptr, c4rhs, vn, s);
      iters++;
      if ((rc = field_halo_exchange(ctx, vn, false))) return rc;
      launch_multi_dot(V, ld, j + 1, vn, n, hc, mask, s, false);                      // h_i = V_i . w
      if ((rc = allsum(hc, j + 1))) return rc;
      launch_orthogonalize(V, ld, j + 1, hc, vn, n, hc + j + 2, mask, s, false);  // w -= sum h_i V_i, |w|^2
      if ((rc = allsum(hc + j + 2, 1))) return rc;
      launch_axpby(n, 1.0, vn, 0.0, nullptr, hc + j + 2, vn, s);                // V_{j+1} = w / |w|
      ctx->launches += 4;
      j++;
      if (j < blind && j < restart && iters < max_iter) continue;
      CK(cudaMemcpyAsync(ctx->h_hcol + (size_t)done * hs, ctx->d_hcol + (size_t)done * hs, sizeof(double) * (size_t)(j - done) * hs, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      for (; done < j; done++) {
        const int c = done;
        const double *hh = ctx->h_hcol + (size_t)c * hs;
        for (int i = 0; i <= c; i++) H[(size_t)i * restart + c] = hh[i];
        H[(size_t)(c + 1) * restart + c] = sqrt(hh[c + 2]);
        for (int i = 0; i < c; i++) {
          const double t = cs[i] * H[(size_t)i * restart + c] + sn[i] * H[(size_t)(i + 1) * restart + c];
          H[(size_t)(i + 1) * restart + c] = -sn[i] * H[(size_t)i * restart + c] + cs[i] * H[(size_t)(i + 1) * restart + c];
          H[(size_t)i * restart + c] = t;
        }
        const double a = H[(size_t)c * restart + c], b = H[(size_t)(c + 1) * restart + c], d = sqrt(a * a + b * b);
        cs[c] = a / d, sn[c] = b / d;
        H[(size_t)c * restart + c] = d, H[(size_t)(c + 1) * restart + c] = 0.0;
        g[c + 1] = -sn[c] * g[c];
        g[c] = cs[c] * g[c];
        rel = fabs(g[c + 1]) / r0norm;
        if (rel <= tol) {
          iters -= j - (c + 1);  // the products past convergence are not iterations of the solve
          j = c + 1;
          converged = true;
          break;
        }
      }
    }
    for (int i = j - 1; i >= 0; i--) {
      double t = g[i];
      for (int q = i + 1; q < j; q++) t -= H[(size_t)i * restart + q] * y[q];
      y[i] = t / H[(size_t)i * restart + i];
    }
    double *yPinned = ctx->h_hcol + (size_t)(restart + 1) * (restart + 4);  // pinned and owned by the context: no wait for the copy
    for (int i = 0; i < j; i++) yPinned[i] = y[i];
    CK(cudaMemcpyAsync(ctx->d_ycoef, yPinned, sizeof(double) * j, cudaMemcpyHostToDevice, s));
    launch_combine(V, ld, j, ctx->d_ycoef, x, n, s);
    ctx->launches++;
    if (rel <= tol) break;
  }
  // ProcessFinalSolution: E^{n+theta} = E^n + x
  launch_axpby(n, 1.0, ctx->d_E, 1.0, x, nullptr, ctx->d_Ehalf, s);
  // UpdateB: B^{n+1} into the buffer that held B_prev, then the two swap roles (CurrentBOffset <-> PrevBOffset)
  launch_update_B(m.nCenters, ctx->d_fZc, ctx->d_Ehalf, ctx->d_Bcur, c4b, ctx->d_Bprev, s);
  std::swap(ctx->d_Bcur, ctx->d_Bprev);
  if ((rc = field_halo_exchange(ctx, ctx->d_Bcur, true))) return rc;  // the ghost cells of this rank are cells of its neighbours
  // UpdateE: E^{n+1} = (E^{n+theta} - (1 - theta) E^n) / theta
  launch_axpby(n, 1.0 / theta, ctx->d_Ehalf, -(1.0 - theta) / theta, ctx->d_E, nullptr, ctx->d_E, s);
  launch_stage_tiles(m, false, ctx->d_Ehalf, ctx->d_Bprev, ctx->d_Bcur, ctx->d_eTile, ctx->d_bPrevTile, ctx->d_bCurTile, s);
  ctx->launches += 4;
  CK(cudaGetLastError());
  ctx->eReady = true;
  ctx->fieldWarmValid = true, ctx->fieldWarmN = n;  // x stays in the Krylov workspace for a warm start of the next step
  ctx->lastFieldIters = iters;
  if (iterations) *iterations = iters;
  if (rel_residual) *rel_residual = rel;
  return AMPS_GPU_OK;
}

int amps_gpu_background_upload(amps_gpu_ctx *ctx, const double *E_center, const double *B_center) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "background_upload before mesh_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  const DevMesh &m = ctx->dm;
  int rc;
  if (!ctx->d_bgTile) {
    if ((rc = dev_alloc(ctx, &ctx->d_bgE, (size_t)3 * m.nCenters))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_bgB, (size_t)3 * m.nCenters))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_bgTile, (size_t)m.nLeaves * m.nCenterLocal * 6))) return rc;
    CK(cudaMemsetAsync(ctx->d_bgTile, 0, sizeof(double) * (size_t)m.nLeaves * m.nCenterLocal * 6, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_bgE, 0, sizeof(double) * 3 * (size_t)m.nCenters, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_bgB, 0, sizeof(double) * 3 * (size_t)m.nCenters, ctx->stream));
  }
  if (E_center) CK(cudaMemcpyAsync(ctx->d_bgE, E_center, sizeof(double) * 3 * (size_t)m.nCenters, cudaMemcpyHostToDevice, ctx->stream));
  if (B_center) CK(cudaMemcpyAsync(ctx->d_bgB, B_center, sizeof(double) * 3 * (size_t)m.nCenters, cudaMemcpyHostToDevice, ctx->stream));
  launch_stage_background(m, E_center ? ctx->d_bgE : nullptr, B_center ? ctx->d_bgB : nullptr, ctx->d_bgTile, ctx->stream);
  ctx->launches++;
  CK(cudaGetLastError());
  ctx->backgroundReady = true;
  return AMPS_GPU_OK;
}

int amps_gpu_background_upload_gca(amps_gpu_ctx *ctx, const double *var15_center) {
  if (!ctx || !var15_center) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "background_upload_gca before mesh_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  const DevMesh &m = ctx->dm;
  int rc;
  if (!ctx->d_gcaTile) {
    if ((rc = dev_alloc(ctx, &ctx->d_gcaVar, (size_t)15 * m.nCenters))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_gcaTile, (size_t)m.nLeaves * m.nCenterLocal * 15))) return rc;
  }
  CK(cudaMemcpyAsync(ctx->d_gcaVar, var15_center, sizeof(double) * 15 * (size_t)m.nCenters, cudaMemcpyHostToDevice, ctx->stream));
  launch_stage_center_table(m, 15, ctx->d_gcaVar, ctx->d_gcaTile, ctx->stream);
  ctx->launches++;
  CK(cudaGetLastError());
  ctx->gcaReady = true;
  return AMPS_GPU_OK;
}

int amps_gpu_background_upload_gradB(amps_gpu_ctx *ctx, const double *gradB_center) {
  if (!ctx || !gradB_center) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "background_upload_gradB before mesh_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  const DevMesh &m = ctx->dm;
  int rc;
  if (!ctx->d_gradBTile) {
    if ((rc = dev_alloc(ctx, &ctx->d_gradBVar, (size_t)9 * m.nCenters))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_gradBTile, (size_t)m.nLeaves * m.nCenterLocal * 9))) return rc;
  }
  CK(cudaMemcpyAsync(ctx->d_gradBVar, gradB_center, sizeof(double) * 9 * (size_t)m.nCenters, cudaMemcpyHostToDevice, ctx->stream));
  launch_stage_center_table(m, 9, ctx->d_gradBVar, ctx->d_gradBTile, ctx->stream);
  ctx->launches++;
  CK(cudaGetLastError());
  ctx->gradBReady = true;
  return AMPS_GPU_OK;
}

int amps_gpu_magnetic_moment_init(amps_gpu_ctx *ctx, int mover_id) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  if (mover_id != AMPS_MOVER_RELATIVISTIC_GCA && mover_id != AMPS_MOVER_GC_FIRST_ORDER && mover_id != AMPS_MOVER_GC_SECOND_ORDER)
    FAIL(AMPS_GPU_ERR_ARG, "magnetic_moment_init: not a guiding-centre mover");
  if (!ctx->cfg.carry_magnetic_moment) FAIL(AMPS_GPU_ERR_STATE, "magnetic_moment_init needs cfg.carry_magnetic_moment");
  const bool gcEcsim = ctx->cfg.gc_fields_ecsim && mover_id != AMPS_MOVER_RELATIVISTIC_GCA;
  if (gcEcsim && (!ctx->meshReady || !ctx->fieldsReady || ctx->meshRefined || ctx->cfg.b_mode != AMPS_B_CENTER_BASED))
    FAIL(AMPS_GPU_ERR_STATE, "magnetic_moment_init with cfg.gc_fields_ecsim needs amps_gpu_fields_upload on a single-level mesh with centre-based B");
  if (!gcEcsim && (!ctx->meshReady || !ctx->backgroundReady)) FAIL(AMPS_GPU_ERR_STATE, "magnetic_moment_init needs the mesh and amps_gpu_background_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemsetAsync(ctx->d_stats, 0, sizeof(DevMoveStats), ctx->stream));
  if (mover_id == AMPS_MOVER_RELATIVISTIC_GCA)
    launch_magnetic_moment_init(ctx->dm, ctx->sp, ctx->cfg.coupler_interpolation, ctx->cfg.speed_of_light, ctx->buf[ctx->cur], ctx->d_n + ctx->cur,
                                ctx->nUpper, ctx->d_bgTile, ctx->d_bgE, ctx->d_bgB, ctx->d_stats, ctx->stream);
  else
    launch_gc_magnetic_moment_init(ctx->dm, ctx->sp, ctx->cfg.coupler_interpolation, ctx->buf[ctx->cur], ctx->d_n + ctx->cur, ctx->nUpper,
                                   ctx->d_bgTile, ctx->d_bgE, ctx->d_bgB, ctx->d_stats, ctx->stream, gcEcsim ? ctx->d_Bcur : nullptr,
                                   gcEcsim ? ctx->d_Bcur : nullptr);
  ctx->launches++;
  CK(cudaGetLastError());
  DevMoveStats h;
  CK(cudaMemcpyAsync(&h, ctx->d_stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (h.n_error) FAIL(AMPS_GPU_ERR_PARTICLE, "magnetic_moment_init: a particle lies outside the cell table of its block");
  return AMPS_GPU_OK;
}

static int upload_by_ptr(amps_gpu_ctx *ctx, bool vpar, const double *mu_by_ptr, int64_t n) {
  CK(cudaSetDevice(ctx->cfg.device));
  double *d = nullptr;
  CK(cudaMallocAsync((void **)&d, sizeof(double) * (size_t)(n ? n : 1), ctx->stream));
  CK(cudaMemcpyAsync(d, mu_by_ptr, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  launch_magnetic_moment_set(ctx->buf[ctx->cur], vpar ? ctx->buf[ctx->cur].vpar : ctx->buf[ctx->cur].mu, ctx->d_n + ctx->cur, ctx->nUpper, d, n,
                             ctx->stream);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaFreeAsync(d, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AMPS_GPU_OK;
}

int amps_gpu_magnetic_moment_upload(amps_gpu_ctx *ctx, const double *mu_by_ptr, int64_t n) {
  if (!ctx || !mu_by_ptr || n < 0) return AMPS_GPU_ERR_ARG;
  if (!ctx->cfg.carry_magnetic_moment) FAIL(AMPS_GPU_ERR_STATE, "magnetic_moment_upload needs cfg.carry_magnetic_moment");
  return upload_by_ptr(ctx, false, mu_by_ptr, n);
}
int amps_gpu_global_stencil_set(amps_gpu_ctx *ctx, int32_t full) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  ctx->sp.globalStencilFull = full ? 1 : 0;
  return AMPS_GPU_OK;
}
int amps_gpu_v_normal_upload(amps_gpu_ctx *ctx, const double *vnormal_by_ptr, int64_t n) {
  if (!ctx || !vnormal_by_ptr || n < 0) return AMPS_GPU_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->d_vnByPtr) cudaFree(ctx->d_vnByPtr);
  ctx->d_vnByPtr = nullptr, ctx->nVn = 0;
  int rc;
  if ((rc = dev_alloc(ctx, &ctx->d_vnByPtr, (size_t)n))) return rc;
  CK(cudaMemcpyAsync(ctx->d_vnByPtr, vnormal_by_ptr, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->nVn = n;
  return AMPS_GPU_OK;
}
int amps_gpu_v_parallel_upload(amps_gpu_ctx *ctx, const double *vpar_by_ptr, int64_t n) {
  if (!ctx || !vpar_by_ptr || n < 0) return AMPS_GPU_ERR_ARG;
  if (!ctx->cfg.carry_v_parallel) FAIL(AMPS_GPU_ERR_STATE, "v_parallel_upload needs cfg.carry_v_parallel");
  return upload_by_ptr(ctx, true, vpar_by_ptr, n);
}

static int download_current_order(amps_gpu_ctx *ctx, bool vpar, double *mu, int64_t n_max, int64_t *n) {
  CK(cudaSetDevice(ctx->cfg.device));
  int cnt = 0;
  CK(cudaMemcpyAsync(&cnt, ctx->d_n + ctx->cur, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *n = cnt;
  if (cnt > n_max) FAIL(AMPS_GPU_ERR_CAPACITY, "magnetic_moment_download: n_max too small");
  if (mu && cnt > 0)
    CK(cudaMemcpyAsync(mu, vpar ? ctx->buf[ctx->cur].vpar : ctx->buf[ctx->cur].mu, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AMPS_GPU_OK;
}
int amps_gpu_magnetic_moment_download(amps_gpu_ctx *ctx, double *mu, int64_t n_max, int64_t *n) {
  if (!ctx || !n) return AMPS_GPU_ERR_ARG;
  if (!ctx->cfg.carry_magnetic_moment) FAIL(AMPS_GPU_ERR_STATE, "magnetic_moment_download needs cfg.carry_magnetic_moment");
  return download_current_order(ctx, false, mu, n_max, n);
}
int amps_gpu_v_parallel_download(amps_gpu_ctx *ctx, double *vpar, int64_t n_max, int64_t *n) {
  if (!ctx || !n) return AMPS_GPU_ERR_ARG;
  if (!ctx->cfg.carry_v_parallel) FAIL(AMPS_GPU_ERR_STATE, "v_parallel_download needs cfg.carry_v_parallel");
  return download_current_order(ctx, true, vpar, n_max, n);
}

int amps_gpu_exit_records(amps_gpu_ctx *ctx, amps_gpu_exit_record *buf, int64_t max_records, int64_t *n) {
  if (!ctx || !n) return AMPS_GPU_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  unsigned long long cnt = 0;
  CK(cudaMemcpyAsync(&cnt, ctx->d_exitCount, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *n = (int64_t)cnt;
  long long have = (long long)cnt < ctx->cfg.exit_record_capacity ? (long long)cnt : ctx->cfg.exit_record_capacity;
  if (have > max_records) have = max_records;
  if (buf && have > 0) CK(cudaMemcpyAsync(buf, ctx->d_exitBuf, sizeof(amps_gpu_exit_record) * have, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_exitCount, 0, sizeof(unsigned long long), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AMPS_GPU_OK;
}

// after a sort the residents are exactly slots [0, count): the count travels to pinned host memory behind the sort and
// tighten_upper() adopts it once it has arrived, so the slot bound follows the resident population instead of growing with
// every arrival (false ERR_CAPACITY after enough steps on several ranks)
static int request_sorted_count(amps_gpu_ctx *ctx) {
  if (!ctx->h_nSorted) {
    CK(cudaMallocHost(&ctx->h_nSorted, sizeof(int)));
    CK(cudaEventCreateWithFlags(&ctx->evSorted, cudaEventDisableTiming));
  }
  CK(cudaMemcpyAsync(ctx->h_nSorted, ctx->d_n + ctx->cur, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (ctx->peerMigrate) {  // the error flag of the exchange travels with it (nobody waited for it during the step)
    CK(cudaMemcpyAsync(ctx->h_errLazy, ctx->d_errFlag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_errFlag, 0, sizeof(int), ctx->stream));
  }
  CK(cudaEventRecord(ctx->evSorted, ctx->stream));
  ctx->nSortedPending = true;
  return AMPS_GPU_OK;
}
static void tighten_upper(amps_gpu_ctx *ctx, bool wait) {
  if (!ctx->nSortedPending) return;
  if (wait) cudaEventSynchronize(ctx->evSorted);
  else if (cudaEventQuery(ctx->evSorted) != cudaSuccess) return;
  ctx->nUpper = *ctx->h_nSorted;
  ctx->nSortedPending = false;
  if (ctx->peerMigrate && *ctx->h_errLazy) ctx->pendingErr |= *ctx->h_errLazy, *ctx->h_errLazy = 0;
}

static int do_sort(amps_gpu_ctx *ctx) {
  ProfScope prof(ctx, AMPS_GPU_PHASE_SORT);
  ParticleSoA &src = ctx->buf[ctx->cur], &dst = ctx->buf[1 - ctx->cur];
  launch_sort(ctx->dm, src, dst, ctx->d_n + ctx->cur, ctx->d_cellCount, ctx->d_cellStart, ctx->d_cellFill, ctx->d_n + (1 - ctx->cur), ctx->nUpper,
              ctx->countValid, ctx->d_scanTmp, nullptr, ctx->stream, &ctx->launches);
  CK(cudaGetLastError());
  ctx->cur = 1 - ctx->cur;
  ctx->sorted = true;
  ctx->countValid = false;
  return request_sorted_count(ctx);
}

// amps_gpu_step: the counting sort only builds the permutation (8 B per particle); the deposit gathers through it and
// writes the sorted copy while it has the particle in registers.  Same end state as do_sort + do_deposit.
// several ranks: the leaves that touch a corner shared with another rank are deposited first, so that the exchange of those
// corners (pack, send/recv) runs next to the deposit of the interior leaves
static int rebuild_deposit_order(amps_gpu_ctx *ctx) {
  const DevMesh &m = ctx->dm;
  std::vector<char> shared((size_t)m.nCorners, 0);
  for (auto &v : ctx->h_sharedUid)
    for (int u : v) shared[u] = 1;
  std::vector<int> bnd, inner, ghost;
  for (int l = 0; l < m.nLeaves; l++) {
    if (ctx->h_leafGhost[l]) {
      ghost.push_back(l);
      continue;
    }
    const int *cu = ctx->h_leafCornerUid.data() + (size_t)l * m.nCornerLocal;
    bool touches = false;
    for (int kk = 0; kk <= m.N[2] && !touches; kk++)
      for (int jj = 0; jj <= m.N[1] && !touches; jj++)
        for (int ii = 0; ii <= m.N[0]; ii++) {
          const int u = cu[ii + m.g[0] + (m.TN[0] + 1) * (jj + m.g[1] + (kk + m.g[2]) * (m.TN[1] + 1))];
          if (u >= 0 && shared[u]) {
            touches = true;
            break;
          }
        }
    (touches ? bnd : inner).push_back(l);
  }
  ctx->nDepBoundary = (int)bnd.size();
  bnd.insert(bnd.end(), inner.begin(), inner.end());
  bnd.insert(bnd.end(), ghost.begin(), ghost.end());
  CK(cudaMemcpy(ctx->d_depLeaf, bnd.data(), sizeof(int) * bnd.size(), cudaMemcpyHostToDevice));
  ctx->depDirty = false;
  return AMPS_GPU_OK;
}

// Guiding-centre species of the deposit (cfg.gc_species_mask): deposit_kernel sees them with zero charge (nothing to J, M) and runs
// without its diagnostics; launch_gc_deposit then adds their explicit current, the magnetisation closure and the diagnostics of all
// species from the sorted store.
static int gc_check(amps_gpu_ctx *ctx) {
  if (!ctx->cfg.gc_species_mask) return AMPS_GPU_OK;
  if (!ctx->cfg.carry_magnetic_moment) FAIL(AMPS_GPU_ERR_STATE, "cfg.gc_species_mask needs cfg.carry_magnetic_moment (the closure deposits mu b)");
  if (ctx->nRanks > 1) FAIL(AMPS_GPU_ERR_STATE, "the guiding-centre species of the ECSIM deposit are built for one rank");
  return AMPS_GPU_OK;
}
static DevSpecies deposit_species(const amps_gpu_ctx *ctx) {
  DevSpecies sp = ctx->sp;
  for (int s = 0; s < sp.n; s++)
    if ((ctx->cfg.gc_species_mask >> s) & 1) sp.charge[s] = 0.0;
  return sp;
}
static unsigned gc_dep_flags(const amps_gpu_ctx *ctx) { return ctx->cfg.gc_species_mask ? (unsigned)DEP_NO_DIAG : 0u; }
static int gc_pass(amps_gpu_ctx *ctx, const ParticleSoA &sorted) {
  if (!ctx->cfg.gc_species_mask) return AMPS_GPU_OK;
  launch_gc_deposit(ctx->dm, ctx->sp, (unsigned)ctx->cfg.gc_species_mask, sorted, ctx->d_cellStart, ctx->d_bCurTile, ctx->d_vnByPtr, ctx->nVn, ctx->d_J,
                    ctx->d_energy, ctx->d_cfl, ctx->nSM, ctx->stream);
  ctx->launches++;
  CK(cudaGetLastError());
  return AMPS_GPU_OK;
}

static int do_sort_deposit_fused(amps_gpu_ctx *ctx) {
  if (!ctx->meshReady || !ctx->fieldsReady) FAIL(AMPS_GPU_ERR_STATE, "deposit before mesh/fields upload");
  if (ctx->meshRefined && ctx->cfg.b_mode == AMPS_B_CENTER_BASED)
    FAIL(AMPS_GPU_ERR_STATE, "ECSIM on a refined mesh needs _PIC_FIELD_SOLVER_B_CORNER_BASED_ (see amps_gpu_move)");
  int rc;
  if (!ctx->d_perm && (rc = dev_alloc(ctx, &ctx->d_perm, (size_t)ctx->cfg.capacity))) return rc;
  if ((rc = gc_check(ctx))) return rc;
  if (ctx->depDirty && ctx->nRanks > 1) {
    CK(cudaStreamSynchronize(ctx->stream));
    if ((rc = rebuild_deposit_order(ctx))) return rc;
  }
  ParticleSoA &src = ctx->buf[ctx->cur], &dst = ctx->buf[1 - ctx->cur];
  {
    ProfScope prof(ctx, AMPS_GPU_PHASE_SORT);
    launch_sort(ctx->dm, src, dst, ctx->d_n + ctx->cur, ctx->d_cellCount, ctx->d_cellStart, ctx->d_cellFill, ctx->d_n + (1 - ctx->cur), ctx->nUpper,
                ctx->countValid, ctx->d_scanTmp, ctx->d_perm, ctx->stream, &ctx->launches);
    CK(cudaGetLastError());
  }
  {
    ProfScope prof(ctx, AMPS_GPU_PHASE_DEPOSIT);
    const unsigned zero = ctx->jmZeroed ? 0u : (unsigned)DEP_ZERO_JM;
    const int nB = ctx->nDepBoundary;
    if (ctx->nRanks > 1 && ctx->comm && nB > 0 && nB < ctx->dm.nDepReal) {
      launch_deposit(ctx->dm, ctx->sp, src, ctx->d_cellStart, ctx->d_bCurTile, ctx->d_J, ctx->d_M, ctx->d_energy, ctx->d_cfl, ctx->nSM, ctx->d_perm,
                     dst, 0, nB, zero | DEP_ZERO_DIAG | DEP_GHOST_PASS, ctx->stream, &ctx->launches);
      if (!ctx->evBoundary) {
        CK(cudaEventCreateWithFlags(&ctx->evBoundary, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->evRecv, cudaEventDisableTiming));
      }
      CK(cudaEventRecord(ctx->evBoundary, ctx->stream));  // every shared corner has its local sum
      ctx->overlapJM = true;
      launch_deposit(ctx->dm, ctx->sp, src, ctx->d_cellStart, ctx->d_bCurTile, ctx->d_J, ctx->d_M, ctx->d_energy, ctx->d_cfl, ctx->nSM, ctx->d_perm,
                     dst, nB, -1, DEP_FINAL | DEP_SPARE_SMS, ctx->stream, &ctx->launches);
    } else {
      launch_deposit(ctx->dm, deposit_species(ctx), src, ctx->d_cellStart, ctx->d_bCurTile, ctx->d_J, ctx->d_M, ctx->d_energy, ctx->d_cfl, ctx->nSM,
                     ctx->d_perm, dst, 0, -1, zero | DEP_ZERO_DIAG | DEP_GHOST_PASS | DEP_FINAL | gc_dep_flags(ctx), ctx->stream, &ctx->launches);
    }
    CK(cudaGetLastError());
    if ((rc = gc_pass(ctx, dst))) return rc;
  }
  ctx->cur = 1 - ctx->cur;
  ctx->sorted = true;
  ctx->countValid = false;
  return request_sorted_count(ctx);
}

int amps_gpu_sort(amps_gpu_ctx *ctx) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "sort before mesh_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  return do_sort(ctx);
}

static int upload_soa(amps_gpu_ctx *ctx, const double *x, const double *v, const double *w, const uint8_t *species, const int32_t *cells,
                      const int32_t *ptrs, int64_t n, bool append);
int amps_gpu_particles_upload_soa(amps_gpu_ctx *ctx, const double *x, const double *v, const double *w, const uint8_t *species,
                                  const int32_t *cells, const int32_t *ptrs, int64_t n) {
  return upload_soa(ctx, x, v, w, species, cells, ptrs, n, false);
}
// the same behind the resident particles (injection between epochs; large populations handed over in pieces)
int amps_gpu_particles_append_soa(amps_gpu_ctx *ctx, const double *x, const double *v, const double *w, const uint8_t *species,
                                  const int32_t *cells, const int32_t *ptrs, int64_t n) {
  return upload_soa(ctx, x, v, w, species, cells, ptrs, n, true);
}
static int upload_soa(amps_gpu_ctx *ctx, const double *x, const double *v, const double *w, const uint8_t *species, const int32_t *cells,
                      const int32_t *ptrs, int64_t n, bool append) {
  if (!ctx || !x || !v || !species || !cells || n < 0) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "particles_upload before mesh_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  int64_t base = 0;
  if (append) {
    if (!ctx->sorted) FAIL(AMPS_GPU_ERR_STATE, "particles_append needs the sorted layout (the residents are slots [0, count))");
    int rc0 = amps_gpu_particle_count(ctx, &base);
    if (rc0) return rc0;
  }
  if (base + n > ctx->cfg.capacity) FAIL(AMPS_GPU_ERR_CAPACITY, "particle capacity exceeded");
  for (int64_t i = 0; i < n; i++) {
    if (cells[i] < 0 || cells[i] >= ctx->nCells) FAIL(AMPS_GPU_ERR_ARG, "particle cell out of range");
    // the kernels index the species tables with bits 0-5 (SetI exits on an out-of-range species, pic.h ParticleBuffer::SetI)
    if ((species[i] & 0x3f) >= ctx->sp.n) FAIL(AMPS_GPU_ERR_ARG, "particle species out of range");
  }
  ParticleSoA &b = ctx->buf[ctx->cur];
  cudaStream_t s = ctx->stream;
  for (int d = 0; d < 3; d++) {
    CK(cudaMemcpyAsync(b.x[d] + base, x + (size_t)d * n, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.v[d] + base, v + (size_t)d * n, sizeof(double) * n, cudaMemcpyHostToDevice, s));
  }
  std::vector<double> ones;
  std::vector<int32_t> iota;
  if (!w) {
    ones.assign((size_t)n, 1.0);
    w = ones.data();
  }
  if (!ptrs) {
    iota.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) iota[i] = (int32_t)(base + i);
    ptrs = iota.data();
  }
  CK(cudaMemcpyAsync(b.w + base, w, sizeof(double) * n, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(b.spec + base, species, n, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(b.key + base, cells, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(b.ptr + base, ptrs, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
  if (append) {
    if (b.mu) CK(cudaMemsetAsync(b.mu + base, 0, sizeof(double) * n, s));
    if (b.vpar) CK(cudaMemsetAsync(b.vpar + base, 0, sizeof(double) * n, s));
  }
  const int n32 = (int)(base + n);
  CK(cudaMemcpyAsync(ctx->d_n + ctx->cur, &n32, sizeof(int), cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));  // host temporaries
  ctx->nUpper = base + n;
  ctx->nSortedPending = false;
  ctx->countValid = false;
  if (append) ctx->h_uploadedSlots.insert(ctx->h_uploadedSlots.end(), ptrs, ptrs + n);
  else ctx->h_uploadedSlots.assign(ptrs, ptrs + n);
  return do_sort(ctx);
}

int amps_gpu_particles_upload_aos(amps_gpu_ctx *ctx, const void *records, const int64_t *ptrs, const int32_t *cells, int64_t n,
                                  const amps_gpu_aos_layout *lay) {
  if (!ctx || !records || !cells || !lay || n < 0) return AMPS_GPU_ERR_ARG;
  // AoS -> SoA on the host (one pass over the records), then the SoA path
  std::vector<double> x((size_t)3 * n), v((size_t)3 * n), w((size_t)n);
  std::vector<uint8_t> sp((size_t)n);
  std::vector<int32_t> pt((size_t)n);
  const unsigned char *base = (const unsigned char *)records;
  for (int64_t i = 0; i < n; i++) {
    const int64_t slot = ptrs ? ptrs[i] : i;
    const unsigned char *r = base + slot * lay->stride;
    double t[3];
    memcpy(t, r + lay->off_x, 24);
    x[i] = t[0], x[n + i] = t[1], x[2 * n + i] = t[2];
    memcpy(t, r + lay->off_v, 24);
    v[i] = t[0], v[n + i] = t[1], v[2 * n + i] = t[2];
    if (lay->off_w >= 0) memcpy(&w[i], r + lay->off_w, 8);
    else w[i] = 1.0;
    sp[i] = r[lay->off_species] & 0x7f;  // species + InitFlag (bit 6)
    pt[i] = (int32_t)slot;
  }
  int rc = amps_gpu_particles_upload_soa(ctx, x.data(), v.data(), w.data(), sp.data(), cells, pt.data(), n);
  if (rc) return rc;
  if (ctx->cfg.carry_magnetic_moment && lay->off_mu >= 0 && n > 0) {  // _PIC_PARTICLE_DATA__MAGNETIC_MOMENT_OFFSET_
    int64_t maxSlot = 0;
    for (int64_t i = 0; i < n; i++) maxSlot = pt[i] > maxSlot ? pt[i] : maxSlot;
    std::vector<double> mu((size_t)maxSlot + 1, 0.0);
    for (int64_t i = 0; i < n; i++) memcpy(&mu[pt[i]], base + (int64_t)pt[i] * lay->stride + lay->off_mu, 8);
    rc = amps_gpu_magnetic_moment_upload(ctx, mu.data(), maxSlot + 1);
  }
  if (!rc && ctx->cfg.carry_v_parallel && lay->off_vpar >= 0 && n > 0) {  // _PIC_PARTICLE_DATA__V_PARALLEL_OFFSET_
    int64_t maxSlot = 0;
    for (int64_t i = 0; i < n; i++) maxSlot = pt[i] > maxSlot ? pt[i] : maxSlot;
    std::vector<double> vp((size_t)maxSlot + 1, 0.0);
    for (int64_t i = 0; i < n; i++) memcpy(&vp[pt[i]], base + (int64_t)pt[i] * lay->stride + lay->off_vpar, 8);
    rc = amps_gpu_v_parallel_upload(ctx, vp.data(), maxSlot + 1);
  }
  return rc;
}

int amps_gpu_particle_count(amps_gpu_ctx *ctx, int64_t *n) {
  if (!ctx || !n) return AMPS_GPU_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  int n32 = 0;
  CK(cudaMemcpyAsync(&n32, ctx->d_n + ctx->cur, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *n = n32;
  return AMPS_GPU_OK;
}

int amps_gpu_particles_download_soa(amps_gpu_ctx *ctx, double *x, double *v, double *w, uint8_t *species, int32_t *cells, int32_t *ptrs,
                                    int64_t n_max, int64_t *n_out) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  int64_t n = 0;
  int rc = amps_gpu_particle_count(ctx, &n);
  if (rc) return rc;
  if (n_out) *n_out = n;
  if (n > n_max) FAIL(AMPS_GPU_ERR_CAPACITY, "download buffer too small");
  const ParticleSoA &b = ctx->buf[ctx->cur];
  cudaStream_t s = ctx->stream;
  // component-major with stride n_max so that the caller can size buffers before knowing n
  for (int d = 0; d < 3; d++) {
    if (x) CK(cudaMemcpyAsync(x + (size_t)d * n_max, b.x[d], sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    if (v) CK(cudaMemcpyAsync(v + (size_t)d * n_max, b.v[d], sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  }
  if (w) CK(cudaMemcpyAsync(w, b.w, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  if (species) CK(cudaMemcpyAsync(species, b.spec, n, cudaMemcpyDeviceToHost, s));
  if (cells) CK(cudaMemcpyAsync(cells, b.key, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  if (ptrs) CK(cudaMemcpyAsync(ptrs, b.ptr, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return AMPS_GPU_OK;
}

int amps_gpu_particles_download_aos(amps_gpu_ctx *ctx, void *records, int64_t *first_cell_particle, int64_t n_max, const amps_gpu_aos_layout *lay,
                                    int64_t *n_out) {
  if (!ctx || !records || !lay) return AMPS_GPU_ERR_ARG;
  int64_t n = 0;
  int rc = amps_gpu_particle_count(ctx, &n);
  if (rc) return rc;
  if (n_out) *n_out = n;
  std::vector<double> x((size_t)3 * n + 1), v((size_t)3 * n + 1), w((size_t)n + 1);
  std::vector<uint8_t> sp((size_t)n + 1);
  std::vector<int32_t> key((size_t)n + 1), pt((size_t)n + 1);
  int64_t nn;
  rc = amps_gpu_particles_download_soa(ctx, x.data(), v.data(), w.data(), sp.data(), key.data(), pt.data(), n, &nn);
  if (rc) return rc;
  unsigned char *base = (unsigned char *)records;
  std::vector<double> mu;
  if (ctx->cfg.carry_magnetic_moment && lay->off_mu >= 0) {
    mu.resize((size_t)n + 1);
    int64_t nm;
    if ((rc = amps_gpu_magnetic_moment_download(ctx, mu.data(), n, &nm))) return rc;
  }
  std::vector<double> vpar;
  if (ctx->cfg.carry_v_parallel && lay->off_vpar >= 0) {
    vpar.resize((size_t)n + 1);
    int64_t nm;
    if ((rc = amps_gpu_v_parallel_download(ctx, vpar.data(), n, &nm))) return rc;
  }
  if (first_cell_particle)
    for (int64_t c = 0; c < ctx->nCells; c++) first_cell_particle[c] = -1;
  for (int64_t i = 0; i < n; i++) {
    const int64_t slot = pt[i];
    if (slot < 0) FAIL(AMPS_GPU_ERR_STATE, "a resident record has no ParticleBuffer slot (it arrived by migration): amps_gpu_particles_slot_delta + _assign_slots first");
    if (slot >= n_max) FAIL(AMPS_GPU_ERR_ARG, "particle slot outside the caller's buffer");
    unsigned char *r = base + slot * lay->stride;
    double t[3] = {x[i], x[n + i], x[2 * n + i]};
    memcpy(r + lay->off_x, t, 24);
    double u[3] = {v[i], v[n + i], v[2 * n + i]};
    memcpy(r + lay->off_v, u, 24);
    if (lay->off_w >= 0) memcpy(r + lay->off_w, &w[i], 8);
    if (!mu.empty()) memcpy(r + lay->off_mu, &mu[i], 8);
    if (!vpar.empty()) memcpy(r + lay->off_vpar, &vpar[i], 8);
    r[lay->off_species] = (unsigned char)((r[lay->off_species] & 0x80) | (sp[i] & 0x7f));
    if (first_cell_particle && key[i] >= 0) {
      // push on the cell list like the movers do (pic_mover_boris.cpp:1333-1343)
      int64_t *first = first_cell_particle + key[i];
      const int64_t old = *first, none = -1;
      memcpy(r + lay->off_next, &old, 8);
      memcpy(r + lay->off_prev, &none, 8);
      if (old != -1) memcpy(base + old * lay->stride + lay->off_prev, &slot, 8);
      *first = slot;
    }
  }
  return AMPS_GPU_OK;
}

// Slot bookkeeping of the caller's PIC::ParticleBuffer at an epoch end: arrivals of amps_gpu_migrate carry ptr = -1
// (no record on this rank yet: GetNewParticle, pic_pbuffer.cpp:371-437), and the records of particles that a mover / boundary
// deleted or that migrated away must go back to the free list (DeleteParticle, pic_pbuffer.cpp:594-666).
int amps_gpu_particles_slot_delta(amps_gpu_ctx *ctx, int64_t *n_new, int64_t *released, int64_t max_released, int64_t *n_released) {
  if (!ctx || !n_new || !n_released || max_released < 0 || (max_released > 0 && !released)) return AMPS_GPU_ERR_ARG;
  int64_t n = 0;
  int rc = amps_gpu_particle_count(ctx, &n);
  if (rc) return rc;
  std::vector<int32_t> pt((size_t)n + 1);
  CK(cudaMemcpy(pt.data(), ctx->buf[ctx->cur].ptr, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  int32_t maxSlot = -1;
  for (int32_t v : ctx->h_uploadedSlots) maxSlot = v > maxSlot ? v : maxSlot;
  std::vector<char> present((size_t)maxSlot + 1, 0);
  int64_t nNew = 0;
  for (int64_t i = 0; i < n; i++) {
    if (pt[i] < 0) nNew++;
    else if (pt[i] <= maxSlot) present[pt[i]] = 1;
  }
  int64_t nRel = 0;
  for (int32_t v : ctx->h_uploadedSlots)
    if (v >= 0 && !present[v]) {
      if (nRel < max_released) released[nRel] = v;
      nRel++;
    }
  *n_new = nNew, *n_released = nRel;
  if (nRel <= max_released) {  // delivered: the released slots leave the books
    std::vector<int32_t> keep;
    keep.reserve(ctx->h_uploadedSlots.size());
    for (int32_t v : ctx->h_uploadedSlots)
      if (v >= 0 && present[v]) keep.push_back(v);
    ctx->h_uploadedSlots.swap(keep);
  }
  return AMPS_GPU_OK;
}

int amps_gpu_particles_assign_slots(amps_gpu_ctx *ctx, const int64_t *slots, int64_t n_slots) {
  if (!ctx || n_slots < 0 || (n_slots > 0 && !slots)) return AMPS_GPU_ERR_ARG;
  int64_t n = 0;
  int rc = amps_gpu_particle_count(ctx, &n);
  if (rc) return rc;
  std::vector<int32_t> pt((size_t)n + 1);
  CK(cudaMemcpy(pt.data(), ctx->buf[ctx->cur].ptr, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  int64_t k = 0;
  for (int64_t i = 0; i < n; i++)
    if (pt[i] < 0) {
      if (k >= n_slots) FAIL(AMPS_GPU_ERR_ARG, "assign_slots: fewer slots than records without one (see amps_gpu_particles_slot_delta)");
      if (slots[k] < 0 || slots[k] > 0x7fffffffLL) FAIL(AMPS_GPU_ERR_ARG, "assign_slots: slot out of range");
      pt[i] = (int32_t)slots[k++];
      ctx->h_uploadedSlots.push_back(pt[i]);
    }
  if (k != n_slots) FAIL(AMPS_GPU_ERR_ARG, "assign_slots: more slots than records without one");
  CK(cudaMemcpy(ctx->buf[ctx->cur].ptr, pt.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice));
  return AMPS_GPU_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// f3 (restart half): PIC::Restart::SaveParticleData / ReadParticleData (pic_restart.cpp:248-420, :430-600) straight from / into the
// sorted device store, in the reference's own file format, so a restart file written here is read by AMPS and vice versa:
//   [header written by the caller: user data, the 43-byte end marker, long ParticleDataLength, double GlobalParticleWeight[nS]]
//   per block that holds particles:  cAMRnodeID (id_bytes) | int nTotal | int ParticleNumberTable[Nx][Ny][Nz] (k fastest) |
//                                    nTotal records of ParticleDataLength bytes, cells in the loop order i, j, k (k innermost)
// The caller supplies the node id of every leaf (node->AMRnodeID, opaque here).  Record fields the device does not hold stay zero.
// ---------------------------------------------------------------------------------------------------------------------------
int amps_gpu_restart_save(amps_gpu_ctx *ctx, const char *fname, const void *header, int64_t header_bytes, const void *leaf_node_ids,
                          int32_t id_bytes, const amps_gpu_aos_layout *lay, int64_t *n_saved) {
  if (!ctx || !fname || !leaf_node_ids || id_bytes < 1 || !lay || header_bytes < 0 || (header_bytes > 0 && !header)) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "restart_save before mesh_upload");
  if (!ctx->sorted) FAIL(AMPS_GPU_ERR_STATE, "restart_save needs the sorted layout: call amps_gpu_sort");
  int64_t n = 0;
  int rc = amps_gpu_particle_count(ctx, &n);
  if (rc) return rc;
  std::vector<double> x((size_t)3 * n + 1), v((size_t)3 * n + 1), w((size_t)n + 1), mu, vpar;
  std::vector<uint8_t> sp((size_t)n + 1);
  int64_t nn;
  if ((rc = amps_gpu_particles_download_soa(ctx, x.data(), v.data(), w.data(), sp.data(), nullptr, nullptr, n, &nn))) return rc;
  if (ctx->cfg.carry_magnetic_moment && lay->off_mu >= 0) {
    mu.resize((size_t)n + 1);
    if ((rc = amps_gpu_magnetic_moment_download(ctx, mu.data(), n, &nn))) return rc;
  }
  if (ctx->cfg.carry_v_parallel && lay->off_vpar >= 0) {
    vpar.resize((size_t)n + 1);
    if ((rc = amps_gpu_v_parallel_download(ctx, vpar.data(), n, &nn))) return rc;
  }
  std::vector<int> start((size_t)ctx->nCells + 1);
  CK(cudaMemcpy(start.data(), ctx->d_cellStart, sizeof(int) * ((size_t)ctx->nCells + 1), cudaMemcpyDeviceToHost));
  FILE *f = fopen(fname, "wb");
  if (!f) FAIL(AMPS_GPU_ERR_ARG, "restart_save: cannot open the file");
  if (header_bytes) fwrite(header, 1, (size_t)header_bytes, f);
  const DevMesh &m = ctx->dm;
  const int Nx = m.N[0], Ny = m.N[1], Nz = m.N[2], C = m.cellsPerBlock;
  std::vector<int> table((size_t)C);
  std::vector<unsigned char> rec((size_t)lay->stride);
  int64_t saved = 0;
  for (int l = 0; l < m.nLeaves; l++) {
    const int total = start[(size_t)(l + 1) * C] - start[(size_t)l * C];
    if (total == 0) continue;
    fwrite((const unsigned char *)leaf_node_ids + (size_t)l * id_bytes, 1, (size_t)id_bytes, f);
    fwrite(&total, sizeof(int), 1, f);
    for (int i = 0; i < Nx; i++)
      for (int j = 0; j < Ny; j++)
        for (int k = 0; k < Nz; k++) {
          const size_t c = (size_t)l * C + i + Nx * (j + Ny * k);
          table[(size_t)k + Nz * (j + Ny * i)] = start[c + 1] - start[c];
        }
    fwrite(table.data(), sizeof(int), (size_t)C, f);
    for (int i = 0; i < Nx; i++)
      for (int j = 0; j < Ny; j++)
        for (int k = 0; k < Nz; k++) {
          const size_t c = (size_t)l * C + i + Nx * (j + Ny * k);
          for (int q = start[c]; q < start[c + 1]; q++) {
            std::fill(rec.begin(), rec.end(), 0);
            const int64_t none = -1;
            memcpy(rec.data() + lay->off_next, &none, 8), memcpy(rec.data() + lay->off_prev, &none, 8);
            const double xx[3] = {x[q], x[n + q], x[2 * n + q]}, vv[3] = {v[q], v[n + q], v[2 * n + q]};
            memcpy(rec.data() + lay->off_x, xx, 24), memcpy(rec.data() + lay->off_v, vv, 24);
            if (lay->off_w >= 0) memcpy(rec.data() + lay->off_w, &w[q], 8);
            if (!mu.empty()) memcpy(rec.data() + lay->off_mu, &mu[q], 8);
            if (!vpar.empty()) memcpy(rec.data() + lay->off_vpar, &vpar[q], 8);
            rec[lay->off_species] = (unsigned char)(0x80 | (sp[q] & 0x7f));  // allocated flag (bit 7) + InitFlag + species id
            fwrite(rec.data(), 1, rec.size(), f);
            saved++;
          }
        }
  }
  fclose(f);
  if (n_saved) *n_saved = saved;
  return AMPS_GPU_OK;
}

int amps_gpu_restart_read(amps_gpu_ctx *ctx, const char *fname, int64_t header_bytes, const void *leaf_node_ids, int32_t id_bytes,
                          const amps_gpu_aos_layout *lay, int64_t *n_loaded) {
  if (!ctx || !fname || !leaf_node_ids || id_bytes < 1 || !lay || header_bytes < 0) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "restart_read before mesh_upload");
  FILE *f = fopen(fname, "rb");
  if (!f) FAIL(AMPS_GPU_ERR_ARG, "restart_read: cannot open the file");
  fseek(f, (long)header_bytes, SEEK_SET);
  const DevMesh &m = ctx->dm;
  const int Nx = m.N[0], Ny = m.N[1], Nz = m.N[2], C = m.cellsPerBlock;
  std::map<std::string, int> leafOf;
  for (int l = 0; l < m.nLeaves; l++) leafOf[std::string((const char *)leaf_node_ids + (size_t)l * id_bytes, (size_t)id_bytes)] = l;
  std::vector<double> x[3], v[3], w, mu, vpar;
  std::vector<uint8_t> sp;
  std::vector<int32_t> cells;
  std::vector<char> id((size_t)id_bytes);
  std::vector<int> table((size_t)C);
  std::vector<unsigned char> rec((size_t)lay->stride);
  const bool wantMu = ctx->cfg.carry_magnetic_moment && lay->off_mu >= 0, wantVp = ctx->cfg.carry_v_parallel && lay->off_vpar >= 0;
  while (fread(id.data(), 1, (size_t)id_bytes, f) == (size_t)id_bytes) {
    int total = 0;
    if (fread(&total, sizeof(int), 1, f) != 1) break;
    if (total == 0) continue;
    auto it = leafOf.find(std::string(id.data(), (size_t)id_bytes));
    const bool mine = it != leafOf.end() && (ctx->nRanks <= 1 || ctx->h_leafOwnerHost.empty() || ctx->h_leafOwnerHost[it->second] == ctx->rank);
    if (!mine) {  // a block of another rank (ReadParticleDataBlock skips it the same way, :457-465)
      fseek(f, (long)(sizeof(int) * (size_t)C + (size_t)total * (size_t)lay->stride), SEEK_CUR);
      continue;
    }
    if (fread(table.data(), sizeof(int), (size_t)C, f) != (size_t)C) {
      fclose(f);
      FAIL(AMPS_GPU_ERR_ARG, "restart_read: truncated file");
    }
    const int l = it->second;
    for (int i = 0; i < Nx; i++)
      for (int j = 0; j < Ny; j++)
        for (int k = 0; k < Nz; k++)
          for (int q = 0; q < table[(size_t)k + Nz * (j + Ny * i)]; q++) {
            if (fread(rec.data(), 1, rec.size(), f) != rec.size()) {
              fclose(f);
              FAIL(AMPS_GPU_ERR_ARG, "restart_read: truncated file");
            }
            double t[3];
            memcpy(t, rec.data() + lay->off_x, 24);
            for (int d = 0; d < 3; d++) x[d].push_back(t[d]);
            memcpy(t, rec.data() + lay->off_v, 24);
            for (int d = 0; d < 3; d++) v[d].push_back(t[d]);
            double ww = 1.0;
            if (lay->off_w >= 0) memcpy(&ww, rec.data() + lay->off_w, 8);
            w.push_back(ww);
            sp.push_back(rec[lay->off_species] & 0x7f);
            cells.push_back((int32_t)((size_t)l * C + i + Nx * (j + Ny * k)));
            if (wantMu) {
              double a;
              memcpy(&a, rec.data() + lay->off_mu, 8);
              mu.push_back(a);
            }
            if (wantVp) {
              double a;
              memcpy(&a, rec.data() + lay->off_vpar, 8);
              vpar.push_back(a);
            }
          }
  }
  fclose(f);
  const int64_t n = (int64_t)w.size();
  std::vector<double> xs((size_t)3 * n + 1), vs((size_t)3 * n + 1);
  for (int d = 0; d < 3; d++)
    for (int64_t i = 0; i < n; i++) xs[(size_t)d * n + i] = x[d][i], vs[(size_t)d * n + i] = v[d][i];
  int rc = amps_gpu_particles_upload_soa(ctx, xs.data(), vs.data(), w.data(), sp.data(), cells.data(), nullptr, n);
  if (rc) return rc;
  // upload_soa sorts; the optional state goes in by ParticleBuffer slot (= the order read)
  if (wantMu && n && (rc = amps_gpu_magnetic_moment_upload(ctx, mu.data(), n))) return rc;
  if (wantVp && n && (rc = amps_gpu_v_parallel_upload(ctx, vpar.data(), n))) return rc;
  if (n_loaded) *n_loaded = n;
  return AMPS_GPU_OK;
}

int amps_gpu_cell_table_download(amps_gpu_ctx *ctx, int64_t *cell_start, int64_t n_cells_plus_1) {
  if (!ctx || !cell_start) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady || n_cells_plus_1 != ctx->nCells + 1) FAIL(AMPS_GPU_ERR_ARG, "cell table size mismatch");
  if (!ctx->sorted) FAIL(AMPS_GPU_ERR_STATE, "cell table is stale: call amps_gpu_sort first");
  CK(cudaSetDevice(ctx->cfg.device));
  std::vector<int> tmp((size_t)n_cells_plus_1);
  CK(cudaMemcpyAsync(tmp.data(), ctx->d_cellStart, sizeof(int) * n_cells_plus_1, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int64_t i = 0; i < n_cells_plus_1; i++) cell_start[i] = tmp[i];
  return AMPS_GPU_OK;
}

// Structure cache of the coupler's AMR stencil (cplr_stencil.cuh): neighbour nodes of every block and, for every block that has a
// finer neighbour (the coarse block of a multi-block stencil), the cells behind the 8 logical centres of each dual cell.  Built at
// the first test-particle move on a refined mesh; AMPS_GPU_CPLR_CACHE=0 keeps the uncached path (the equivalence test).
static int build_cplr_cache(amps_gpu_ctx *ctx) {
  ctx->cplrCacheTried = true;
  const char *sw = getenv("AMPS_GPU_CPLR_CACHE");
  if (sw && atoi(sw) == 0) return AMPS_GPU_OK;
  DevMesh &m = ctx->dm;
  std::vector<int> slot((size_t)m.nLeaves, -1), tabLeaf;
  for (int l = 0; l < m.nLeaves; l++) {
    const int mx = (int)(short)((ctx->h_leafNeib[l] >> 16) & 0xffff);
    if (mx > ctx->h_leafLevel[l]) slot[l] = (int)tabLeaf.size(), tabLeaf.push_back(l);
  }
  const size_t bytes = cplr_cache_table_bytes(m) * tabLeaf.size();
  if (bytes > ((size_t)8 << 30)) return AMPS_GPU_OK;  // declined: the stencils are built per particle
  int rc;
  if ((rc = dev_alloc(ctx, &ctx->d_neib26, (size_t)m.nLeaves * 27))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_mbSlot, (size_t)m.nLeaves))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_mbLeaf, tabLeaf.size() + 1))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_mbTab, bytes + 8))) return rc;
  CK(cudaMemcpyAsync(ctx->d_mbSlot, slot.data(), sizeof(int) * slot.size(), cudaMemcpyHostToDevice, ctx->stream));
  if (!tabLeaf.empty()) CK(cudaMemcpyAsync(ctx->d_mbLeaf, tabLeaf.data(), sizeof(int) * tabLeaf.size(), cudaMemcpyHostToDevice, ctx->stream));
  launch_build_cplr_cache(m, ctx->d_neib26, ctx->d_mbLeaf, (int)tabLeaf.size(), ctx->d_mbTab, ctx->stream);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));  // slot / tabLeaf are locals
  m.neib26 = ctx->d_neib26, m.mbSlot = ctx->d_mbSlot, m.mbTab = ctx->d_mbTab;
  return AMPS_GPU_OK;
}

static int do_move(amps_gpu_ctx *ctx, int mover_id) {
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "move before mesh upload");
  tighten_upper(ctx, false);
  if (ctx->pendingErr) {
    const int e = ctx->pendingErr;
    ctx->pendingErr = 0;
    if (e & 1) FAIL(AMPS_GPU_ERR_CAPACITY, "the previous particle exchange overflowed a receive region (more than capacity/32 leavers from one rank to another)");
    FAIL(AMPS_GPU_ERR_STATE, "the previous particle exchange dropped arrivals (unknown leaf, foreign owner or no free slot)");
  }
  if (!ctx->sorted) FAIL(AMPS_GPU_ERR_STATE, "move needs the (block,cell)-sorted layout: call amps_gpu_sort");
  if (mover_id != AMPS_MOVER_LAPENTA2017 && mover_id != AMPS_MOVER_RELATIVISTIC_BORIS && mover_id != AMPS_MOVER_BORIS &&
      mover_id != AMPS_MOVER_RELATIVISTIC_GCA && mover_id != AMPS_MOVER_GC_FIRST_ORDER && mover_id != AMPS_MOVER_GC_SECOND_ORDER &&
      mover_id != AMPS_MOVER_MARKIDIS2010 && mover_id != AMPS_MOVER_GYROKINETIC_FIRST_ORDER && mover_id != AMPS_MOVER_GYROKINETIC_SECOND_ORDER)
    FAIL(AMPS_GPU_ERR_ARG, "unknown mover id");
  if (mover_id == AMPS_MOVER_GYROKINETIC_FIRST_ORDER || mover_id == AMPS_MOVER_GYROKINETIC_SECOND_ORDER) {
    if (!ctx->cfg.carry_magnetic_moment || !ctx->cfg.carry_v_parallel)
      FAIL(AMPS_GPU_ERR_STATE, "the gyrokinetic movers need cfg.carry_magnetic_moment and cfg.carry_v_parallel");
    if (!ctx->gradBReady) FAIL(AMPS_GPU_ERR_STATE, "PIC::GYROKINETIC needs amps_gpu_background_upload_gradB");
  }
  const bool gcEcsim = ctx->cfg.gc_fields_ecsim && (mover_id == AMPS_MOVER_GC_FIRST_ORDER || mover_id == AMPS_MOVER_GC_SECOND_ORDER);
  if (mover_id == AMPS_MOVER_GC_FIRST_ORDER || mover_id == AMPS_MOVER_GC_SECOND_ORDER) {
    if (!ctx->cfg.carry_magnetic_moment) FAIL(AMPS_GPU_ERR_STATE, "the guiding-centre movers need cfg.carry_magnetic_moment");
    if (gcEcsim) {
      if (!ctx->fieldsReady || !ctx->eReady) FAIL(AMPS_GPU_ERR_STATE, "cfg.gc_fields_ecsim needs amps_gpu_fields_upload (B_cur) and the current E (amps_gpu_E_upload / field_step)");
      if (ctx->meshRefined || ctx->cfg.b_mode != AMPS_B_CENTER_BASED)
        FAIL(AMPS_GPU_ERR_STATE, "cfg.gc_fields_ecsim: single-level meshes with centre-based B (ECSIM::GetMagneticField reads the centre nodes)");
    } else if (!ctx->gradBReady) FAIL(AMPS_GPU_ERR_STATE, "GuidingCenter needs amps_gpu_background_upload_gradB");
  }
  if (mover_id == AMPS_MOVER_RELATIVISTIC_GCA) {
    if (!ctx->cfg.carry_magnetic_moment) FAIL(AMPS_GPU_ERR_STATE, "the guiding-centre movers need cfg.carry_magnetic_moment");
    if (!ctx->gcaReady) FAIL(AMPS_GPU_ERR_STATE, "Relativistic::GuidingCenter needs amps_gpu_background_upload_gca");
    if (ctx->cfg.boundary_mode != AMPS_BOUNDARY_DELETE)
      FAIL(AMPS_GPU_ERR_STATE, "Relativistic::GuidingCenter implements the DELETE boundary only (reference :262-279)");
  }
  if (mover_id == AMPS_MOVER_LAPENTA2017 && !ctx->fieldsReady) FAIL(AMPS_GPU_ERR_STATE, "Lapenta2017 needs amps_gpu_fields_upload");
  if (mover_id == AMPS_MOVER_LAPENTA2017 && ctx->meshRefined && ctx->cfg.b_mode == AMPS_B_CENTER_BASED)
    FAIL(AMPS_GPU_ERR_STATE,
         "ECSIM on a refined mesh needs _PIC_FIELD_SOLVER_B_CORNER_BASED_: with centre-based B the reference indexes the start block's "
         "buffer with stencil ids of other blocks (pic_mover_boris.cpp:975-990)");
  if (mover_id != AMPS_MOVER_LAPENTA2017 && !gcEcsim && !ctx->backgroundReady) FAIL(AMPS_GPU_ERR_STATE, "the test-particle movers need amps_gpu_background_upload");
  if (mover_id != AMPS_MOVER_LAPENTA2017 && ctx->meshRefined && !ctx->cplrCacheTried &&
      ctx->cfg.coupler_interpolation == AMPS_CPLR_CELL_CENTERED_LINEAR) {
    int rc;
    if ((rc = build_cplr_cache(ctx))) return rc;
  }
  const DevMesh &m = ctx->dm;
  ProfScope prof(ctx, AMPS_GPU_PHASE_MOVE);
  CK(cudaMemsetAsync(ctx->d_cellCount, 0, sizeof(int) * (size_t)ctx->nCells, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_stats, 0, sizeof(DevMoveStats), ctx->stream));
  if (mover_id == AMPS_MOVER_BORIS || mover_id == AMPS_MOVER_MARKIDIS2010) {
    launch_move_boris(m, ctx->sp, mover_id == AMPS_MOVER_MARKIDIS2010, ctx->cfg.coupler_interpolation, ctx->cfg.backward_time_integration, ctx->cfg.speed_of_light,
                      ctx->cfg.internal_sphere_radius, ctx->cfg.exit_record_capacity, ctx->cfg.gravity_gm, ctx->buf[ctx->cur], ctx->d_n + ctx->cur,
                      ctx->nUpper, ctx->d_bgTile, ctx->d_bgE, ctx->d_bgB, ctx->d_cellCount, ctx->d_stats, ctx->d_exitBuf, ctx->d_exitCount,
                      ctx->stream);
    ctx->launches++;
    CK(cudaGetLastError());
    ctx->sorted = false;
    ctx->countValid = true;
    return AMPS_GPU_OK;
  }
  if (mover_id == AMPS_MOVER_GC_FIRST_ORDER || mover_id == AMPS_MOVER_GC_SECOND_ORDER) {
    launch_move_guiding_center(m, ctx->sp, mover_id == AMPS_MOVER_GC_SECOND_ORDER ? 2 : 1, ctx->cfg.coupler_interpolation, ctx->cfg.ideal_mhd,
                               ctx->cfg.internal_sphere_radius, ctx->cfg.exit_record_capacity, ctx->buf[ctx->cur], ctx->d_n + ctx->cur, ctx->nUpper,
                               ctx->d_bgTile, ctx->d_gradBTile, ctx->d_bgE, ctx->d_bgB, ctx->d_gradBVar, ctx->d_cellCount, ctx->d_stats, ctx->d_exitBuf, ctx->d_exitCount, ctx->stream,
                               gcEcsim ? ctx->d_E : nullptr, gcEcsim ? ctx->d_Bcur : nullptr);
    ctx->launches++;
    CK(cudaGetLastError());
    ctx->sorted = false;
    ctx->countValid = true;
    return AMPS_GPU_OK;
  }
  if (mover_id == AMPS_MOVER_GYROKINETIC_FIRST_ORDER || mover_id == AMPS_MOVER_GYROKINETIC_SECOND_ORDER) {
    launch_move_gyrokinetic(m, ctx->sp, mover_id == AMPS_MOVER_GYROKINETIC_SECOND_ORDER ? 2 : 1, ctx->cfg.coupler_interpolation,
                            ctx->cfg.internal_sphere_radius, ctx->buf[ctx->cur], ctx->d_n + ctx->cur, ctx->nUpper, ctx->d_bgTile, ctx->d_gradBTile,
                            ctx->d_bgE, ctx->d_bgB, ctx->d_gradBVar, ctx->d_cellCount, ctx->d_stats, ctx->stream);
    ctx->launches++;
    CK(cudaGetLastError());
    ctx->sorted = false;
    ctx->countValid = true;
    return AMPS_GPU_OK;
  }
  if (mover_id == AMPS_MOVER_RELATIVISTIC_GCA) {
    launch_move_relativistic_gca(m, ctx->sp, ctx->cfg.coupler_interpolation, ctx->cfg.speed_of_light, ctx->cfg.internal_sphere_radius,
                                 ctx->cfg.exit_record_capacity, ctx->buf[ctx->cur], ctx->d_n + ctx->cur, ctx->nUpper, ctx->d_bgTile, ctx->d_gcaTile,
                                 ctx->d_bgE, ctx->d_bgB, ctx->d_gcaVar, ctx->d_cellCount, ctx->d_stats, ctx->d_exitBuf, ctx->d_exitCount, ctx->stream);
    ctx->launches++;
    CK(cudaGetLastError());
    ctx->sorted = false;
    ctx->countValid = true;
    return AMPS_GPU_OK;
  }
  if (mover_id == AMPS_MOVER_RELATIVISTIC_BORIS) {
    launch_move_relativistic_boris(m, ctx->sp, ctx->cfg.coupler_interpolation, ctx->cfg.backward_time_integration, ctx->cfg.speed_of_light,
                                   ctx->cfg.internal_sphere_radius, ctx->cfg.exit_record_capacity, ctx->buf[ctx->cur], ctx->d_n + ctx->cur, ctx->nUpper,
                                   ctx->d_bgTile, ctx->d_bgE, ctx->d_bgB, ctx->d_cellCount, ctx->d_stats, ctx->d_exitBuf, ctx->d_exitCount,
                                   ctx->stream);
    ctx->launches++;
    CK(cudaGetLastError());
    ctx->sorted = false;
    ctx->countValid = true;
    return AMPS_GPU_OK;
  }
  long long perLeaf = ctx->nUpper / (m.nLeaves > 0 ? m.nLeaves : 1);
  int slices = (int)((perLeaf + 4095) / 4096);
  if (slices < 1) slices = 1;
  if (slices > 64) slices = 64;
  if (ctx->cfg.exact_arithmetic) {
    launch_move_lapenta(m, ctx->sp, ctx->buf[ctx->cur], ctx->d_cellStart, ctx->d_eTile, ctx->d_bPrevTile, ctx->d_cellCount, ctx->d_stats, slices,
                        ctx->d_exitBuf, ctx->d_exitCount, ctx->cfg.exit_record_capacity, nullptr, nullptr, nullptr, ctx->stream);
    ctx->launches++;
  } else {
    // fast pass (FMA, reciprocals) for every particle that stays clear of cell faces; the exact kernel finishes the flagged rest
    int rc;
    if (!ctx->d_redoMask && (rc = dev_alloc(ctx, &ctx->d_redoMask, (size_t)ctx->cfg.capacity))) return rc;
    if (!ctx->d_leafRedo && (rc = dev_alloc(ctx, &ctx->d_leafRedo, (size_t)2 * m.nLeaves + 1))) return rc;  // (per mesh epoch)
    CK(cudaMemsetAsync(ctx->d_redoMask, 0, (size_t)ctx->nUpper, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_leafRedo, 0, sizeof(int) * ((size_t)2 * m.nLeaves + 1), ctx->stream));
    int *redoList = ctx->d_leafRedo + m.nLeaves, *nRedoLeaves = ctx->d_leafRedo + 2 * (size_t)m.nLeaves;
    launch_move_lapenta_fast(m, ctx->sp, ctx->buf[ctx->cur], ctx->d_cellStart, ctx->d_eTile, ctx->d_bPrevTile, ctx->d_cellCount, ctx->d_stats, slices,
                             ctx->d_redoMask, ctx->d_leafRedo, redoList, nRedoLeaves, ctx->stream);
    launch_move_lapenta(m, ctx->sp, ctx->buf[ctx->cur], ctx->d_cellStart, ctx->d_eTile, ctx->d_bPrevTile, ctx->d_cellCount, ctx->d_stats, slices,
                        ctx->d_exitBuf, ctx->d_exitCount, ctx->cfg.exit_record_capacity, ctx->d_redoMask, redoList, nRedoLeaves, ctx->stream);
    ctx->launches += 2;
  }
  CK(cudaGetLastError());
  ctx->sorted = false;
  ctx->countValid = true;
  return AMPS_GPU_OK;
}

int amps_gpu_move(amps_gpu_ctx *ctx, int mover_id, amps_gpu_move_stats *stats) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  int rc = do_move(ctx, mover_id);
  if (rc) return rc;
  if (stats) {
    DevMoveStats h;
    CK(cudaMemcpyAsync(&h, ctx->d_stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    stats->n_moved = (int64_t)h.n_moved;
    stats->n_cross_cell = (int64_t)h.n_cross_cell;
    stats->n_cross_block = (int64_t)h.n_cross_block;
    stats->n_left_domain = (int64_t)h.n_left_domain;
    stats->n_not_in_use = (int64_t)h.n_not_in_use;
    stats->n_periodic_wrap = (int64_t)h.n_periodic_wrap;
    stats->n_error = (int64_t)h.n_error;
    stats->n_sub_steps = (int64_t)h.n_sub_steps;
    if (h.n_error) FAIL(AMPS_GPU_ERR_PARTICLE, "mover: particle outside its block / cell not found (the reference would exit())");
  }
  return AMPS_GPU_OK;
}

static int do_deposit(amps_gpu_ctx *ctx) {
  if (!ctx->meshReady || !ctx->fieldsReady) FAIL(AMPS_GPU_ERR_STATE, "deposit before mesh/fields upload");
  if (!ctx->sorted) FAIL(AMPS_GPU_ERR_STATE, "deposit needs the (block,cell)-sorted layout: call amps_gpu_sort");
  if (ctx->meshRefined && ctx->cfg.b_mode == AMPS_B_CENTER_BASED)
    FAIL(AMPS_GPU_ERR_STATE, "ECSIM on a refined mesh needs _PIC_FIELD_SOLVER_B_CORNER_BASED_ (see amps_gpu_move)");
  {
    int rcGc;
    if ((rcGc = gc_check(ctx))) return rcGc;
  }
  ProfScope prof(ctx, AMPS_GPU_PHASE_DEPOSIT);
  launch_deposit(ctx->dm, deposit_species(ctx), ctx->buf[ctx->cur], ctx->d_cellStart, ctx->d_bCurTile, ctx->d_J, ctx->d_M, ctx->d_energy, ctx->d_cfl,
                 ctx->nSM, nullptr, ctx->buf[ctx->cur], 0, -1, DEP_ALL | gc_dep_flags(ctx), ctx->stream, &ctx->launches);
  {
    int rcGc;
    if ((rcGc = gc_pass(ctx, ctx->buf[ctx->cur]))) return rcGc;
  }
  CK(cudaGetLastError());
  return AMPS_GPU_OK;
}

int amps_gpu_deposit_JM(amps_gpu_ctx *ctx, double *particle_energy, double *cfl) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  int rc = do_deposit(ctx);
  if (rc) return rc;
  if (particle_energy || cfl) {
    double e;
    unsigned long long c[AMPS_GPU_MAX_SPECIES];
    CK(cudaMemcpyAsync(&e, ctx->d_energy, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(c, ctx->d_cfl, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (particle_energy) *particle_energy = e;
    if (cfl)
      for (int s = 0; s < ctx->cfg.n_species; s++) memcpy(&cfl[s], &c[s], 8);
  }
  return AMPS_GPU_OK;
}

int amps_gpu_sample_cells(amps_gpu_ctx *ctx) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "sample_cells before mesh_upload");
  if (!ctx->sorted) FAIL(AMPS_GPU_ERR_STATE, "sample_cells needs the (block,cell)-sorted layout: call amps_gpu_sort");
  CK(cudaSetDevice(ctx->cfg.device));
  int rc;
  const size_t n = (size_t)ctx->nCells * ctx->sp.n * 13;
  if (!ctx->d_sample) {
    if ((rc = dev_alloc(ctx, &ctx->d_sample, n))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_nSampled, (size_t)AMPS_GPU_MAX_SPECIES))) return rc;
    CK(cudaMemsetAsync(ctx->d_sample, 0, sizeof(double) * n, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_nSampled, 0, sizeof(unsigned long long) * AMPS_GPU_MAX_SPECIES, ctx->stream));
  }
  launch_sample_cells(ctx->dm, ctx->sp, ctx->buf[ctx->cur], ctx->d_cellStart, ctx->d_sample, ctx->d_nSampled, ctx->nSM, ctx->stream);
  ctx->launches++;
  CK(cudaGetLastError());
  return AMPS_GPU_OK;
}

int amps_gpu_sample_download(amps_gpu_ctx *ctx, double *sample, int64_t *n_sampled, int clear) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  if (!ctx->d_sample) FAIL(AMPS_GPU_ERR_STATE, "sample_download before amps_gpu_sample_cells");
  CK(cudaSetDevice(ctx->cfg.device));
  const size_t n = (size_t)ctx->nCells * ctx->sp.n * 13;
  unsigned long long cnt[AMPS_GPU_MAX_SPECIES];
  if (sample) CK(cudaMemcpyAsync(sample, ctx->d_sample, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(cnt, ctx->d_nSampled, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
  if (clear) {
    CK(cudaMemsetAsync(ctx->d_sample, 0, sizeof(double) * n, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_nSampled, 0, sizeof(cnt), ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  if (n_sampled)
    for (int s = 0; s < ctx->sp.n; s++) n_sampled[s] = (int64_t)cnt[s];
  return AMPS_GPU_OK;
}

int amps_gpu_species_moments(amps_gpu_ctx *ctx, double *spec_corner) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "species_moments before mesh_upload");
  if (!ctx->sorted) FAIL(AMPS_GPU_ERR_STATE, "species_moments needs the (block,cell)-sorted layout: call amps_gpu_sort");
  CK(cudaSetDevice(ctx->cfg.device));
  int rc;
  const size_t n = (size_t)ctx->dm.nCorners * 10 * ctx->sp.n;
  if (!ctx->d_spec && (rc = dev_alloc(ctx, &ctx->d_spec, n))) return rc;
  launch_species_moments(ctx->dm, ctx->sp, ctx->buf[ctx->cur], ctx->d_cellStart, ctx->d_spec, ctx->nSM, ctx->stream);
  ctx->launches++;
  CK(cudaGetLastError());
  ctx->specReady = true;
  if (spec_corner) {
    CK(cudaMemcpyAsync(spec_corner, ctx->d_spec, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return AMPS_GPU_OK;
}

int amps_gpu_phi_upload(amps_gpu_ctx *ctx, const double *phi_center) {
  if (!ctx || !phi_center) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "phi_upload before mesh_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  int rc;
  if (!ctx->d_phi && (rc = dev_alloc(ctx, &ctx->d_phi, (size_t)ctx->dm.nCenters))) return rc;
  CK(cudaMemcpyAsync(ctx->d_phi, phi_center, sizeof(double) * (size_t)ctx->dm.nCenters, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->phiReady = true;
  return AMPS_GPU_OK;
}

int amps_gpu_correct_particle_location(amps_gpu_ctx *ctx, double charge_conv, double mass_conv, int64_t *n_displaced, int64_t *n_deleted) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "correct_particle_location before mesh_upload");
  if (!ctx->specReady || !ctx->phiReady) FAIL(AMPS_GPU_ERR_STATE, "correct_particle_location needs amps_gpu_species_moments and amps_gpu_phi_upload");
  if (ctx->meshRefined) FAIL(AMPS_GPU_ERR_STATE, "CorrectParticleLocation is defined on single-level meshes (the reference indexes a block-local array)");
  CK(cudaSetDevice(ctx->cfg.device));
  int rc;
  if (!ctx->d_cplCount && (rc = dev_alloc(ctx, &ctx->d_cplCount, 3))) return rc;
  const double qom0 = (ctx->cfg.charge[0] * charge_conv) / (ctx->cfg.mass[0] * mass_conv);  // :4449-4452
  launch_correct_particle_location(ctx->dm, ctx->sp, ctx->buf[ctx->cur], ctx->d_n + ctx->cur, ctx->nUpper, ctx->d_phi, ctx->d_spec, ctx->d_neibMask,
                                   qom0, ctx->d_cellCount, ctx->d_cplCount, ctx->stream);
  ctx->launches++;
  CK(cudaGetLastError());
  unsigned long long c[3];
  CK(cudaMemcpyAsync(c, ctx->d_cplCount, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->sorted = false;       // the lists of the reference are rebuilt (exchangeParticleLocal): amps_gpu_sort re-files the store
  ctx->countValid = true;    // like a mover, the pass leaves the cell histogram of the survivors (amps_gpu_migrate edits it)
  if (n_displaced) *n_displaced = (int64_t)c[0];
  if (n_deleted) *n_deleted = (int64_t)c[1];
  if (c[2]) FAIL(AMPS_GPU_ERR_PARTICLE, "CorrectParticleLocation: cannot find the cell where a particle is located (the reference exits here)");
  return AMPS_GPU_OK;
}

int amps_gpu_net_charge(amps_gpu_ctx *ctx, double charge_conv, double *rho_center) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "net_charge before mesh_upload");
  if (!ctx->sorted) FAIL(AMPS_GPU_ERR_STATE, "net_charge needs the (block,cell)-sorted layout: call amps_gpu_sort");
  if (ctx->meshRefined) FAIL(AMPS_GPU_ERR_STATE, "ComputeNetCharge is defined on single-level meshes (the reference indexes a block-local array)");
  if ((size_t)ctx->dm.nCenterLocal * sizeof(double) > 48 * 1024) FAIL(AMPS_GPU_ERR_ARG, "block too large for the shared-memory centre tile");
  CK(cudaSetDevice(ctx->cfg.device));
  int rc;
  if (!ctx->d_rho && (rc = dev_alloc(ctx, &ctx->d_rho, (size_t)ctx->dm.nCenters))) return rc;
  launch_net_charge(ctx->dm, ctx->sp, ctx->buf[ctx->cur], ctx->d_cellStart, charge_conv, ctx->d_rho, ctx->nUpper, ctx->stream);
  ctx->launches++;
  CK(cudaGetLastError());
  if (rho_center) CK(cudaMemcpyAsync(rho_center, ctx->d_rho, sizeof(double) * (size_t)ctx->dm.nCenters, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AMPS_GPU_OK;
}

int amps_gpu_diagnostics(amps_gpu_ctx *ctx, double *particle_energy, double *cfl) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  double e;
  unsigned long long c[AMPS_GPU_MAX_SPECIES];
  CK(cudaMemcpyAsync(&e, ctx->d_energy, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(c, ctx->d_cfl, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (particle_energy) *particle_energy = e;
  if (cfl)
    for (int s = 0; s < ctx->cfg.n_species; s++) memcpy(&cfl[s], &c[s], 8);
  return AMPS_GPU_OK;
}

int amps_gpu_JM_download(amps_gpu_ctx *ctx, double *J, double *M) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "JM_download before mesh_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  if (J) CK(cudaMemcpyAsync(J, ctx->d_J, sizeof(double) * 3 * (size_t)ctx->dm.nCorners, cudaMemcpyDeviceToHost, ctx->stream));
  if (M) CK(cudaMemcpyAsync(M, ctx->d_M, sizeof(double) * 243 * (size_t)ctx->dm.nCorners, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AMPS_GPU_OK;
}

int amps_gpu_JM_device(amps_gpu_ctx *ctx, double **J_dev, double **M_dev) {
  if (!ctx || !ctx->meshReady) return AMPS_GPU_ERR_STATE;
  if (J_dev) *J_dev = ctx->d_J;
  if (M_dev) *M_dev = ctx->d_M;
  return AMPS_GPU_OK;
}

int amps_gpu_profile(amps_gpu_ctx *ctx, int enable, double *phase_ms, int64_t *phase_count) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaStreamSynchronize(ctx->stream));
  for (auto &sp : ctx->evSpans) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->evPool[sp.second.first], ctx->evPool[sp.second.second]);
    if (sp.first >= AMPS_GPU_N_PHASES) {  // sub-phases of the exchange (AMPS_GPU_DEBUG_SUBPHASES)
      ctx->subMs[sp.first] += ms;
      continue;
    }
    ctx->phaseMs[sp.first] += ms;
    ctx->phaseCount[sp.first]++;
  }
  if (!ctx->subMs.empty()) {
    if (ctx->rank == 0) {
      static const char *names[] = {"mig_pack", "mig_counts", "mig_sendrecv", "mig_unpack", "jm_pack", "jm_sendrecv", "jm_add", "jm_allreduce"};
      for (auto &kv : ctx->subMs) fprintf(stderr, "[amps_gpu] sub-phase %s: %.3f ms total\n", names[(kv.first - 16) & 7], kv.second);
    }
    ctx->subMs.clear();
  }
  ctx->evSpans.clear();
  ctx->evUsed = 0;
  for (int i = 0; i < AMPS_GPU_N_PHASES; i++) {
    if (phase_ms) phase_ms[i] = ctx->phaseMs[i];
    if (phase_count) phase_count[i] = ctx->phaseCount[i];
    ctx->phaseMs[i] = 0, ctx->phaseCount[i] = 0;
  }
  ctx->profile = enable != 0;
  return AMPS_GPU_OK;
}

int amps_gpu_selftest_division(amps_gpu_ctx *ctx, const double *a, const double *b, int64_t n, int64_t *n_mismatch) {
  if (!ctx || !a || !b || !n_mismatch || n < 1) return AMPS_GPU_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  double *da = nullptr, *db = nullptr;
  unsigned long long *dout = nullptr, h = 0;
  CK(cudaMalloc(&da, n * sizeof(double)));
  CK(cudaMalloc(&db, n * sizeof(double)));
  CK(cudaMalloc(&dout, sizeof(unsigned long long)));
  CK(cudaMemcpyAsync(da, a, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(db, b, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(dout, 0, sizeof(unsigned long long), ctx->stream));
  launch_division_selftest(da, db, (int)n, dout, ctx->stream);
  ctx->launches++;
  CK(cudaMemcpyAsync(&h, dout, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(da), cudaFree(db), cudaFree(dout);
  *n_mismatch = (int64_t)h;
  return AMPS_GPU_OK;
}

int amps_gpu_comm_unique_id(void *id128) {
  if (!id128) return AMPS_GPU_ERR_ARG;
  NcclApi &a = nccl_api();
  if (!a.ok) return AMPS_GPU_ERR_STATE;
  ncclUniqueId id;
  if (a.GetUniqueId(&id) != ncclSuccess) return AMPS_GPU_ERR_CUDA;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return AMPS_GPU_OK;
}

int amps_gpu_comm_init(amps_gpu_ctx *ctx, const void *id128, int rank, int n_ranks) {
  if (!ctx || !id128) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "comm_init before mesh_upload");
  if (rank != ctx->rank || n_ranks != ctx->nRanks) FAIL(AMPS_GPU_ERR_ARG, "rank / n_ranks differ from the mesh description");
  NcclApi &a = nccl_api();
  if (!a.ok) FAIL(AMPS_GPU_ERR_STATE, "libnccl.so.2 could not be loaded");
  CK(cudaSetDevice(ctx->cfg.device));
  // the shared-corner exchange is a handful of 10-50 MB point-to-point messages per step: NCCL's default of 2 channels per
  // peer moves them at ~45 GB/s, 16 channels at ~170 GB/s over NVLink (measured, B200 x2).  NCCL reads the variables once per
  // process, at its first communicator: a host that initialises NCCL earlier (torch) sets them itself (bench.py does).
  setenv("NCCL_MIN_P2P_NCHANNELS", "16", 0);
  setenv("NCCL_MAX_P2P_NCHANNELS", "32", 0);
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NCK(a.CommInitRank(&ctx->comm, n_ranks, id, rank));
  // ---- peer memory for the particle migration (one box: every GPU reaches every other through NVLink / NVSwitch) ----
  // Every rank publishes the IPC handle of its receive buffer; a rank that can map all the others writes its leavers straight
  // into their buffers (pack_leavers_kernel), so the exchange needs no message sizes and the host never waits for counts.
  // All ranks take the same decision (an all-reduce of "could map everything"); otherwise the NCCL send/recv path stays.
  ctx->peerMigrate = false;
  const char *env = getenv("AMPS_GPU_PEER_MIGRATE");
  int want = (env && env[0] == '0') ? 0 : 1;
  {
    cudaIpcMemHandle_t mine;
    unsigned char *d_h = nullptr;
    int *d_ok = nullptr;
    std::vector<cudaIpcMemHandle_t> all((size_t)n_ranks);
    CK(cudaMalloc(&d_h, sizeof(cudaIpcMemHandle_t) * (size_t)(n_ranks + 1)));
    CK(cudaMalloc(&d_ok, sizeof(int)));
    if (cudaIpcGetMemHandle(&mine, ctx->d_recvBuf) != cudaSuccess) {
      cudaGetLastError();
      want = 0;
      memset(&mine, 0, sizeof(mine));
    }
    CK(cudaMemcpy(d_h + sizeof(mine) * (size_t)n_ranks, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    NCK(a.AllGather(d_h + sizeof(mine) * (size_t)n_ranks, d_h, sizeof(mine), ncclChar, ctx->comm, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(all.data(), d_h, sizeof(mine) * (size_t)n_ranks, cudaMemcpyDeviceToHost));
    ctx->h_peerMapped.assign((size_t)n_ranks, nullptr);
    if (want)
      for (int r = 0; r < n_ranks; r++) {
        if (r == rank) continue;
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          want = 0;
          break;
        }
        ctx->h_peerMapped[r] = ptr;
      }
    CK(cudaMemcpy(d_ok, &want, sizeof(int), cudaMemcpyHostToDevice));
    NCK(a.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, ctx->comm, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int okAll = 0;
    CK(cudaMemcpy(&okAll, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_h), cudaFree(d_ok);
    if (okAll) {
      std::vector<double *> tab((size_t)n_ranks, nullptr);
      for (int r = 0; r < n_ranks; r++) tab[r] = (r == rank) ? ctx->d_recvBuf : (double *)ctx->h_peerMapped[r];
      cudaFree(ctx->d_peerRecv);
      CK(cudaMalloc(&ctx->d_peerRecv, sizeof(double *) * (size_t)n_ranks));
      CK(cudaMemcpy(ctx->d_peerRecv, tab.data(), sizeof(double *) * (size_t)n_ranks, cudaMemcpyHostToDevice));
      if (!ctx->d_sentRecv) CK(cudaMalloc(&ctx->d_sentRecv, 2 * sizeof(long long)));
      if (!ctx->h_errLazy) {
        CK(cudaMallocHost(&ctx->h_errLazy, sizeof(int)));
        *ctx->h_errLazy = 0;
      }
      ctx->peerMigrate = true;
    } else {
      for (void *&q : ctx->h_peerMapped)
        if (q) cudaIpcCloseMemHandle(q), q = nullptr;
    }
  }
  return AMPS_GPU_OK;
}

int amps_gpu_comm_uses_peer_memory(amps_gpu_ctx *ctx) { return (ctx && ctx->peerMigrate) ? 1 : 0; }

int amps_gpu_set_shared_corners(amps_gpu_ctx *ctx, int peer, const int32_t *uids, int64_t n) {
  if (!ctx || peer < 0 || n < 0 || (n > 0 && !uids)) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady || ctx->nRanks <= 1 || peer >= ctx->nRanks || peer == ctx->rank) FAIL(AMPS_GPU_ERR_ARG, "bad peer");
  CK(cudaSetDevice(ctx->cfg.device));
  for (int64_t i = 0; i < n; i++)
    if (uids[i] < 0 || uids[i] >= ctx->dm.nCorners) FAIL(AMPS_GPU_ERR_ARG, "shared corner id out of range");
  cudaFree(ctx->d_sharedUid[peer]);
  ctx->d_sharedUid[peer] = nullptr;
  ctx->nShared[peer] = n;
  ctx->h_sharedUid[peer].assign(uids, uids + n);
  ctx->depDirty = true;
  ctx->sharedAllDirty = true;
  if (n) {
    CK(cudaMalloc(&ctx->d_sharedUid[peer], n * sizeof(int)));
    CK(cudaMemcpy(ctx->d_sharedUid[peer], uids, n * sizeof(int), cudaMemcpyHostToDevice));
  }
  long long tot = 0;
  for (long long v : ctx->nShared) tot += v;
  if (tot * 246 > ctx->cornerBufDoubles) {
    cudaFree(ctx->d_cornerSend), cudaFree(ctx->d_cornerRecv);
    ctx->cornerBufDoubles = tot * 246;
    CK(cudaMalloc(&ctx->d_cornerSend, ctx->cornerBufDoubles * sizeof(double)));
    CK(cudaMalloc(&ctx->d_cornerRecv, ctx->cornerBufDoubles * sizeof(double)));
  }
  return AMPS_GPU_OK;
}

static int do_migrate(amps_gpu_ctx *ctx, int64_t *n_sent, int64_t *n_received, bool zeroJM = false, bool barrierFollows = false) {
  if (n_sent) *n_sent = 0;
  if (n_received) *n_received = 0;
  if (ctx->nRanks <= 1) return AMPS_GPU_OK;
  if (!ctx->comm) FAIL(AMPS_GPU_ERR_STATE, "migrate before amps_gpu_comm_init");
  if (!ctx->countValid) FAIL(AMPS_GPU_ERR_STATE, "migrate must follow amps_gpu_move (it edits the mover's cell histogram)");
  ProfScope prof(ctx, AMPS_GPU_PHASE_EXCHANGE);
  NcclApi &a = nccl_api();
  const int R = ctx->nRanks, me = ctx->rank;
  const size_t recLen = (size_t)migration_record_len(ctx->buf[ctx->cur]);
  cudaStream_t s = ctx->stream;
  Sub s0(ctx, 16);
  CK(cudaMemsetAsync(ctx->d_sendCount, 0, sizeof(int) * R, s));
  launch_pack_leavers(ctx->dm, ctx->buf[ctx->cur], ctx->d_n + ctx->cur, ctx->nUpper, ctx->d_leafOwner, ctx->d_leafGlobal, me, ctx->d_sendBuf,
                      ctx->capPerPeer, ctx->d_sendCount, ctx->d_cellCount, ctx->d_errFlag, ctx->peerMigrate ? ctx->d_peerRecv : nullptr, s);
  ctx->launches++;
  s0.end();
  if (ctx->peerMigrate) {
    // the records already lie in the peers' receive buffers.  The all-gather of the counts is the only message: it announces
    // them and, being ordered behind every rank's pack kernel, doubles as the barrier before anybody reads its buffer.
    Sub sp1(ctx, 17);
    NCK(a.AllGather(ctx->d_sendCount, ctx->d_allCounts, R, ncclInt, ctx->comm, s));
    if (zeroJM) {
      CK(cudaMemsetAsync(ctx->d_J, 0, sizeof(double) * 3 * (size_t)ctx->dm.nCorners, s));
      CK(cudaMemsetAsync(ctx->d_M, 0, sizeof(double) * 243 * (size_t)ctx->dm.nCorners, s));
      ctx->jmZeroed = true;
    }
    sp1.end();
    Sub sp3(ctx, 19);
    launch_unpack_arrivals_peer(ctx->dm, ctx->d_recvBuf, ctx->d_allCounts, R, ctx->capPerPeer, ctx->buf[ctx->cur], ctx->d_n + ctx->cur, ctx->d_g2l,
                                ctx->d_leafOwner, me, ctx->cfg.capacity, ctx->d_cellCount, ctx->d_errFlag, ctx->d_sentRecv, s);
    // the next exchange may overwrite this rank's regions in the peers' buffers only after they have unpacked: the all-gather
    // of the NEXT step is behind their unpack kernel in their streams, and this rank's next pack kernel is behind ITS previous
    // all-gather only -- so a second, empty-payload barrier closes the exchange, unless one follows anyway (inside
    // amps_gpu_step the all-reduces of the shared-corner exchange come after every rank's unpack kernel)
    if (!barrierFollows) NCK(a.AllReduce(ctx->d_sendCount, ctx->d_sendCount, 1, ncclInt, ncclMax, ctx->comm, s));
    sp3.end();
    ctx->launches += 2;
    ctx->nUpper = ctx->cfg.capacity;  // the arrivals are counted on the device only; the next sort reports the population
    CK(cudaGetLastError());
    if (n_sent || n_received) {  // (a caller that asks pays the round trip)
      long long sr[2] = {0, 0};
      CK(cudaMemcpyAsync(sr, ctx->d_sentRecv, sizeof(sr), cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      if (n_sent) *n_sent = sr[0];
      if (n_received) *n_received = sr[1];
    }
    return AMPS_GPU_OK;
  }
  Sub s1(ctx, 17);
  // counts: every rank learns the whole R x R matrix (the reference's first message, pic_parallel.cpp:267-301)
  NCK(a.AllGather(ctx->d_sendCount, ctx->d_allCounts, R, ncclInt, ctx->comm, s));
  std::vector<int> all((size_t)R * R);
  int err = 0;
  CK(cudaMemcpyAsync(all.data(), ctx->d_allCounts, sizeof(int) * R * R, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&err, ctx->d_errFlag, sizeof(int), cudaMemcpyDeviceToHost, s));
  if (zeroJM) {
    // inside amps_gpu_step: the device zeroes J and M (SetCornerNodeAssociatedDataValue, :3266-3267) while the host waits for
    // the counts and enqueues the transfers
    if (!ctx->evCounts) CK(cudaEventCreateWithFlags(&ctx->evCounts, cudaEventDisableTiming));
    CK(cudaEventRecord(ctx->evCounts, s));
    CK(cudaMemsetAsync(ctx->d_J, 0, sizeof(double) * 3 * (size_t)ctx->dm.nCorners, s));
    CK(cudaMemsetAsync(ctx->d_M, 0, sizeof(double) * 243 * (size_t)ctx->dm.nCorners, s));
    ctx->jmZeroed = true;
    CK(cudaEventSynchronize(ctx->evCounts));
  } else {
    CK(cudaStreamSynchronize(s));
  }
  // the stream is idle here: the count of the last sort has arrived, the slot bound drops back to the resident population
  tighten_upper(ctx, true);
  if (err) CK(cudaMemsetAsync(ctx->d_errFlag, 0, sizeof(int), s));  // reported once, not sticky
  // Every rank holds the same R x R matrix, so every rank takes the same decision without another message: an overflowing
  // pair (a rank packed more leavers for one peer than a send region holds) fails the exchange on ALL ranks after a matched
  // transfer of the clamped counts -- no rank is left waiting in a receive nobody sends
  bool overflow = false;
  for (size_t i = 0; i < all.size(); i++)
    if (all[i] > ctx->capPerPeer) all[i] = (int)ctx->capPerPeer, overflow = true;
  long long nRecv = 0, nSend = 0;
  std::vector<long long> roff(R, 0);
  for (int r = 0; r < R; r++) {
    roff[r] = nRecv;
    if (r != me) nRecv += all[(size_t)r * R + me], nSend += all[(size_t)me * R + r];
  }
  s1.end();
  Sub s2(ctx, 18);
  NCK(a.GroupStart());
  for (int r = 0; r < R; r++) {
    if (r == me) continue;
    const int ns = all[(size_t)me * R + r], nr = all[(size_t)r * R + me];
    if (ns > 0) NCK(a.Send(ctx->d_sendBuf + (size_t)r * ctx->capPerPeer * recLen, (size_t)ns * recLen, ncclDouble, r, ctx->comm, s));
    if (nr > 0) NCK(a.Recv(ctx->d_recvBuf + (size_t)roff[r] * recLen, (size_t)nr * recLen, ncclDouble, r, ctx->comm, s));
  }
  NCK(a.GroupEnd());
  s2.end();
  // (reported after the matched transfer, so that the peers of this rank are not left in a receive)
  if (err & 2) FAIL(AMPS_GPU_ERR_STATE, "the previous exchange dropped arrivals (unknown leaf, foreign owner or no free slot): the mesh tables of the ranks disagree");
  if (overflow) FAIL(AMPS_GPU_ERR_CAPACITY, "migration send buffer overflow (more than capacity/32 leavers from one rank to another); every rank reports it");
  if (ctx->nUpper + nRecv > ctx->cfg.capacity) FAIL(AMPS_GPU_ERR_CAPACITY, "particle capacity exceeded by arriving particles");
  Sub s3(ctx, 19);
  launch_unpack_arrivals(ctx->dm, ctx->d_recvBuf, (int)nRecv, ctx->buf[ctx->cur], ctx->d_n + ctx->cur, ctx->d_g2l, ctx->d_leafOwner, me,
                         ctx->cfg.capacity, ctx->d_cellCount, ctx->d_errFlag, s);
  s3.end();
  if (nRecv) ctx->launches += 2;
  ctx->nUpper += nRecv;
  CK(cudaGetLastError());
  if (n_sent) *n_sent = nSend;
  if (n_received) *n_received = nRecv;
  return AMPS_GPU_OK;
}

int amps_gpu_migrate(amps_gpu_ctx *ctx, int64_t *n_sent, int64_t *n_received) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  return do_migrate(ctx, n_sent, n_received);
}

static int do_exchange_JM(amps_gpu_ctx *ctx) {
  if (ctx->nRanks <= 1) return AMPS_GPU_OK;
  if (!ctx->comm) FAIL(AMPS_GPU_ERR_STATE, "exchange_JM before amps_gpu_comm_init");
  ProfScope prof(ctx, AMPS_GPU_PHASE_EXCHANGE);
  NcclApi &a = nccl_api();
  const int R = ctx->nRanks, me = ctx->rank;
  // inside amps_gpu_step the shared corners are final once the boundary leaves are deposited (evBoundary): pack and
  // send/recv run on a second stream next to the deposit of the interior leaves, the main stream joins for the add
  const bool overlap = ctx->overlapJM;
  ctx->overlapJM = false;
  if (overlap) {
    if (!ctx->commStream) {
      int lo = 0, hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CK(cudaStreamCreateWithPriority(&ctx->commStream, cudaStreamNonBlocking, hi));
    }
    CK(cudaStreamWaitEvent(ctx->commStream, ctx->evBoundary, 0));
  }
  cudaStream_t s = overlap ? ctx->commStream : ctx->stream;
  // all partial sums are packed before any is added, so every sharer ends with the same total
  Sub j0(ctx, 20);
  if (ctx->sharedAllDirty) {
    std::vector<int> all;
    for (int r = 0; r < R; r++)
      if (r != me) all.insert(all.end(), ctx->h_sharedUid[r].begin(), ctx->h_sharedUid[r].end());
    cudaFree(ctx->d_sharedUidAll);
    ctx->d_sharedUidAll = nullptr;
    ctx->nSharedAll = (long long)all.size();
    if (!all.empty()) {
      CK(cudaMalloc(&ctx->d_sharedUidAll, all.size() * sizeof(int)));
      CK(cudaMemcpy(ctx->d_sharedUidAll, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    ctx->sharedAllDirty = false;
  }
  long long off = 0;
  std::vector<long long> offs(R, 0);
  for (int r = 0; r < R; r++) {
    offs[r] = off;
    if (r == me || ctx->nShared[r] == 0) continue;
    off += ctx->nShared[r];
  }
  if (ctx->nSharedAll > 0) {
    launch_pack_corners(ctx->d_sharedUidAll, (int)ctx->nSharedAll, ctx->d_J, ctx->d_M, ctx->d_cornerSend, s);
    ctx->launches++;
  }
  j0.end();
  Sub j1(ctx, 21);
  NCK(a.GroupStart());
  for (int r = 0; r < R; r++) {
    if (r == me || ctx->nShared[r] == 0) continue;
    NCK(a.Send(ctx->d_cornerSend + offs[r] * 246, (size_t)ctx->nShared[r] * 246, ncclDouble, r, ctx->comm, s));
    NCK(a.Recv(ctx->d_cornerRecv + offs[r] * 246, (size_t)ctx->nShared[r] * 246, ncclDouble, r, ctx->comm, s));
  }
  NCK(a.GroupEnd());
  j1.end();
  if (overlap) {
    CK(cudaEventRecord(ctx->evRecv, s));
    s = ctx->stream;
    CK(cudaStreamWaitEvent(s, ctx->evRecv, 0));
  }
  Sub j2(ctx, 22);
  if (ctx->nSharedAll > 0) {
    launch_add_corners_atomic(ctx->d_sharedUidAll, (int)ctx->nSharedAll, ctx->d_J, ctx->d_M, ctx->d_cornerRecv, s);
    ctx->launches++;
  }
  j2.end();
  Sub j3(ctx, 23);
  // MPI_Reduce(ParticleEnergy, SUM), MPI_Reduce(cfl, MAX): non-negative doubles order like their bit patterns
  NCK(a.AllReduce(ctx->d_energy, ctx->d_energy, 1, ncclDouble, ncclSum, ctx->comm, s));
  NCK(a.AllReduce(ctx->d_cfl, ctx->d_cfl, AMPS_GPU_MAX_SPECIES, ncclUint64, ncclMax, ctx->comm, s));
  j3.end();
  CK(cudaGetLastError());
  return AMPS_GPU_OK;
}

int amps_gpu_exchange_JM(amps_gpu_ctx *ctx) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  return do_exchange_JM(ctx);
}

int amps_gpu_step(amps_gpu_ctx *ctx, int mover_id);
// amps_gpu_step + amps_gpu_JM_download with the download pipelined behind the deposit
// neighbour slots kept by the packed rows: slot = sx + 3 sy + 9 sz with the per-dimension code 0 -> 0, -1 -> 1, +1 -> 2 (:625-637);
// self, then of every pair (d, -d) the one whose highest non-zero dimension (z, then y, then x) points to +1
static const int kPackedSlots[14] = {0, 2, 6, 7, 8, 18, 19, 20, 21, 22, 23, 24, 25, 26};
const int32_t *amps_gpu_JM_packed_slots(void) {
  static const int32_t s[14] = {0, 2, 6, 7, 8, 18, 19, 20, 21, 22, 23, 24, 25, 26};
  return s;
}

static int pack_JM_all(amps_gpu_ctx *ctx, double *JM_packed_host, cudaStream_t s) {
  int rc;
  if (ctx->meshRefined) FAIL(AMPS_GPU_ERR_STATE, "the packed J/M rows pair the neighbour slots of equal cells: single-level meshes only");
  if (!ctx->d_pack && (rc = dev_alloc(ctx, &ctx->d_pack, (size_t)ctx->dm.nCorners * 129))) return rc;
  launch_pack_jm_half(0, ctx->dm.nCorners, ctx->d_J, ctx->d_M, ctx->d_pack, kPackedSlots, s);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(JM_packed_host, ctx->d_pack, sizeof(double) * 129 * (size_t)ctx->dm.nCorners, cudaMemcpyDeviceToHost, s));
  return AMPS_GPU_OK;
}

int amps_gpu_JM_download_packed(amps_gpu_ctx *ctx, double *JM_packed_host) {
  if (!ctx || !JM_packed_host) return AMPS_GPU_ERR_ARG;
  if (!ctx->meshReady) FAIL(AMPS_GPU_ERR_STATE, "JM_download_packed before mesh_upload");
  CK(cudaSetDevice(ctx->cfg.device));
  int rc;
  if ((rc = pack_JM_all(ctx, JM_packed_host, ctx->stream))) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return AMPS_GPU_OK;
}

static int do_step_JM(amps_gpu_ctx *ctx, int mover_id, double *J_host, double *M_host, double *JM_packed_host);
int amps_gpu_step_JM(amps_gpu_ctx *ctx, int mover_id, double *J_host, double *M_host) {
  if (!ctx || !J_host || !M_host) return AMPS_GPU_ERR_ARG;
  return do_step_JM(ctx, mover_id, J_host, M_host, nullptr);
}
int amps_gpu_step_JM_packed(amps_gpu_ctx *ctx, int mover_id, double *JM_packed_host) {
  if (!ctx || !JM_packed_host) return AMPS_GPU_ERR_ARG;
  return do_step_JM(ctx, mover_id, nullptr, nullptr, JM_packed_host);
}
static int do_step_JM(amps_gpu_ctx *ctx, int mover_id, double *J_host, double *M_host, double *JM_packed_host) {
  CK(cudaSetDevice(ctx->cfg.device));
  int rc;
  const bool packed = JM_packed_host != nullptr;
  if (packed && ctx->meshRefined) FAIL(AMPS_GPU_ERR_STATE, "the packed J/M rows pair the neighbour slots of equal cells: single-level meshes only");
  if (ctx->nRanks > 1 || ctx->cfg.gc_species_mask) {  // shared corners change in the exchange / the guiding-centre pass adds to J: no early download
    if ((rc = amps_gpu_step(ctx, mover_id))) return rc;
    if (packed) return amps_gpu_JM_download_packed(ctx, JM_packed_host);
    return amps_gpu_JM_download(ctx, J_host, M_host);
  }
  if (packed && !ctx->d_pack && (rc = dev_alloc(ctx, &ctx->d_pack, (size_t)ctx->dm.nCorners * 129))) return rc;
  if ((rc = do_move(ctx, mover_id))) return rc;
  if (!ctx->meshReady || !ctx->fieldsReady) FAIL(AMPS_GPU_ERR_STATE, "deposit before mesh/fields upload");
  if (ctx->meshRefined && ctx->cfg.b_mode == AMPS_B_CENTER_BASED)
    FAIL(AMPS_GPU_ERR_STATE, "ECSIM on a refined mesh needs _PIC_FIELD_SOLVER_B_CORNER_BASED_ (see amps_gpu_move)");
  if (!ctx->d_perm && (rc = dev_alloc(ctx, &ctx->d_perm, (size_t)ctx->cfg.capacity))) return rc;
  const int nChunks = (int)ctx->dlCellEnd.size();
  if (!ctx->copyStream) CK(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
  while ((int)ctx->dlEvents.size() < nChunks) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->dlEvents.push_back(e);
  }
  ParticleSoA &src = ctx->buf[ctx->cur], &dst = ctx->buf[1 - ctx->cur];
  const bool dbg = getenv("AMPS_GPU_DEBUG_TIMELINE") != nullptr;
  std::vector<cudaEvent_t> tk, tc0, tc1;
  cudaEvent_t t00 = nullptr;
  if (dbg) {
    cudaEventCreate(&t00);
    cudaEventRecord(t00, ctx->stream);
  }
  {
    ProfScope prof(ctx, AMPS_GPU_PHASE_SORT);
    launch_sort(ctx->dm, src, dst, ctx->d_n + ctx->cur, ctx->d_cellCount, ctx->d_cellStart, ctx->d_cellFill, ctx->d_n + (1 - ctx->cur), ctx->nUpper,
                ctx->countValid, ctx->d_scanTmp, ctx->d_perm, ctx->stream, &ctx->launches);
    CK(cudaGetLastError());
  }
  {
    ProfScope prof(ctx, AMPS_GPU_PHASE_DEPOSIT);
    int c0 = 0;
    for (int k = 0; k < nChunks; k++) {
      if (dbg) {
        cudaEvent_t a, b, c;
        cudaEventCreate(&a), cudaEventCreate(&b), cudaEventCreate(&c);
        tk.push_back(a), tc0.push_back(b), tc1.push_back(c);
      }
      launch_deposit(ctx->dm, ctx->sp, src, ctx->d_cellStart, ctx->d_bCurTile, ctx->d_J, ctx->d_M, ctx->d_energy, ctx->d_cfl, ctx->nSM, ctx->d_perm,
                     dst, ctx->h_realBefore[c0 / ctx->dm.cellsPerBlock], ctx->h_realBefore[ctx->dlCellEnd[k] / ctx->dm.cellsPerBlock],
                     (k == 0 ? DEP_ZERO_JM | DEP_ZERO_DIAG | DEP_GHOST_PASS : 0u) | (k == nChunks - 1 ? (unsigned)DEP_FINAL : 0u), ctx->stream,
                     &ctx->launches);
      CK(cudaGetLastError());
      CK(cudaEventRecord(ctx->dlEvents[k], ctx->stream));
      if (dbg) cudaEventRecord(tk[k], ctx->stream);
      CK(cudaStreamWaitEvent(ctx->copyStream, ctx->dlEvents[k], 0));
      if (dbg) cudaEventRecord(tc0[k], ctx->copyStream);
      for (int r = ctx->dlRunStart[k]; r < ctx->dlRunStart[k + 1]; r++) {
        const size_t u0 = (size_t)ctx->dlRuns[r].uid0, n = (size_t)ctx->dlRuns[r].n;
        if (packed) {
          launch_pack_jm_half((int)u0, (int)n, ctx->d_J, ctx->d_M, ctx->d_pack, kPackedSlots, ctx->copyStream);
          ctx->launches++;
          CK(cudaMemcpyAsync(JM_packed_host + 129 * u0, ctx->d_pack + 129 * u0, sizeof(double) * 129 * n, cudaMemcpyDeviceToHost, ctx->copyStream));
        } else {
          CK(cudaMemcpyAsync(M_host + 243 * u0, ctx->d_M + 243 * u0, sizeof(double) * 243 * n, cudaMemcpyDeviceToHost, ctx->copyStream));
        }
      }
      if (dbg) cudaEventRecord(tc1[k], ctx->copyStream);
      c0 = ctx->dlCellEnd[k];
    }
    // J is 1% of the volume: one copy behind the last range (the packed rows carry it)
    if (!packed) CK(cudaMemcpyAsync(J_host, ctx->d_J, sizeof(double) * 3 * (size_t)ctx->dm.nCorners, cudaMemcpyDeviceToHost, ctx->copyStream));
  }
  ctx->cur = 1 - ctx->cur;
  ctx->sorted = true;
  ctx->countValid = false;
  if ((rc = request_sorted_count(ctx))) return rc;
  CK(cudaStreamSynchronize(ctx->copyStream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (dbg) {
    for (int k = 0; k < nChunks; k++) {
      float a, b, c;
      cudaEventElapsedTime(&a, t00, tk[k]), cudaEventElapsedTime(&b, t00, tc0[k]), cudaEventElapsedTime(&c, t00, tc1[k]);
      size_t bytes = 0;
      for (int r = ctx->dlRunStart[k]; r < ctx->dlRunStart[k + 1]; r++) bytes += (size_t)ctx->dlRuns[r].n * 243 * 8;
      fprintf(stderr, "range %d: kernel done %.3f ms, copy %.3f -> %.3f ms, %.1f MB in %d runs\n", k, a, b, c, bytes / 1e6,
              ctx->dlRunStart[k + 1] - ctx->dlRunStart[k]);
      cudaEventDestroy(tk[k]), cudaEventDestroy(tc0[k]), cudaEventDestroy(tc1[k]);
    }
    cudaEventDestroy(t00);
  }
  return AMPS_GPU_OK;
}

int amps_gpu_step(amps_gpu_ctx *ctx, int mover_id) {
  if (!ctx) return AMPS_GPU_ERR_ARG;
  CK(cudaSetDevice(ctx->cfg.device));
  int rc;
  ctx->jmZeroed = false;
  if ((rc = do_move(ctx, mover_id))) return rc;
  if ((rc = do_migrate(ctx, nullptr, nullptr, ctx->meshReady && ctx->fieldsReady, true))) return rc;
  rc = do_sort_deposit_fused(ctx);
  ctx->jmZeroed = false;
  if (rc) return rc;
  return do_exchange_JM(ctx);
}

}  // extern "C"
