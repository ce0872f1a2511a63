/*
 * amps_gpu.h -- C ABI of the B200 (sm_100a) charged-particle hot path of AMPS.
 *
 * Drop-in boundary for ONE path of SWMFsoftware/AMPS: the particle movers
 * (PIC::Mover::*) and the ECSIM current / mass-matrix deposition
 * (PIC::FieldSolver::Electromagnetic::ECSIM::UpdateJMassMatrix).  AMPS has no
 * plugin ABI of its own; the hooks these entry points sit behind are
 *
 *   PIC::Mover::UserDefinedMoverManager            src/pic/pic.h:5919-5920, pic_mover.cpp:589-592
 *   ECSIM::UpdateJMassMatrix (_CUDA_MODE_ branch)  src/pic/pic_field_solver_ecsim.cpp:3254-3257
 *   PIC::ParticleBuffer (AoS byte records)         src/pic/picParticleDataMacro.h:18-90
 *
 * (all reference paths are relative to the AMPS source tree).  See
 * INTEGRATION.md for the C++ shim a maintainer adds on the AMPS side.
 *
 * Conventions: plain C structs, host pointers unless the name says `_dev`,
 * every call returns an int status (AMPS_GPU_OK == 0); the library never calls
 * exit()/abort() (the reference does: exit(__LINE__,__FILE__,msg) -> MPI_Abort).
 * Device memory is owned by the context; host memory by the caller.  One caller
 * thread per context (the reference calls these paths from the rank's main
 * thread, src/pic/pic_time_step.cpp:403-551).
 */
#ifndef AMPS_GPU_H
#define AMPS_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMPS_GPU_MAX_SPECIES 8

/* ---- status codes ------------------------------------------------------- */
enum {
  AMPS_GPU_OK = 0,
  AMPS_GPU_ERR_CUDA = 1,          /* a CUDA runtime call failed; see amps_gpu_last_error() */
  AMPS_GPU_ERR_ARG = 2,           /* invalid argument / inconsistent description            */
  AMPS_GPU_ERR_CAPACITY = 3,      /* particle capacity exceeded                              */
  AMPS_GPU_ERR_STATE = 4,         /* call sequence error (e.g. move before mesh upload)      */
  AMPS_GPU_ERR_PARTICLE = 5,      /* device-side particle error: the reference would have
                                     called exit("cannot find the cell ...") pic_mover_boris.cpp:1286 */
  AMPS_GPU_ERR_NO_DEVICE = 6      /* no CUDA device: there is NO CPU fallback               */
};

/* ---- movers (PIC::Mover::*, bound by _PIC_PARTICLE_MOVER__MOVE_PARTICLE_TIME_STEP_) --- */
enum {
  AMPS_MOVER_LAPENTA2017 = 0,         /* pic_mover_boris.cpp:876-1393 (ECSIM push)             */
  AMPS_MOVER_BORIS = 1,               /* pic_mover_boris.cpp:126-553                           */
  AMPS_MOVER_RELATIVISTIC_BORIS = 2,  /* pic_mover_relativistic_boris.cpp:16-588               */
  AMPS_MOVER_GC_FIRST_ORDER = 3,      /* pic_mover_guiding_center.cpp:629-849                  */
  AMPS_MOVER_GC_SECOND_ORDER = 4,     /* pic_mover_guiding_center.cpp:293-627                  */
  AMPS_MOVER_RELATIVISTIC_GCA = 5,    /* pic_mover_relativistic_guiding_center.cpp:96-409      */
  AMPS_MOVER_MARKIDIS2010 = 6,        /* pic_mover_boris.cpp:557-835 (energy-conserving scheme, coupler fields) */
  AMPS_MOVER_GYROKINETIC_FIRST_ORDER = 7,  /* gyro/gyro_mover.cpp:386-544  PIC::GYROKINETIC::Mover_FirstOrder     */
  AMPS_MOVER_GYROKINETIC_SECOND_ORDER = 8  /* gyro/gyro_mover.cpp:544-720  PIC::GYROKINETIC::Mover_SecondOrder    */
};

/* return codes of a per-particle mover, src/pic/pic.h:5955-5960 */
enum {
  AMPS_PARTICLE_LEFT_THE_DOMAIN = 2,
  AMPS_PARTICLE_MOTION_FINISHED = 3,
  AMPS_PARTICLE_IN_NOT_IN_USE_NODE = 4
};

/* _PIC_PARTICLE_DOMAIN_BOUNDARY_INTERSECTION_PROCESSING_MODE_ */
enum {
  AMPS_BOUNDARY_DELETE = 0,
  AMPS_BOUNDARY_SPECULAR_REFLECTION = 1,
  AMPS_BOUNDARY_USER_FUNCTION = 2   /* exit record appended; host replays the callback */
};

/* _PIC_FIELD_SOLVER_B_MODE_ */
enum { AMPS_B_CENTER_BASED = 0, AMPS_B_CORNER_BASED = 1 };

/* _SIMULATION_TIME_STEP_MODE_ (pic_mover_boris.cpp:895-904) */
enum { AMPS_DT_SINGLE_GLOBAL = 0, AMPS_DT_SPECIES_GLOBAL = 1 };

/* node-flag bits of amps_gpu_mesh::node_flags */
enum {
  AMPS_NODE_USED = 1,           /* cTreeNodeAMR::IsUsedInCalculationFlag                     */
  AMPS_NODE_PERIODIC_GHOST = 2  /* cTreeNodeAMR::IsGhostNodeFlag (pic_bc_periodic.cpp:763-766) */
};

/* ---- configuration: every field is a compile-time macro or a namespace global
 *      in the reference (ampsConfig.pl rewrites them); here they are run-time. ---- */
typedef struct amps_gpu_config {
  int32_t block_cells[3];        /* _BLOCK_CELLS_X/Y/Z_                                        */
  int32_t ghost_cells[3];        /* _GHOST_CELLS_X/Y/Z_                                        */
  int32_t n_species;             /* PIC::nTotalSpecies                                         */
  int32_t b_mode;                /* AMPS_B_CENTER_BASED | AMPS_B_CORNER_BASED                  */
  int32_t periodic;              /* _PIC_BC__PERIODIC_MODE_                                    */
  int32_t boundary_mode;         /* AMPS_BOUNDARY_*                                            */
  int32_t time_step_mode;        /* AMPS_DT_*                                                  */
  int32_t device;                /* CUDA device ordinal                                        */
  int64_t capacity;              /* PIC::ParticleBuffer::MaxNPart (device SoA capacity)        */
  double charge[AMPS_GPU_MAX_SPECIES];   /* ElectricChargeTable after picunits::si2no_q when NORM units */
  double mass[AMPS_GPU_MAX_SPECIES];     /* MolMass after picunits::si2no_m when NORM units            */
  double species_weight[AMPS_GPU_MAX_SPECIES]; /* block->GetLocalParticleWeight(spec)                 */
  double time_step[AMPS_GPU_MAX_SPECIES];      /* PIC::ParticleWeightTimeStep::GlobalTimeStep[]       */
  /* ECSIM namespace globals, pic_field_solver_ecsim.cpp:186-212 */
  double ecsim_dt_total;         /* ECSIM::dtTotal                                             */
  double ecsim_B_conv;           /* ECSIM::B_conv                                              */
  double ecsim_length_conv;      /* ECSIM::length_conv                                         */
  double ecsim_light_speed;      /* ECSIM::LightSpeed                                          */
  /* test-particle movers (Boris / Relativistic::Boris / guiding centre): fields come from the coupler's
     centre-node table through PIC::CPLR::InitInterpolationStencil (pic_swmf.cpp:76-90)          */
  int32_t coupler_interpolation;     /* AMPS_CPLR_CELL_CENTERED_CONSTANT | _LINEAR  (_PIC_COUPLER__INTERPOLATION_MODE_) */
  int32_t backward_time_integration; /* BackwardTimeIntegrationMode (pic_mover_relativistic_boris.cpp:128-132)           */
  double speed_of_light;             /* SpeedOfLight (constants.h), SI                                                     */
  double internal_sphere_radius;     /* max(_RADIUS_(_TARGET_),Planet->Radius), 0 = no internal sphere (:272)            */
  int64_t exit_record_capacity;      /* records kept for the host callbacks (0 = only count)                              */
  double gravity_gm;                 /* GravityConstant*_MASS_(_TARGET_) of BorisSplitAcceleration_default (:110-118), 0 = off */
  int32_t carry_magnetic_moment;     /* _USE_MAGNETIC_MOMENT_: particles carry mu (picParticleDataMacro.h:178-187); needed by the GCA movers */
  int32_t exact_arithmetic;          /* 1: Lapenta2017 rounds every operation like the CPU build (no FMA contraction, IEEE quotients):
                                        x', v' bit-identical.  0 (default): contracted arithmetic for every particle whose x' stays clear
                                        of cell faces, the exact kernel for the rest -> keys/counters still bit-exact, x', v' to ~1e-14 */
  int32_t carry_v_parallel;          /* particles carry v_parallel (_PIC_PARTICLE_DATA__V_PARALLEL_OFFSET_, picParticleDataMacro.h): the reduced state
                                        of the gyrokinetic movers (needs carry_magnetic_moment as well)                     */
  int32_t ideal_mhd;                 /* _PIC__IDEAL_MHD_MODE_ (picGlobal.dfn:339, default ON): E.b = 0 in the guiding-centre parallel force */
  int32_t gc_species_mask;           /* bit s: species s is a guiding-centre species of PIC::GYROKINETIC (IsGuidingCenterSpecies, pic.h:5052) in
                                        ECSIM::ProcessCell (pic_field_solver_ecsim.cpp:2084): explicit current q v_eff, no mass matrix, the
                                        magnetisation current curl(M) of :1828 and |v_normal|^2 in the energy / cfl diagnostics.  Needs
                                        carry_magnetic_moment; one rank.  0 = _PIC_GYROKINETIC_MODEL_MODE_ off                       */
  int32_t gc_fields_ecsim;           /* 1: the guiding-centre movers (GC_FIRST_ORDER / GC_SECOND_ORDER) read the ECSIM arrays like the reference
                                        built with _PIC_FIELD_SOLVER_MODE__ELECTROMAGNETIC__ECSIM_ (pic_mover_guiding_center.cpp:103, :179-184,
                                        :727): E = the current E on the corners (amps_gpu_E_upload / amps_gpu_field_step) through the corner
                                        stencil, B = B_cur on the centres through the centre stencil, grad B = ECSIM::GetMagneticFieldGradient
                                        (pic_field_solver_ecsim.cpp:7473, differences over half a cell).  Single-level meshes, centre-based B.
                                        0: the coupler's background tables.
                                        NOTE: the guiding-centre movers take charge[] / mass[] as PIC::MolecularData::GetElectricCharge /
                                        GetMass, the RAW species tables (pic_mover_guiding_center.cpp:137, :216-217, :643), while Lapenta2017
                                        and ProcessCell use the picunits::si2no values: in a NORM-unit ECSIM run the context that moves the
                                        guiding-centre species is configured with the raw tables (tests/test_reference_gyrokinetic.py)  */
} amps_gpu_config;

/* _PIC_COUPLER__INTERPOLATION_MODE_ */
enum { AMPS_CPLR_CELL_CENTERED_CONSTANT = 0, AMPS_CPLR_CELL_CENTERED_LINEAR = 1 };

/* one particle that left through a domain face or hit the internal sphere: what
 * fProcessOutsideDomainParticles (pic.h:6043) / ParticleSphereInteraction (pic_mover_relativistic_boris.cpp:290)
 * receive; the host replays the callbacks (e.g. Earth::CutoffRigidity::ProcessOutsideDomainParticles,
 * srcEarth/CutoffRigidity.cpp:129-230)                                                        */
typedef struct amps_gpu_exit_record {
  int32_t ptr;       /* ParticleBuffer slot                                         */
  int32_t species;
  int32_t face;      /* nIntersectionFace 0..5, AMPS_EXIT_SPHERE for the internal sphere */
  int32_t leaf;      /* local leaf the callback gets as newNode                     */
  double x[3], v[3]; /* xInit, vInit at the intersection                            */
} amps_gpu_exit_record;
enum { AMPS_EXIT_SPHERE = 6 };

/* ---- flattened AMR mesh (host builds it from cMeshAMRgeneric; K7 in SURVEY 2.5) ----
 * The tree is a forest: n_root[0..2] root blocks tile [x_global_min,x_global_max];
 * n_root = {1,1,1} is exactly the reference's single octree
 * (src/meshAMR/meshAMRgeneric.h:2328-2393).  Node fields mirror cTreeNodeAMR
 * (meshAMRgeneric.h:825-838).                                                      */
typedef struct amps_gpu_mesh {
  int32_t n_root[3];
  int32_t max_refinement_level;     /* _MAX_REFINMENT_LEVEL_ (lattice depth, not deepest leaf) */
  double  x_global_min[3], x_global_max[3];
  double  dx_max_refinement[3];     /* meshAMRgeneric.h:2365                                   */
  double  dx_root_block[3];         /* dxRootBlock (per root block)                            */
  double  eps;                      /* cMeshAMRgeneric::EPS, meshAMRgeneric.h:2340             */
  int32_t n_nodes;
  const int32_t *node_parent;       /* [n_nodes]      upNode or -1                             */
  const int32_t *node_child;        /* [n_nodes][8]   downNode[i+2*(j+2*k)] or -1              */
  const int32_t *node_level;        /* [n_nodes]      RefinmentLevel                           */
  const int32_t *node_imin;         /* [n_nodes][3]   xMinGlobalIndex                          */
  const int32_t *node_isize;        /* [n_nodes]      NodeGeometricSizeIndex                   */
  const double  *node_xmin;         /* [n_nodes][3]                                            */
  const double  *node_xmax;         /* [n_nodes][3]                                            */
  const int32_t *node_leaf;         /* [n_nodes]      leaf id or -1                            */
  const int32_t *node_flags;        /* [n_nodes]      AMPS_NODE_*                              */
  const int32_t *node_thread;       /* [n_nodes]      cTreeNodeAMR::Thread (owner rank)        */
  const int32_t *root_node;         /* [n_root0*n_root1*n_root2] node id of root (i+n0*(j+n1*k)) */
  int32_t n_leaves;
  const int32_t *leaf_node;         /* [n_leaves]     tree node of the leaf                    */
  const int32_t *leaf_real;         /* [n_leaves]     periodic: paired real leaf of a ghost leaf, else -1
                                                      (BlockPairTable, pic_bc_periodic.cpp:555-571) */
  const int32_t *leaf_face_boundary;/* [n_leaves]     bit f set: GetNeibFace(f)==NULL          */
  /* unique corner / centre nodes; block-local number = _getCornerNodeLocalNumber /
     _getCenterNodeLocalNumber incl. ghost layers (meshAMRgeneric.h:74-75); -1 = no node */
  int32_t n_corners, n_centers;
  const int32_t *leaf_corner_uid;   /* [n_leaves][(Nx+2g+1)(Ny+2g+1)(Nz+2g+1)]                 */
  const int32_t *leaf_center_uid;   /* [n_leaves][(Nx+2g)(Ny+2g)(Nz+2g)]                       */
  /* domain decomposition (one rank per GPU).  n_ranks <= 1: the three tables may be NULL.     */
  int32_t this_rank, n_ranks;       /* PIC::ThisThread, PIC::nTotalThreads                     */
  int32_t n_global_leaves;          /* leaves of the whole tree (same numbering on every rank) */
  const int32_t *leaf_owner;        /* [n_leaves]  cTreeNodeAMR::Thread of each LOCAL leaf: own blocks, the boundary
                                       layer (DomainBoundaryLayerNodesList, meshAMRgeneric.h:1812) and real images of
                                       adjacent periodic ghost blocks                           */
  const int32_t *leaf_global_id;    /* [n_leaves]  global leaf number of each local leaf        */
  const int32_t *global_leaf_to_local; /* [n_global_leaves] local leaf or -1                    */
} amps_gpu_mesh;

/* AoS record description of PIC::ParticleBuffer (picParticleDataMacro.h:18-330) */
typedef struct amps_gpu_aos_layout {
  int64_t stride;           /* ParticleDataLength                                             */
  int32_t off_species;      /* _PIC_PARTICLE_DATA__SPECIES_ID_OFFSET_ (low 6 bits = id)       */
  int32_t off_v;            /* _PIC_PARTICLE_DATA__VELOCITY_OFFSET_                           */
  int32_t off_x;            /* _PIC_PARTICLE_DATA__POSITION_OFFSET_                           */
  int32_t off_w;            /* _PIC_PARTICLE_DATA__WEIGHT_CORRECTION_OFFSET_ or -1            */
  int32_t off_mu;           /* _PIC_PARTICLE_DATA__MAGNETIC_MOMENT_OFFSET_ or -1              */
  int32_t off_next;         /* _PIC_PARTICLE_DATA__NEXT_OFFSET_ (download rebuilds the lists) */
  int32_t off_prev;
  int32_t off_vpar;         /* _PIC_PARTICLE_DATA__V_PARALLEL_OFFSET_ or -1 (gyrokinetic movers; v_normal and the stored drift
                               velocity are functions of the state and are not carried)          */
} amps_gpu_aos_layout;

/* counters of one MoveParticles() call */
typedef struct amps_gpu_move_stats {
  int64_t n_moved;          /* particles processed                                            */
  int64_t n_cross_cell;     /* ended in another cell of the same block                        */
  int64_t n_cross_block;    /* ended in another block (incl. periodic wrap)                   */
  int64_t n_left_domain;    /* _PARTICLE_LEFT_THE_DOMAIN_                                     */
  int64_t n_not_in_use;     /* _PARTICLE_IN_NOT_IN_USE_NODE_                                  */
  int64_t n_periodic_wrap;  /* landed in a periodic ghost block and was shifted               */
  int64_t n_error;          /* reference would have exit()ed                                  */
  int64_t n_sub_steps;      /* Relativistic::Boris: gyro-period sub-steps taken (pic_mover_relativistic_boris.cpp:115-125); 0 for the
                               movers that do not sub-cycle                                                          */
} amps_gpu_move_stats;

typedef struct amps_gpu_ctx amps_gpu_ctx;

/* ---- life cycle ---------------------------------------------------------- */
/* <- PIC::Init_AfterParser / PIC::ParticleBuffer::Init (pic_pbuffer.cpp:41-222) */
int amps_gpu_init(const amps_gpu_config *cfg, amps_gpu_ctx **out);
int amps_gpu_finalize(amps_gpu_ctx *ctx);
const char *amps_gpu_last_error(const amps_gpu_ctx *ctx);
/* number of kernels this context has launched so far (bench.py "gpu_launches") */
int64_t amps_gpu_launch_count(const amps_gpu_ctx *ctx);
/* diagnostic: particles the last Lapenta2017 move handed from the contracted-arithmetic kernel to the exact one
 * (0 with exact_arithmetic = 1 or before the first move)                                          */
int amps_gpu_last_move_redo(amps_gpu_ctx *ctx, int64_t *n);
/* cudaStream_t the context launches on (as void*) */
void *amps_gpu_stream(amps_gpu_ctx *ctx);

/* ---- mesh: <- DomainBlockDecomposition::UpdateBlockTable (pic_mesh.cpp:1644),
 *      PIC::Mesh::GPU::CopyMeshHost2Device (pic.h:4794-4860)                  ----
 * Calling it again on a live context starts a new mesh epoch (the reference rebuilds BlockTable whenever
 * nMeshModificationCounter changed): everything sized by the old mesh is released, and the resident particles, fields,
 * background tables and shared-corner lists are dropped with it (their keys and node ids named the old blocks) - the host
 * uploads them again.  The communicator, the capacity and the species tables of amps_gpu_init stay.                       */
int amps_gpu_mesh_upload(amps_gpu_ctx *ctx, const amps_gpu_mesh *mesh);

/* ---- fields: replaces the per-thread SetBlock_E / SetBlock_B gathers
 *      (pic_mover.cpp:86-166).  Inputs are per UNIQUE node, 3 doubles each:
 *      E_half = corner OffsetE_HalfTimeStep, B_prev/B_cur = centre (or corner when
 *      b_mode is corner based) PrevBOffset/CurrentBOffset
 *      (pic_field_solver_ecsim.cpp:484-538).  Any pointer may be NULL = keep.   ---- */
int amps_gpu_fields_upload(amps_gpu_ctx *ctx, const double *E_half, const double *B_prev,
                           const double *B_cur);

/* ---- the field half of the ECSIM step on the device (single-level mesh, centre-based B,
 * normalised units).  Replaces PIC::FieldSolver::Electromagnetic::ECSIM::TimeStep (pic_field_solver_ecsim.cpp:6004-6157):
 * UpdateRhs (ecsim/update_rhs.cpp), UpdateMatrixElement (:6004-6020), cLinearSystemCornerNode::Solve / MultiplyVector
 * (srcInterface/LinearSystemCornerNode.h:3212, :2749; GMRES of the SWMF library), ProcessFinalSolution, UpdateB (:5160),
 * UpdateE (:5909).  J and M stay on the device: the 2 KB per corner the host path downloads every step never cross PCIe.
 * One rank, or several with amps_gpu_field_halo_set (every Krylov vector is refreshed on the neighbours' copies after each
 * product; the inner products are all-reduced).
 *
 * field_solver_init: node adjacency on the unique nodes (the host derives it from the leaves' node tables, like the row
 * builder GetStencil walks GetCornerNode(i+di, j+dj, k+dk)):
 *   corner_nb[n_corners][27]      neighbour corner per mass-matrix slot sx + 3 sy + 9 sz (0 -> 0, -1 -> 1, +1 -> 2); -1 = none:
 *                                 such a corner gets the boundary row dE = 0
 *   corner_cells[n_corners][8]    centre node of the cell at corner index + (a, b, c), a, b, c in {-1, 0}: (a+1) + 2 (b+1) + 4 (c+1)
 *   center_corners[n_centers][8]  corner node at cell index + (ii, jj, kk) in {0, 1}^3: ii + 2 jj + 4 kk                      */
int amps_gpu_field_solver_init(amps_gpu_ctx *ctx, const int32_t *corner_nb, const int32_t *corner_cells,
                               const int32_t *center_corners);
/* Several ranks (after amps_gpu_comm_init): the node values that cross rank boundaries in the field solve -- the reference's
 * PIC::Parallel::UpdateGhostBlockData (pic_bc_periodic.cpp:298-303) / ParallelBlockDataExchange for E and B.  corner_send / corner_recv:
 * local unique corner ids whose value this rank sends to / receives from `peer` after every operator product (the sender is the
 * corner's PRIMARY rank: the lowest rank that deposits into it); center_send / center_recv: the same for B after UpdateB (the
 * sender owns the cell).  Both sides list the nodes of a pair in the same order (ascending global key).  primary[n_corners] = 1
 * where this rank counts the corner in the all-reduced inner products of the GMRES.                                          */
int amps_gpu_field_halo_set(amps_gpu_ctx *ctx, int peer, const int32_t *corner_send, int64_t n_corner_send,
                            const int32_t *corner_recv, int64_t n_corner_recv, const int32_t *center_send,
                            int64_t n_center_send, const int32_t *center_recv, int64_t n_center_recv);
int amps_gpu_field_primary_set(amps_gpu_ctx *ctx, const uint8_t *primary);
/* E^n on the unique corners [n_corners][3] (CurrentEOffset); B^n is B_cur of amps_gpu_fields_upload */
int amps_gpu_E_upload(amps_gpu_ctx *ctx, const double *E_cur);
/* one field step with J, M of the last deposit: GMRES(restart; <= 0 = 30) from x0 = 0 until |r| <= tol |r0| or max_iter
 * products; afterwards E_half = E^{n+theta}, E = E^{n+1}, B_prev = B^n, B_cur = B^{n+1} on the device and in the tiles the
 * movers / the deposit read, i.e. amps_gpu_step may follow directly (the order of PIC::TimeStep).  warm_start != 0 starts
 * from the increment of the previous step instead of the reference's zero guess (SetInitialGuess, :6566); the stopping test
 * is |r| <= tol |rhs| in both cases.                                                                                       */
int amps_gpu_field_step(amps_gpu_ctx *ctx, double theta, double tol, int max_iter, int restart, int warm_start,
                        int *iterations, double *rel_residual);
/* any pointer may be NULL; E_cur, E_half [n_corners][3], B_cur [n_centers][3] */
int amps_gpu_fields_download(amps_gpu_ctx *ctx, double *E_cur, double *E_half, double *B_cur);

/* background E, B of the coupler on the unique centre nodes, [n_centers][3] each (DATAFILE::Offset::ElectricField /
 * MagneticField of cDataCenterNode, pic.h:8338-8425); either may be NULL = keep / zero            */
int amps_gpu_background_upload(amps_gpu_ctx *ctx, const double *E_center, const double *B_center);
/* the 15 tabulated drift variables of the relativistic guiding-centre mover on the unique centre nodes, [n_centers][15]:
 * b.grad(b), vE.grad(b), b.grad(vE), vE.grad(vE), grad(kappa*B), 3 components each in that order
 * (PIC::CPLR::GetVarForRelativisticGCA, pic.h:8643-8680; produced by DATAFILE, pic_datafile.cpp:1164-1340)   */
int amps_gpu_background_upload_gca(amps_gpu_ctx *ctx, const double *var15_center);
/* grad B of the coupler on the unique centre nodes, [n_centers][9] = {d/dx,d/dy,d/dz} of Bx, then By, then Bz
 * (PIC::CPLR::GetBackgroundMagneticFieldGradient, pic.h:8434-8470); read by the guiding-centre movers          */
int amps_gpu_background_upload_gradB(amps_gpu_ctx *ctx, const double *gradB_center);
/* InitiateMagneticMoment for every resident particle, as InitiateParticle does (pic_pbuffer.cpp:986-997):
 * mover_id = AMPS_MOVER_RELATIVISTIC_GCA -> Relativistic::GuidingCenter (pic_mover_relativistic_guiding_center.cpp:19-93);
 * AMPS_MOVER_GC_FIRST_ORDER/_SECOND_ORDER -> GuidingCenter (pic_mover_guiding_center.cpp:85-144; this one also aligns
 * v with B).  Needs carry_magnetic_moment and the background table.  GuidingCenter::Mover_FirstOrder does the
 * same by itself for particles whose InitFlag (bit 6 of the species byte) is clear.                          */
int amps_gpu_magnetic_moment_init(amps_gpu_ctx *ctx, int mover_id);
/* mu_by_ptr[ptr] -> device (SetMagneticMoment on the records the particles came from); n = length of mu_by_ptr */
int amps_gpu_magnetic_moment_upload(amps_gpu_ctx *ctx, const double *mu_by_ptr, int64_t n);
/* current device order (pair with the ptrs of amps_gpu_particles_download_soa) */
int amps_gpu_magnetic_moment_download(amps_gpu_ctx *ctx, double *mu, int64_t n_max, int64_t *n);
/* v_parallel of the gyrokinetic reduced state (PB::SetVParallel / GetVParallel), same conventions as the magnetic moment */
int amps_gpu_v_parallel_upload(amps_gpu_ctx *ctx, const double *vpar_by_ptr, int64_t n);
/* PB::GetVNormal of the guiding-centre species (cfg.gc_species_mask), by ParticleBuffer slot: read by the deposit's energy / cfl
 * diagnostics (pic_field_solver_ecsim.cpp:2232-2235); the device never changes it.  Slots beyond n read 0. */
int amps_gpu_v_normal_upload(amps_gpu_ctx *ctx, const double *vnormal_by_ptr, int64_t n);
/* State of the reference's global StencilTable (PIC::InterpolationRoutines::CellCentered::StencilTable): GetTriliniarInterpolationStencil
 * normalises a caller's stencil unless the GLOBAL table holds an 8-cell stencil (pic_interpolation_routines.cpp:903), and only
 * ComputeNetCharge fills it (:4783).  full = 1 after the first ComputeNetCharge of a run with the div-E correction: the exact
 * Lapenta2017 kernel and the ECSIM-field guiding-centre movers then leave a full B stencil un-normalised like the reference (x', v'
 * change by <= 2 ulp; the contracted production mover and the deposit are not affected).  Default 0. */
int amps_gpu_global_stencil_set(amps_gpu_ctx *ctx, int32_t full);
int amps_gpu_v_parallel_download(amps_gpu_ctx *ctx, double *vpar, int64_t n_max, int64_t *n);
/* exit records accumulated by the movers since the last call (clears them) */
int amps_gpu_exit_records(amps_gpu_ctx *ctx, amps_gpu_exit_record *buf, int64_t max_records, int64_t *n);

/* ---- particle store: PIC::ParticleBuffer as a device SoA sorted by (block,cell) ---- */
/* AoS records + the cell each one is attached to (global cell = leaf*Nx*Ny*Nz +
 * i+Nx*(j+Ny*k), i.e. the FirstCellParticleTable slot it hangs on).              */
int amps_gpu_particles_upload_aos(amps_gpu_ctx *ctx, const void *records, const int64_t *ptrs,
                                  const int32_t *cells, int64_t n, const amps_gpu_aos_layout *lay);
/* SoA upload: x[3][n], v[3][n] component-major; w may be NULL (w=1); ptrs = ParticleBuffer
 * slot of each particle (NULL = 0..n-1), carried through sorts so a download can be
 * matched to the caller's records.                                                 */
int amps_gpu_particles_upload_soa(amps_gpu_ctx *ctx, const double *x, const double *v,
                                  const double *w, const uint8_t *species, const int32_t *cells,
                                  const int32_t *ptrs, int64_t n);
/* the same, appended behind the resident particles (injection between epochs, PIC::ParticleBuffer::GetNewParticle +
 * InitiateParticle pic_pbuffer.cpp:371-437, :939-1027; large populations handed over in pieces). ptrs NULL = count.. */
int amps_gpu_particles_append_soa(amps_gpu_ctx *ctx, const double *x, const double *v,
                                  const double *w, const uint8_t *species, const int32_t *cells,
                                  const int32_t *ptrs, int64_t n);
int amps_gpu_particle_count(amps_gpu_ctx *ctx, int64_t *n);
/* current device order; any output pointer may be NULL */
int amps_gpu_particles_download_soa(amps_gpu_ctx *ctx, double *x, double *v, double *w,
                                    uint8_t *species, int32_t *cells, int32_t *ptrs, int64_t n_max,
                                    int64_t *n);
/* writes records back (ptr i -> slot i of `records`) and threads next/prev lists +
 * first_cell_particle[n_cells] like FirstCellParticleTable                        */
int amps_gpu_particles_download_aos(amps_gpu_ctx *ctx, void *records, int64_t *first_cell_particle,
                                    int64_t n_max, const amps_gpu_aos_layout *lay, int64_t *n);
/* Slot bookkeeping of the caller's ParticleBuffer before a download_aos when particles were deleted (boundary,
 * movers) or exchanged between ranks.  n_new = resident records without a slot of the caller's buffer (arrivals of
 * amps_gpu_migrate): the caller takes that many records with GetNewParticle (pic_pbuffer.cpp:371-437) and hands them
 * over with amps_gpu_particles_assign_slots.  released[] = slots that were uploaded but whose particle is gone (deleted,
 * migrated away): the caller returns each with DeleteParticle (pic_pbuffer.cpp:594-666).  When max_released is too small
 * only n_released is reported and nothing changes.                                   */
int amps_gpu_particles_slot_delta(amps_gpu_ctx *ctx, int64_t *n_new, int64_t *released, int64_t max_released,
                                  int64_t *n_released);
int amps_gpu_particles_assign_slots(amps_gpu_ctx *ctx, const int64_t *slots, int64_t n_slots);
/* SURVEY 8f row f3, restart half: PIC::Restart::SaveParticleData / ReadParticleData (pic_restart.cpp:248-420, :430-600) from /
 * into the sorted device store in the reference's own file format (AMPS reads what this writes and vice versa): after the caller's
 * header (user data, end marker, ParticleDataLength, GlobalParticleWeight[]) one section per block with particles:
 * cAMRnodeID | int nTotal | int ParticleNumberTable[Nx][Ny][Nz] | nTotal records of lay->stride bytes, cells in the loop order
 * i, j, k.  leaf_node_ids[n_leaves][id_bytes] = node->AMRnodeID of every leaf (opaque bytes).  Record fields the device does not
 * hold are written as zero.  restart_read replaces the resident particles by those of this rank's blocks.                       */
int amps_gpu_restart_save(amps_gpu_ctx *ctx, const char *fname, const void *header, int64_t header_bytes,
                          const void *leaf_node_ids, int32_t id_bytes, const amps_gpu_aos_layout *lay, int64_t *n_saved);
int amps_gpu_restart_read(amps_gpu_ctx *ctx, const char *fname, int64_t header_bytes, const void *leaf_node_ids,
                          int32_t id_bytes, const amps_gpu_aos_layout *lay, int64_t *n_loaded);
/* per-cell particle ranges after a sort: cell_start[n_cells+1] (CreateParticleTable,
 * pic_pbuffer.cpp:1160-1310)                                                      */
int amps_gpu_cell_table_download(amps_gpu_ctx *ctx, int64_t *cell_start, int64_t n_cells_plus_1);

/* counting sort by (block,cell); drops deleted particles. Replaces the temp->first list
 * swap of MoveParticles (pic_mover.cpp:1056-1088) and CreateParticleTable.       */
int amps_gpu_sort(amps_gpu_ctx *ctx);

/* ---- push: <- PIC::Mover::MoveParticles() (pic_mover.cpp:580-1088) through
 *      PIC::Mover::UserDefinedMoverManager.  In-place: particle slot i keeps slot i
 *      until amps_gpu_sort().  Periodic ghost->real wrap (pic_bc_periodic.cpp:86-178)
 *      is folded into the mover epilogue.  stats may be NULL (no host sync then). ---- */
int amps_gpu_move(amps_gpu_ctx *ctx, int mover_id, amps_gpu_move_stats *stats);

/* ---- deposit: <- ECSIM::UpdateJMassMatrix() (pic_field_solver_ecsim.cpp:3244-3995).
 *      Zeroes J[3],M[243] per unique corner, accumulates ProcessCell over all cells of
 *      real (non periodic-ghost) blocks, which with the unique-corner table also is the
 *      periodic/ghost corner reduction (ProcessJMassMatrix :1383).  energy (1) and
 *      cfl (n_species) may be NULL.                                               ---- */
int amps_gpu_deposit_JM(amps_gpu_ctx *ctx, double *particle_energy, double *cfl);
/* J[n_corners][3], M[n_corners][243] (neighbour-major, 9 per neighbour as in
 * IndexMatrix, pic_field_solver_ecsim.cpp:1377-1380); either may be NULL          */
/* amps_gpu_step followed by amps_gpu_JM_download, with the download pipelined: the deposit runs in ranges of blocks and the
 * corners whose last contributing block lies in a finished range travel to the host (pinned memory for real overlap) while the
 * next range is deposited.  Same results as the two calls.                                         */
int amps_gpu_step_JM(amps_gpu_ctx *ctx, int mover_id, double *J_host, double *M_host);
/* ECSIM::ComputeNetCharge() (pic_field_solver_ecsim.cpp:4690-4828, the particle pass of divECorrection): rho_new on the unique
 * centre nodes, [n_centers]; charge_conv multiplies cfg.charge[] (ECSIM::charge_conv).  rho_center may be NULL (result stays
 * on the device).  Single-rank: the sum over ranks of shared centres is the caller's (ProcessNetCharge).            */
int amps_gpu_net_charge(amps_gpu_ctx *ctx, double charge_conv, double *rho_center);

/* J and M in packed rows: JM_packed[n_corners][129] = J[3], then the 9 doubles of the 14 neighbour slots listed by
 * amps_gpu_JM_packed_slots() (self, and of every pair of opposite neighbours the one whose highest non-zero dimension is +1).
 * The mass matrix is symmetric by construction - ProcessCell adds the same 3x3 block to corner c under neighbour c' and to corner
 * c' under neighbour c (pic_field_solver_ecsim.cpp:2411-2420), i.e. M[c][slot(d)] == M[c+d][slot(-d)] - so the host rebuilds the
 * other 13 slots while it scatters the rows into the corner buffers, and 270 MB instead of 516 MB cross PCIe per step at 64^3
 * cells.  amps_gpu_step_JM_packed is amps_gpu_step_JM with this row format (same pipelining behind the deposit).               */
const int32_t *amps_gpu_JM_packed_slots(void);
int amps_gpu_JM_download_packed(amps_gpu_ctx *ctx, double *JM_packed_host);
int amps_gpu_step_JM_packed(amps_gpu_ctx *ctx, int mover_id, double *JM_packed_host);

/* PIC::Sampling::SamplingManager() + ProcessCell (pic.cpp:1045-1082, :705-990) on the device store: one more sample is ADDED to the
 * collecting buffer sample[n_leaves*cells][n_species][13] = {ParticleWeight, ParticleNumber, NumberDensity (w / cell volume),
 * ParticleVelocity[3] (w v), ParticleVelocity2[3] (w v_i^2), ParticleSpeed (w |v|), ParticleVelocity2Tensor[3] (w v_i v_(i+1)%3)} --
 * the sampled datums of pic.h:4117-4130 with the velocity tensor on; parallel/tangential temperature, internal degrees of freedom,
 * dust and user sampling are not sampled.  Every block is sampled, ghost blocks included, like the reference.  Needs the sorted
 * layout.  amps_gpu_sample_download copies the buffer (may be NULL) and the number of particles sampled per species since the
 * last clear (localSimulatedSpeciesParticleNumber); clear != 0 starts a new collecting period.                                  */
int amps_gpu_sample_cells(amps_gpu_ctx *ctx);
int amps_gpu_sample_download(amps_gpu_ctx *ctx, double *sample, int64_t *n_sampled, int clear);

/* The per-species corner moments UpdateJMassMatrix samples when _PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_ is on
 * (pic_field_solver_ecsim.cpp:2270-2300 per particle, :2384-2392 per cell, :3874-3879 flush; corner buffer from
 * SpeciesDataIndex[0] = 9+243 on, :527-531): spec_corner[n_corners][n_species][10] =
 * {Rho, RhoUx, RhoUy, RhoUz, RhoUxUx, RhoUyUy, RhoUzUz, RhoUxUy, RhoUyUz, RhoUxUz} (mass density moments / cell volume).
 * Needs the sorted layout.  The result stays on the device for amps_gpu_correct_particle_location; spec_corner may be NULL. */
int amps_gpu_species_moments(amps_gpu_ctx *ctx, double *spec_corner);

/* phi of the div-E correction on the unique centre nodes (centre buffer, phiIndex; written by the reference's Poisson solve,
 * divECorrection :4341-4352): phi_center[n_centers].                                                                   */
int amps_gpu_phi_upload(amps_gpu_ctx *ctx, const double *phi_center);

/* ECSIM::CorrectParticleLocation() (pic_field_solver_ecsim.cpp:4440-4688), the particle shift of divECorrection: species 0 moves
 * by -0.9 grad(phi) / (4 pi rho_e), at most 0.1 cell, rho_e = species-0 density on the closest corner times q/m (charge_conv,
 * mass_conv as in :4449-4452); every particle is re-filed (call amps_gpu_sort afterwards).  Particles that leave the domain are
 * deleted (n_deleted).  Single-level meshes; the reference runs it with periodic boundaries only (:4361-4363).             */
int amps_gpu_correct_particle_location(amps_gpu_ctx *ctx, double charge_conv, double mass_conv, int64_t *n_displaced, int64_t *n_deleted);
/* particle energy and per-species cfl of the last deposit (amps_gpu_deposit_JM or amps_gpu_step; after
 * amps_gpu_exchange_JM they are the all-reduced values)                                            */
int amps_gpu_diagnostics(amps_gpu_ctx *ctx, double *particle_energy, double *cfl);
int amps_gpu_JM_download(amps_gpu_ctx *ctx, double *J, double *M);
/* device pointers for an on-device consumer (field solve) */
int amps_gpu_JM_device(amps_gpu_ctx *ctx, double **J_dev, double **M_dev);

/* ---- multi-GPU: one rank per GPU, NCCL over NVLink (libnccl is resolved with dlopen at the first call) ----
 * <- MPI_GLOBAL_COMMUNICATOR.  Rank 0 creates the id, the host broadcasts its 128 bytes (MPI_Bcast /
 *    torch.distributed), every rank joins.                                                     */
int amps_gpu_comm_unique_id(void *id128);
int amps_gpu_comm_init(amps_gpu_ctx *ctx, const void *id128, int rank, int n_ranks);
/* 1 when the ranks of the communicator could map each other's receive buffers (CUDA IPC over NVLink / NVSwitch): the particle
 * migration then writes the leavers straight into the owner's buffer and never waits for counts on the host; 0 = NCCL
 * send / recv with a count round trip (other boxes, or AMPS_GPU_PEER_MIGRATE=0)                                       */
int amps_gpu_comm_uses_peer_memory(amps_gpu_ctx *ctx);
/* corners that this rank and `peer` both deposit into: local unique-corner ids, both sides ordered by the
 * same global corner key (amps_b200/mesh.py: shared_corner_lists)                               */
int amps_gpu_set_shared_corners(amps_gpu_ctx *ctx, int peer, const int32_t *uids, int64_t n);
/* <- PIC::Parallel::ExchangeParticleData (pic_parallel.cpp:50-488): call between move and sort.
 *    n_sent / n_received may be NULL.                                                          */
int amps_gpu_migrate(amps_gpu_ctx *ctx, int64_t *n_sent, int64_t *n_received);
/* <- SyncMassMatrix / ProcessCornerBlockBoundaryNodes (ecsim/halo_sync.cpp:79-124) + the MPI_Reduce of the
 *    particle energy (SUM) and cfl (MAX) (pic_field_solver_ecsim.cpp:3972-3977): call after deposit.  */
int amps_gpu_exchange_JM(amps_gpu_ctx *ctx);

/* one fused ECSIM particle phase: move (+ migrate) + sort + deposit (+ corner exchange), no host sync inside
 * on a single rank; the migration needs one small device->host read of the counts          */
int amps_gpu_step(amps_gpu_ctx *ctx, int mover_id);

/* Phase timing with CUDA events recorded on the context's stream (what the reference prints from
 * ECSIM::CumulativeTiming::{ParticleMoverTime,UpdateJMassMatrixTime}, pic_field_solver_ecsim.cpp:154-181).
 * Synchronises, returns the ms and call counts accumulated per phase since the last call, clears
 * them, and enables/disables recording for the following calls.                          */
enum { AMPS_GPU_PHASE_MOVE = 0, AMPS_GPU_PHASE_SORT = 1, AMPS_GPU_PHASE_DEPOSIT = 2, AMPS_GPU_PHASE_EXCHANGE = 3, AMPS_GPU_N_PHASES = 4 };
int amps_gpu_profile(amps_gpu_ctx *ctx, int enable, double *phase_ms, int64_t *phase_count);

/* Device self test of the mover's shared-reciprocal quotients: counts (a[i],b[j]) pairs whose result
 * differs from the IEEE fp64 division (must be 0; the cell assignment parity depends on it).   */
int amps_gpu_selftest_division(amps_gpu_ctx *ctx, const double *a, const double *b, int64_t n, int64_t *n_mismatch);

/* block until all queued work of the context is complete */
int amps_gpu_synchronize(amps_gpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* AMPS_GPU_H */
