/*
 * amps_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A literal CPU restatement of the AMPS reference algorithms on the hot path
 * (movers + ECSIM J/mass-matrix deposition) using the reference's own data
 * structures: AoS byte particle records with next/prev links, per-cell lists
 * (FirstCellParticleTable / tempParticleMovingListTable), a pointer tree of
 * cTreeNodeAMR-like nodes, corner/centre node objects with associated-data
 * buffers.  Each function cites the reference file:line it follows.
 *
 * PARITY PINNING: the reference's golden outputs for this path live in an
 * external, un-vendored data repository, and the reference
 * cannot be compiled here (needs MPI, the Perl-generated build/ tree and
 * un-vendored SWMF share/ headers).  For the ECSIM push/deposit the oracle is
 * therefore pinned only by the structural invariants the reference states
 * (sum W = 1, symmetric mass matrix scatter, CPU-variant agreement) =>
 * "parity unpinned" for those; the relativistic Boris momentum rotation is pinned
 * against srcEarth/gridless, and the tree search / cell index / neighbour probes /
 * neighbour level limits against the reference's own mesh class, both compiled
 * from the reference sources into oracle/_ref (tests/test_reference_mesh.py,
 * tests/test_relativistic_boris.py).  See DESIGN.md.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef AMPS_ORACLE_H
#define AMPS_ORACLE_H

#include <stdint.h>
#include "../include/amps_gpu.h" /* POD descriptions only (config, mesh, stats) */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_ctx oracle_ctx;

/* PIC::ParticleBuffer::Init + mesh construction from the flattened description */
oracle_ctx *oracle_create(const amps_gpu_config *cfg, const amps_gpu_mesh *mesh);
void oracle_destroy(oracle_ctx *);
const char *oracle_last_error(const oracle_ctx *);

/* byte layout of one particle record (packed, picParticleDataMacro.h:55-81) */
int64_t oracle_particle_data_length(const oracle_ctx *);

/* corner E_half[n_corners][3], centre B_prev/B_cur[n_centers][3] */
void oracle_set_fields(oracle_ctx *, const double *E_half, const double *B_prev, const double *B_cur);

/* coupler table for the test-particle movers: E, B on the unique centre nodes [n_centers][3] */
void oracle_set_background(oracle_ctx *, const double *E_center, const double *B_center);
/* the 15 tabulated variables of the relativistic GCA on the unique centre nodes [n_centers][15] */
void oracle_set_background_gca(oracle_ctx *, const double *var15);
/* grad B on the unique centre nodes [n_centers][9] = {d/dx,d/dy,d/dz} of Bx, By, Bz */
void oracle_set_background_gradB(oracle_ctx *, const double *gradB);
/* (Relativistic::)GuidingCenter::InitiateMagneticMoment for every listed particle (mover_id picks the variant);
 * mu_out[ptr] if not NULL */
int oracle_magnetic_moment_init(oracle_ctx *, int mover_id, double *mu_out, int64_t n);
/* by ptr: magnetic moment and InitFlag (bit 6 of the species byte) */
void oracle_get_magnetic_moment(const oracle_ctx *, double *mu, uint8_t *init_flag, int64_t n);
/* gyrokinetic reduced state by ptr: mu and v_parallel (either may be NULL) */
void oracle_set_reduced_state(oracle_ctx *, const double *mu, const double *vpar, int64_t n);
/* v_normal by ptr (PB::SetVNormal): read by ProcessCell for the guiding-centre species of cfg.gc_species_mask */
void oracle_set_v_normal(oracle_ctx *o, const double *vnormal, int64_t n);
/* the current E on the unique corners: read by the guiding-centre movers when cfg.gc_fields_ecsim */
void oracle_set_E_current(oracle_ctx *o, const double *E);
/* Length of the stencil in the reference's global StencilTable (8 once ComputeNetCharge has run on a periodic box): the movers and
 * ProcessCell skip Normalize() of their B stencil then (pic_interpolation_routines.cpp:903) */
void oracle_set_global_stencil_length(oracle_ctx *o, int length);
/* ECSIM::GetElectricField / GetMagneticField / GetMagneticFieldGradient at n points (x[n][3], each in its leaf) */
int oracle_ecsim_fields(const oracle_ctx *o, int64_t n, const double *x, const int32_t *leaf, double *E, double *B, double *gradB);
void oracle_get_v_parallel(const oracle_ctx *, double *vpar, int64_t n);
/* exit records (domain faces / internal sphere) accumulated since the last call; returns their number */
int64_t oracle_exit_records(oracle_ctx *, amps_gpu_exit_record *buf, int64_t max_records);

/* InitiateParticle(...,ADD2LIST) for n particles: particle i gets ptr i */
int oracle_add_particles(oracle_ctx *, const double *x, const double *v, const double *w,
                         const uint8_t *species, const int32_t *cells, int64_t n);
int64_t oracle_particle_count(const oracle_ctx *);

/* PIC::Mover::MoveParticles() with the Lapenta2017 (or other) mover, followed by
 * PIC::BC::ExternalBoundary::Periodic::ExchangeParticles() when periodic.
 * n_threads>1 = OpenMP split by blocks (pic_mover.cpp:716-766).
 * per-particle outputs (indexed by ptr, may be NULL): return code, final cell (-1 deleted) */
int oracle_move(oracle_ctx *, int mover_id, int n_threads, amps_gpu_move_stats *stats,
                int32_t *ret_code, int32_t *final_cell);

/* read back by ptr (slot i): x[3][n] v[3][n] component-major like the SoA upload */
void oracle_get_particles(const oracle_ctx *, double *x, double *v, double *w, uint8_t *species,
                          int32_t *cells, uint8_t *alive, int64_t n);

/* ECSIM::UpdateJMassMatrix(): J[n_corners][3], M[n_corners][243] */
int oracle_deposit_JM(oracle_ctx *, int n_threads, double *J, double *M, double *energy, double *cfl);

/* ECSIM::ComputeNetCharge(): rho_new on the unique centre nodes [n_centers] */
int oracle_net_charge(oracle_ctx *, double charge_conv, double *rho);
/* the _PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_ part of UpdateJMassMatrix: [n_corners][10*n_species] (spec may be NULL) */
int oracle_species_moments(oracle_ctx *, double *spec);
/* probes of the restated src/general/specfunc.h helpers */
double oracle_probe_gyro_frequency(const double *v, double m, double q, const double *B, double c);
void oracle_probe_normalize(double *x);
/* PIC::Sampling::SamplingManager(): adds one sample to the collecting buffer [n_leaves*cells][n_species][13] and returns it */
int oracle_sample_cells(oracle_ctx *, double *sample, int64_t *n_sampled);
/* phi of the div-E correction on the unique centre nodes [n_centers] */
void oracle_set_phi(oracle_ctx *, const double *phi_center);
/* ECSIM::CorrectParticleLocation(): needs oracle_species_moments + oracle_set_phi; final_cell[MaxNPart] = block*cells+cell or -1 */
int oracle_correct_particle_location(oracle_ctx *, double charge_conv, double mass_conv, int32_t *final_cell, int64_t *n_displaced,
                                     int64_t *n_deleted);

/* stand-alone pieces for unit parity tests */
int oracle_find_tree_node(const oracle_ctx *, const double *x, int start_leaf); /* leaf id or -1 */
int oracle_find_cell_index(const oracle_ctx *, const double *x, int leaf, int *ijk);
/* CornerBased::InitStencil: W[8] in cell-corner order, local ids[8], normalised weights[8]; returns Length */
int oracle_corner_stencil(const oracle_ctx *, double *x_inout, int leaf, double *W, int *ids, double *wnorm);
/* CellCentered::Linear::InitStencil: returns Length */
int oracle_center_stencil(const oracle_ctx *, const double *x, int leaf, int *ids, double *w);

/* PIC::CPLR::InitInterpolationStencil (AMR capable): unique centre ids + weights (<=64); returns Length, -1 = reference exit() */
int oracle_coupler_stencil(const oracle_ctx *, const double *x, int leaf, int *uids, double *w);
void oracle_neib_levels(const oracle_ctx *, int leaf, int *minmax);
int oracle_neib(const oracle_ctx *, int leaf, int kind, int idx, double *lo, double *hi, int *level);

/* ParticleBuffer list checks (CheckParticleList, pic_pbuffer.cpp:807) : 0 ok */
int oracle_check_particle_lists(const oracle_ctx *);

#ifdef __cplusplus
}
#endif
#endif
