// ref_pic_shim.cpp -- OURS (test infrastructure).  Runs the reference's OWN PIC code as a one-rank checker of the oracle:
// the translation units of src/pic, src/general, src/meshAMR, ... are compiled from the tree that the reference's
// ampsConfig.pl generates for input/test/fast-wave.input (ECSIM, Lapenta2017, periodic box, 16x8x4-cell blocks), with our
// stand-ins for mpi.h and the three un-vendored SWMF share headers (oracle/ref_pic/*.h, oracle/ref_mesh/mpi.h).  This file
// includes the reference's fast-wave test program for its initial conditions (its main() renamed, never called) and exports a C
// interface that drives  PIC::Mover::MoveParticles (-> Lapenta2017, CornerBased / CellCentered InitStencil, findTreeNode,
// periodic wrap)  and  ECSIM::UpdateJMassMatrix (-> ProcessCell, ProcessJMassMatrix)  and reads the state back.
#define main ref_fastwave_main_never_called
#include "main.cpp"  // test/srcFastWave/main.cpp of the reference (found through -I, not copied)
#undef main

// ---- symbols of translation units that are not part of this build (the SWMF fluid coupler, Fortran readers, user models) ----
// (linear_solver_matvec_c and the GMRES behind linear_solver_wrapper: gmres_single.cpp)
long int PIC::CPLR::FLUID::iCycle = 0;
bool PIC::CPLR::FLUID::IsRestart = false;
double PIC::CPLR::FLUID::EFieldTol = 1.0e-6;
double PIC::CPLR::FLUID::EFieldIter = 200;
void PIC::CPLR::FLUID::write_output(double, bool) {}
void PIC::CPLR::FLUID::check_max_mem_usage(std::string) {}
void PIC::CPLR::InitInterpolationStencil(double *, cTreeNodeAMR<PIC::Mesh::cDataBlockAMR> *) { abort(); }
double Exosphere::GetSurfaceTemperature(double, double *) { abort(); }
double Exosphere::SurfaceInteraction::StickingProbability(int, double &, double) { abort(); }
int Exosphere::ColumnIntegral::GetVariableList(char *) { return 0; }
void Exosphere::ColumnIntegral::ProcessColumnIntegrationVector(double *, int) {}
void Exosphere::ColumnIntegral::CoulumnDensityIntegrant(double *, int, double *, cTreeNodeAMR<PIC::Mesh::cDataBlockAMR> *) {}
double Exosphere::OrbitalMotion::GetTAA(SpiceDouble) { return 0.0; }
extern "C" {
void batsrus2amps_set_mpi_parameters_(...) { abort(); }
void batsrus2amps_read_file_header_(...) { abort(); }
void batsrus2amps_openfile_(...) { abort(); }
void batsrus2amps_get_nvar_(...) { abort(); }
void batsrus2amps_get_namevardata_(...) { abort(); }
void batsrus2amps_get_nameunitdata_(...) { abort(); }
void batsrus2amps_get_data_point_(...) { abort(); }
void batsrus2amps_domain_limits_(...) { abort(); }
void batsrus2amps_closefile_(...) { abort(); }
}

extern double TotalParticleEnergy;  // file-scope global of pic_field_solver_ecsim.cpp (:228)
namespace {
typedef cTreeNodeAMR<PIC::Mesh::cDataBlockAMR> Node;
std::vector<Node *> g_blocks;  // every allocated block in BranchBottomNodeList order
bool is_ghost(Node *node) {    // the reference's own test (PrepopulateDomain, UpdateJMassMatrix): a block at the domain boundary
  for (int iface = 0; iface < 6; iface++)
    if (node->GetNeibFace(iface, 0, 0, PIC::Mesh::mesh) == NULL) return true;
  return false;
}
const int NX = _BLOCK_CELLS_X_, NY = _BLOCK_CELLS_Y_, NZ = _BLOCK_CELLS_Z_, GX = _GHOST_CELLS_X_, GY = _GHOST_CELLS_Y_, GZ = _GHOST_CELLS_Z_;
}  // namespace

extern "C" {
// the set-up of the reference's main(), up to its time loop (file outputs left out)
int ref_pic_init(void) {
  PIC::InitMPI();
  PIC::Init_BeforeParser();
  rnd_seed(100);
  PIC::Mesh::mesh->AllowBlockAllocation = false;
  PIC::BC::ExternalBoundary::Periodic::Init(xmin, xmax, BulletLocalResolution);
  PIC::Mesh::mesh->buildMesh();
  PIC::Mesh::initCellSamplingDataBuffer();
  PIC::Mesh::mesh->CreateNewParallelDistributionLists();
  PIC::Mesh::mesh->AllowBlockAllocation = true;
  PIC::Mesh::mesh->AllocateTreeBlocks();
  PIC::Mesh::mesh->InitCellMeasure();
  PIC::Init_AfterParser();
  PIC::Mover::Init();
  PIC::ParticleWeightTimeStep::LocalTimeStep = localTimeStep;
  PIC::ParticleWeightTimeStep::initTimeStep();
  PIC::BC::ExternalBoundary::Periodic::InitBlockPairTable();
  PIC::ParticleWeightTimeStep::SetGlobalParticleWeight(0, 1e-2 * 0.0795774715459477);
  PIC::ParticleWeightTimeStep::SetGlobalParticleWeight(1, 1e-2 * 0.0795774715459477);
  PIC::DomainBlockDecomposition::UpdateBlockTable();
  PIC::BC::ExternalBoundary::UpdateData();
  PIC::FieldSolver::Electromagnetic::ECSIM::SetIC = ::SetIC;
  PIC::FieldSolver::Electromagnetic::ECSIM::Init_IC();
  PIC::BC::ExternalBoundary::UpdateData();
  CleanParticles();
  PrepopulateDomain();
  PIC::BC::ExternalBoundary::UpdateData();
  g_blocks.clear();
  for (Node *node = PIC::Mesh::mesh->BranchBottomNodeList; node != NULL; node = node->nextBranchBottomNode)
    if (node->block != NULL) g_blocks.push_back(node);
  return (int)g_blocks.size();
}

// out[0..5] block cells and ghost cells, [6] blocks, [7] species, [8] particles, [9] ParticleDataLength
void ref_pic_dims(long *out) {
  out[0] = NX, out[1] = NY, out[2] = NZ, out[3] = GX, out[4] = GY, out[5] = GZ;
  out[6] = (long)g_blocks.size(), out[7] = PIC::nTotalSpecies, out[8] = PIC::ParticleBuffer::GetAllPartNum(), out[9] = PIC::ParticleBuffer::ParticleDataLength;
}
// species tables and the constants ProcessCell / Lapenta2017 use: out = charge[nS] mass[nS] weight[nS] dt, LightSpeed, and the unit factors
void ref_pic_constants(double *out) {
  using namespace PIC::FieldSolver::Electromagnetic::ECSIM;
  int n = 0;
  for (int s = 0; s < PIC::nTotalSpecies; s++) out[n++] = PIC::MolecularData::GetElectricCharge(s);
  for (int s = 0; s < PIC::nTotalSpecies; s++) out[n++] = PIC::MolecularData::GetMass(s);
  for (int s = 0; s < PIC::nTotalSpecies; s++) out[n++] = PIC::ParticleWeightTimeStep::GlobalParticleWeight[s];
  out[n++] = PIC::ParticleWeightTimeStep::GlobalTimeStep[0];
  out[n++] = LightSpeed;
  out[n++] = B_conv, out[n++] = E_conv, out[n++] = length_conv, out[n++] = charge_conv, out[n++] = mass_conv, out[n++] = cDt, out[n++] = theta;
  out[n++] = _AMU_, out[n++] = ElectronCharge;
}
// charge and mass of the species as Lapenta2017 (pic_mover_boris.cpp:1181-1183) and ProcessCell (:2239-2241) form them
void ref_pic_species(double *q_no, double *m_no) {
  for (int s = 0; s < PIC::nTotalSpecies; s++) {
    q_no[s] = picunits::si2no_q(PIC::MolecularData::GetElectricCharge(s), PIC::Units::Factors);
    m_no[s] = picunits::si2no_m(PIC::MolecularData::GetMass(s), PIC::Units::Factors);
  }
}
void ref_pic_blocks(double *bxmin, double *bxmax, int *ghost) {
  for (size_t b = 0; b < g_blocks.size(); b++) {
    for (int d = 0; d < 3; d++) bxmin[3 * b + d] = g_blocks[b]->xmin[d], bxmax[3 * b + d] = g_blocks[b]->xmax[d];
    ghost[b] = is_ghost(g_blocks[b]) ? 1 : 0;
  }
}
// corner data of every block, all nodes incl. the ghost layers, [block][k][j][i][len]; what: 0 E (current), 1 E at the half step,
// 2 J (3 values), 3 the mass matrix (243 values).  Missing nodes read as NaN.
static int corner_slice(int what, int *off) {
  using namespace PIC::FieldSolver::Electromagnetic::ECSIM;
  switch (what) {
    case 0: *off = CurrentEOffset / (int)sizeof(double); return 3;
    case 1: *off = OffsetE_HalfTimeStep / (int)sizeof(double); return 3;
    case 2: *off = JxOffsetIndex; return 3;
    case 3: *off = MassMatrixOffsetIndex; return 243;
  }
  return 0;
}
long ref_pic_corner_rw(int what, double *buf, int write) {
  int off = 0;
  const int len = corner_slice(what, &off);
  long n = 0;
  for (Node *node : g_blocks)
    for (int k = -GZ; k <= NZ + GZ; k++)
      for (int j = -GY; j <= NY + GY; j++)
        for (int i = -GX; i <= NX + GX; i++) {
          PIC::Mesh::cDataCornerNode *c = node->block->GetCornerNode(_getCornerNodeLocalNumber(i, j, k));
          double *p = c ? (double *)(c->GetAssociatedDataBufferPointer() + PIC::CPLR::DATAFILE::Offset::ElectricField.RelativeOffset) + off : NULL;
          for (int q = 0; q < len; q++, n++) {
            if (write) { if (p && buf[n] == buf[n]) p[q] = buf[n]; }
            else buf[n] = p ? p[q] : NAN;
          }
        }
  return n;
}
// centre data [block][k][j][i][3] incl. ghost cells; what: 0 B (current), 1 B (previous)
long ref_pic_center_rw(int what, double *buf, int write) {
  using namespace PIC::FieldSolver::Electromagnetic::ECSIM;
  const int off = (what == 0 ? CurrentBOffset : PrevBOffset) / (int)sizeof(double);
  long n = 0;
  for (Node *node : g_blocks)
    for (int k = -GZ; k < NZ + GZ; k++)
      for (int j = -GY; j < NY + GY; j++)
        for (int i = -GX; i < NX + GX; i++) {
          PIC::Mesh::cDataCenterNode *c = node->block->GetCenterNode(_getCenterNodeLocalNumber(i, j, k));
          double *p = c ? (double *)(c->GetAssociatedDataBufferPointer() + PIC::CPLR::DATAFILE::Offset::MagneticField.RelativeOffset) + off : NULL;
          for (int q = 0; q < 3; q++, n++) {
            if (write) { if (p && buf[n] == buf[n]) p[q] = buf[n]; }
            else buf[n] = p ? p[q] : NAN;
          }
        }
  return n;
}
// one double of the ECSIM centre data of every centre node incl. ghost cells, [block][k][j][i]; index in doubles from
// MagneticField.RelativeOffset: netChargeOldIndex 6, netChargeNewIndex 7, divEIndex 8, phiIndex 9 (pic_field_solver_ecsim.cpp:499-502)
long ref_pic_center_scalar_rw(int index, double *buf, int write) {
  long n = 0;
  for (Node *node : g_blocks)
    for (int k = -GZ; k < NZ + GZ; k++)
      for (int j = -GY; j < NY + GY; j++)
        for (int i = -GX; i < NX + GX; i++, n++) {
          PIC::Mesh::cDataCenterNode *c = node->block->GetCenterNode(_getCenterNodeLocalNumber(i, j, k));
          double *p = c ? (double *)(c->GetAssociatedDataBufferPointer() + PIC::CPLR::DATAFILE::Offset::MagneticField.RelativeOffset) + index : NULL;
          if (write) { if (p && buf[n] == buf[n]) *p = buf[n]; }
          else buf[n] = p ? *p : NAN;
        }
  return n;
}
// ECSIM::ComputeNetCharge (pic_field_solver_ecsim.cpp:4690): the charge density on the centre nodes (read it with index 7 above)
void ref_pic_compute_net_charge(int update_old) { PIC::FieldSolver::Electromagnetic::ECSIM::ComputeNetCharge(update_old != 0); }
// 1 when ProcessCell samples the species moments on the corners (_PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_, on in the gk variant)
int ref_pic_samples_species_on_corners(void) { return _PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_ == _PIC_MODE_ON_ ? 1 : 0; }
#if _PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_ == _PIC_MODE_ON_
// the 10 moments per species UpdateJMassMatrix leaves on the corners (SpeciesDataIndex, :531), [block][k][j][i][10 nSpecies]
long ref_pic_species_moments(double *buf) {
  using namespace PIC::FieldSolver::Electromagnetic::ECSIM;
  const int len = 10 * PIC::nTotalSpecies;
  long n = 0;
  for (Node *node : g_blocks)
    for (int k = -GZ; k <= NZ + GZ; k++)
      for (int j = -GY; j <= NY + GY; j++)
        for (int i = -GX; i <= NX + GX; i++) {
          PIC::Mesh::cDataCornerNode *c = node->block->GetCornerNode(_getCornerNodeLocalNumber(i, j, k));
          double *p = c ? (double *)(c->GetAssociatedDataBufferPointer() + PIC::CPLR::DATAFILE::Offset::ElectricField.RelativeOffset) + SpeciesDataIndex[0] : NULL;
          for (int q = 0; q < len; q++, n++) buf[n] = p ? p[q] : NAN;
        }
  return n;
}
// the particle half of ECSIM::divECorrection (:4354-4355): CorrectParticleLocation, then the rank exchange
void ref_pic_correct_particle_location(void) {
  PIC::FieldSolver::Electromagnetic::ECSIM::CorrectParticleLocation();
  PIC::Parallel::ExchangeParticleData();
}
#endif
// PIC::Sampling::SamplingManager (pic.cpp:1049 -> ProcessCell :705): one more sample of every cell into the collecting buffer.
// out[block][cell i + Nx (j + Ny k)][species][10] = weight, number, number density, velocity[3], velocity2[3], speed as the collecting buffer
// holds them afterwards (NAN for a datum this configuration does not sample); counts[species] = particles sampled by this call
void ref_pic_sample_cells(double *out, long *counts) {
  const int nS = PIC::nTotalSpecies;
  std::vector<int> tab(nS, 0);
  int *row = tab.data();
  PIC::Sampling::SamplingManager(&row);
  for (int s = 0; s < nS; s++) counts[s] = tab[s];
  PIC::Datum::cDatumSampled *datum[6] = {&PIC::Mesh::DatumParticleWeight, &PIC::Mesh::DatumParticleNumber, &PIC::Mesh::DatumNumberDensity,
                                         &PIC::Mesh::DatumParticleVelocity, &PIC::Mesh::DatumParticleVelocity2, &PIC::Mesh::DatumParticleSpeed};
  const int at[6] = {0, 1, 2, 3, 6, 9};
  long n = 0;
  for (Node *node : g_blocks)
    for (int k = 0; k < NZ; k++)
      for (int j = 0; j < NY; j++)
        for (int i = 0; i < NX; i++) {
          PIC::Mesh::cDataCenterNode *c = node->block->GetCenterNode(_getCenterNodeLocalNumber(i, j, k));
          for (int s = 0; s < nS; s++, n += 10) {
            for (int q = 0; q < 10; q++) out[n + q] = NAN;
            if (!c) continue;
            char *base = c->GetAssociatedDataBufferPointer() + PIC::Mesh::collectingCellSampleDataPointerOffset;
            for (int d = 0; d < 6; d++)
              if (datum[d]->offset >= 0)
                for (int q = 0; q < datum[d]->length; q++) out[n + at[d] + q] = *(q + datum[d]->length * s + (double *)(base + datum[d]->offset));
          }
        }
}
// every particle on a cell list: ParticleBuffer slot, x, v, individual weight correction, species, block (index of ref_pic_blocks)
// and cell i + Nx (j + Ny k), in the reference's own iteration order (block, k, j, i, list order)
long ref_pic_particles(long max_n, long *ptr, double *x, double *v, double *w, int *spec, int *block, int *cell) {
  long n = 0;
  for (size_t b = 0; b < g_blocks.size(); b++) {
    long int *first = g_blocks[b]->block->FirstCellParticleTable;
    if (!first) continue;
    for (int k = 0; k < NZ; k++)
      for (int j = 0; j < NY; j++)
        for (int i = 0; i < NX; i++)
          for (long int p = first[i + NX * (j + NY * k)]; p != -1; p = PIC::ParticleBuffer::GetNext(p)) {
            if (n < max_n) {
              ptr[n] = p;
              memcpy(x + 3 * n, PIC::ParticleBuffer::GetX(p), 24);
              memcpy(v + 3 * n, PIC::ParticleBuffer::GetV(p), 24);
              w[n] = PIC::ParticleBuffer::GetIndividualStatWeightCorrection(p);
              spec[n] = PIC::ParticleBuffer::GetI(p);
              block[n] = (int)b, cell[n] = i + NX * (j + NY * k);
            }
            n++;
          }
  }
  return n;
}
// keep every keep_every-th particle of the walk above, delete the others (PIC::ParticleBuffer::DeleteParticle): a population
// small enough to commit as a fixture together with everything the reference computes from it
long ref_pic_thin(int keep_every) {
  long n = 0, kept = 0;
  for (Node *node : g_blocks) {
    long int *first = node->block->FirstCellParticleTable;
    if (!first) continue;
    for (int c = 0; c < NX * NY * NZ; c++) {
      long int p = first[c];
      while (p != -1) {
        const long int next = PIC::ParticleBuffer::GetNext(p);
        if (n++ % keep_every) PIC::ParticleBuffer::DeleteParticle(p, first[c]);
        else kept++;
        p = next;
      }
    }
  }
  return kept;
}
// ---- PIC::Restart with the reference's own writer / reader (pic_restart.cpp:248, :560) ----
int ref_pic_node_id_bytes(void) { return (int)sizeof(cAMRnodeID); }
void ref_pic_node_ids(unsigned char *out) {  // AMRnodeID of every block of ref_pic_blocks
  for (size_t b = 0; b < g_blocks.size(); b++) memcpy(out + b * sizeof(cAMRnodeID), &g_blocks[b]->AMRnodeID, sizeof(cAMRnodeID));
}
void ref_pic_save_restart(const char *fname) { PIC::Restart::SaveParticleData(fname); }
// delete every particle, then load the file
long ref_pic_read_restart(const char *fname) {
  CleanParticles();
  PIC::Restart::ReadParticleData(fname);
  return PIC::ParticleBuffer::GetAllPartNum();
}
// offsets of the particle record (picParticleDataMacro.h): stride, species, v, x, weight correction, next, prev
void ref_pic_record_layout(long *out) {
  out[0] = PIC::ParticleBuffer::ParticleDataLength, out[1] = _PIC_PARTICLE_DATA__SPECIES_ID_OFFSET_, out[2] = _PIC_PARTICLE_DATA__VELOCITY_OFFSET_;
  out[3] = _PIC_PARTICLE_DATA__POSITION_OFFSET_, out[4] = _PIC_PARTICLE_DATA__WEIGHT_CORRECTION_OFFSET_;
  out[5] = _PIC_PARTICLE_DATA__NEXT_OFFSET_, out[6] = _PIC_PARTICLE_DATA__PREV_OFFSET_;
}
// ---- the gyrokinetic variant of the library (REF_PIC_VARIANT=gk: _PIC_GYROKINETIC_MODEL_MODE_ and _USE_MAGNETIC_MOMENT_ on) ----
// 1 when this build carries the reduced state in the particle record and ProcessCell / MoveParticles take their guiding-centre branches
int ref_pic_gyrokinetic(void) { return (_PIC_GYROKINETIC_MODEL_MODE_ == _PIC_MODE_ON_ && _USE_MAGNETIC_MOMENT_ == _PIC_MODE_ON_) ? 1 : 0; }
#if _PIC_GYROKINETIC_MODEL_MODE_ == _PIC_MODE_ON_ && _USE_MAGNETIC_MOMENT_ == _PIC_MODE_ON_
// offsets of the reduced state in the particle record: magnetic moment, v_parallel, v_normal
void ref_pic_reduced_layout(long *out) {
  out[0] = _PIC_PARTICLE_DATA__MAGNETIC_MOMENT_OFFSET_, out[1] = _PIC_PARTICLE_DATA__V_PARALLEL_OFFSET_, out[2] = _PIC_PARTICLE_DATA__V_NORMAL_OFFSET_;
}
void ref_pic_set_gc_species(int spec, int on) { PIC::GYROKINETIC::SetGuidingCenterSpecies(spec, on != 0); }
// what MoveParticles calls per particle in this variant: 0 = PIC::GYROKINETIC::Mover (pic_gyrokinetic.cpp:49: GuidingCenter::Mover_FirstOrder for
// the guiding-centre species, Lapenta2017 for the rest), 1 = the same routing with GuidingCenter::Mover_SecondOrder
static int g_mover_mode = 0;
void ref_pic_set_mover_mode(int mode) { g_mover_mode = mode; }
int ref_pic_user_mover(long int ptr, double dt, void *node) {
  Node *n = (Node *)node;
  if (g_mover_mode == 1) {
    const int spec = PIC::ParticleBuffer::GetI(ptr);
    if (PIC::GYROKINETIC::IsGuidingCenterSpecies(spec)) return PIC::Mover::GuidingCenter::Mover_SecondOrder(ptr, dt, n);
    return PIC::Mover::Lapenta2017(ptr, dt, n);
  }
  return PIC::GYROKINETIC::Mover(ptr, dt, n);
}
// PB::SetMagneticMoment / SetVNormal / the InitFlag of GuidingCenter::Mover_FirstOrder (:640-644), by ParticleBuffer slot; NULL = leave
void ref_pic_set_reduced(long n, const long *ptr, const double *mu, const double *vnormal, const int *init_flag) {
  for (long i = 0; i < n; i++) {
    if (mu) PIC::ParticleBuffer::SetMagneticMoment(mu[i], ptr[i]);
    if (vnormal) PIC::ParticleBuffer::SetVNormal(vnormal[i], ptr[i]);
    if (init_flag) PIC::ParticleBuffer::SetInitFlag(init_flag[i] != 0, PIC::ParticleBuffer::GetParticleDataPointer(ptr[i]));
  }
}
void ref_pic_get_reduced(long n, const long *ptr, double *mu, double *vnormal, int *init_flag) {
  for (long i = 0; i < n; i++) {
    if (mu) mu[i] = PIC::ParticleBuffer::GetMagneticMoment(ptr[i]);
    if (vnormal) vnormal[i] = PIC::ParticleBuffer::GetVNormal(PIC::ParticleBuffer::GetParticleDataPointer(ptr[i]));
    if (init_flag) init_flag[i] = PIC::ParticleBuffer::TestInitFlag(PIC::ParticleBuffer::GetParticleDataPointer(ptr[i])) ? 1 : 0;
  }
}
#endif
void ref_pic_set_weight_correction(long n, const long *ptr, const double *w) {
  for (long i = 0; i < n; i++) PIC::ParticleBuffer::SetIndividualStatWeightCorrection(w[i], ptr[i]);
}
// PIC::TimeStep's particle phase: MoveParticles (Lapenta2017 per particle), the rank exchange, the periodic ghost -> real hand-off
void ref_pic_move(void) {
  PIC::Mover::MoveParticles();
  PIC::Parallel::ExchangeParticleData();
  PIC::BC::ExternalBoundary::Periodic::ExchangeParticles();
}
// ---- the field half of ECSIM::TimeStep (pic_field_solver_ecsim.cpp:6449-6560): UpdateRhs, UpdateMatrixElement, the GMRES solve for
// x = E_conv (E^{n+theta} - E^n), ProcessFinalSolution, UpdateB, UpdateE.  (Its first call also runs UpdateJMassMatrix + BuildMatrix.)
int ref_pic_field_step(double tol, int max_iter, double *rel_residual) {
  PIC::CPLR::FLUID::EFieldTol = tol;
  PIC::CPLR::FLUID::EFieldIter = max_iter;
  PIC::FieldSolver::Electromagnetic::ECSIM::TimeStep();
  if (rel_residual) *rel_residual = ref_gmres_last_relative_residual;
  return ref_gmres_last_iterations;
}
// the rows of the reference's linear system in its own order: block (index of ref_pic_blocks), corner i, j, k, component and the
// right-hand side of the last UpdateRhs; returns the number of rows
long ref_pic_solver_rows(long max_n, int *block, int *ijk, int *ivar, double *rhs) {
  auto *S = PIC::FieldSolver::Electromagnetic::ECSIM::Solver;
  long n = 0;
  for (auto *row = S->MatrixRowListFirst; row != NULL; row = row->next, n++) {
    if (n >= max_n) continue;
    int b = -1;
    for (size_t q = 0; q < g_blocks.size(); q++)
      if (g_blocks[q] == row->node) b = (int)q;
    block[n] = b, ijk[3 * n] = row->i, ijk[3 * n + 1] = row->j, ijk[3 * n + 2] = row->k, ivar[n] = row->iVar, rhs[n] = row->Rhs;
  }
  return n;
}
// y = A x with the reference's operator (vectors in the order of ref_pic_solver_rows: 3 components per corner)
void ref_pic_matvec(double *x, double *y, int n) { PIC::FieldSolver::Electromagnetic::ECSIM::matvec(x, y, n); }
// the reference's constant operator tables (InitDiscritizationStencil, pic_field_solver_ecsim.cpp:7451): kind 0 = LaplacianStencil[p],
// kind 1 = GradDivStencil[p][q]; returns the number of taps (offsets di,dj,dk and the coefficient a)
int ref_pic_stencil(int kind, int p, int q, int max_n, int *ijk, double *a) {
  using namespace PIC::FieldSolver::Electromagnetic::ECSIM;
  cStencil::cStencilData *st = (kind == 0) ? LaplacianStencil + p : &GradDivStencil[p][q];
  for (int it = 0; it < st->Length && it < max_n; it++)
    ijk[3 * it] = st->Data[it].i, ijk[3 * it + 1] = st->Data[it].j, ijk[3 * it + 2] = st->Data[it].k, a[it] = st->Data[it].a;
  return st->Length;
}
double ref_pic_theta(void) { return PIC::FieldSolver::Electromagnetic::ECSIM::theta; }
// the field getters the guiding-centre movers call when the field solver is ECSIM (pic_mover_guiding_center.cpp:103, :179-184, :727):
// E[n][3], B[n][3], gradB[n][9] at the points x[n][3], each inside block[n] (BranchBottomNodeList order)
void ref_pic_ecsim_fields(long n, const double *x, const int *block, double *E, double *B, double *gradB) {
  using namespace PIC::FieldSolver::Electromagnetic::ECSIM;
  for (long i = 0; i < n; i++) {
    Node *node = g_blocks[block[i]];
    double xx[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
    GetElectricField(E + 3 * i, xx, node);
    xx[0] = x[3 * i], xx[1] = x[3 * i + 1], xx[2] = x[3 * i + 2];
    GetMagneticField(B + 3 * i, xx, node);
    GetMagneticFieldGradient(gradB + 9 * i, xx, node);
  }
}

// (the per-species cfl maxima are locals of UpdateJMassMatrix: it prints them, "max cfl number for spec s :value")
void ref_pic_update_JM(double *energy) {
  PIC::FieldSolver::Electromagnetic::ECSIM::UpdateJMassMatrix();
  if (energy) *energy = TotalParticleEnergy;
}
}
