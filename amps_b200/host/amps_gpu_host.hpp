// amps_gpu_host.hpp -- C++ host layer above the C ABI (include/amps_gpu.h): the reference's entry points of this path
// on the reference's own data structures, so that AMPS calls these instead of its CPU loops.
//
//   PIC::ParticleBuffer            one byte buffer MaxNPart x ParticleDataLength, per-cell doubly linked lists headed by
//                                  block->FirstCellParticleTable (pic_pbuffer.cpp:41-222, pic.h:4547)  -> ParticleBufferView
//   PIC::Mover::MoveParticles()    pic_mover.cpp:580-1088 (UserDefinedMoverManager hook, pic.h:5919)   -> MoveParticles()
//   ECSIM::UpdateJMassMatrix()     pic_field_solver_ecsim.cpp:3244-3995                                  -> UpdateJMassMatrix()
//   PIC::Mover::SetBlock_E/B       pic_mover.cpp:86-166                                                  -> SetFields()
//   ECSIM::ComputeNetCharge()      pic_field_solver_ecsim.cpp:4690-4828                                  -> ComputeNetCharge()
//   ECSIM::CorrectParticleLocation :4440-4688 (with the species corner moments of ProcessCell :2270-2300) -> CorrectParticleLocation()
//   PIC::Sampling::SamplingManager pic.cpp:1045-1082, ProcessCell :705-990                               -> Sampling(), SampledData()
//
// Header only; needs nothing but amps_gpu.h and the C++ standard library.  Errors become std::runtime_error carrying
// amps_gpu_last_error (AMPS maps them to exit(__LINE__,__FILE__,msg)).  There is no CPU fallback.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "amps_gpu.h"

namespace amps_b200 {

// the caller's PIC::ParticleBuffer
struct ParticleBufferView {
  unsigned char *ParticleDataBuffer;  // PIC::ParticleBuffer::ParticleDataBuffer
  long int MaxNPart;                  // PIC::ParticleBuffer::MaxNPart
  amps_gpu_aos_layout layout;         // ParticleDataLength and the _PIC_PARTICLE_DATA__*_OFFSET_ macros
};

class EcsimHost {
 public:
  EcsimHost(const amps_gpu_config &cfg, const amps_gpu_mesh &mesh) : n_corners_(mesh.n_corners), n_species_(cfg.n_species) {
    n_cells_ = (int64_t)mesh.n_leaves * cfg.block_cells[0] * cfg.block_cells[1] * cfg.block_cells[2];
    int rc = amps_gpu_init(&cfg, &ctx_);
    if (rc != AMPS_GPU_OK) {
      std::string msg = ctx_ ? amps_gpu_last_error(ctx_) : "amps_gpu_init failed (no CUDA device? there is no CPU fallback)";
      if (ctx_) amps_gpu_finalize(ctx_);
      ctx_ = nullptr;
      throw std::runtime_error(msg);
    }
    check(amps_gpu_mesh_upload(ctx_, &mesh));
  }
  ~EcsimHost() {
    if (ctx_) amps_gpu_finalize(ctx_);
  }
  EcsimHost(const EcsimHost &) = delete;
  EcsimHost &operator=(const EcsimHost &) = delete;

  // epoch start (after injection / restart / sampling touched the AoS buffer): walk every cell list like the reference's
  // loops do (pic_mover.cpp:934-980) and hand the records to the device store.  FirstCellParticleTable[cell], cell = leaf*Nx*Ny*Nz + i+Nx*(j+Ny*k)
  void UploadParticles(const ParticleBufferView &pb, const long int *FirstCellParticleTable) {
    std::vector<int64_t> ptrs;
    std::vector<int32_t> cells;
    for (int64_t c = 0; c < n_cells_; c++) {
      long int ptr = FirstCellParticleTable[c];
      while (ptr != -1) {
        if (ptr < 0 || ptr >= pb.MaxNPart) throw std::runtime_error("corrupt particle list");
        ptrs.push_back(ptr);
        cells.push_back((int32_t)c);
        int64_t next;
        __builtin_memcpy(&next, pb.ParticleDataBuffer + ptr * pb.layout.stride + pb.layout.off_next, 8);  // GetNext, pic.h:2808
        ptr = (long int)next;
      }
    }
    check(amps_gpu_particles_upload_aos(ctx_, pb.ParticleDataBuffer, ptrs.data(), cells.data(), (int64_t)ptrs.size(), &pb.layout));
  }

  // E^{n+theta} on the unique corner nodes, B^n / B^{n+1} on the unique centre (or corner) nodes
  void SetFields(const double *E_half, const double *B_prev, const double *B_cur) { check(amps_gpu_fields_upload(ctx_, E_half, B_prev, B_cur)); }

  // PIC::Mover::MoveParticles(): push every particle, periodic exchange, temp->first list swap (= the device sort)
  amps_gpu_move_stats MoveParticles(int mover = AMPS_MOVER_LAPENTA2017) {
    amps_gpu_move_stats st;
    check(amps_gpu_move(ctx_, mover, &st));
    check(amps_gpu_migrate(ctx_, nullptr, nullptr));
    check(amps_gpu_sort(ctx_));
    return st;
  }

  // ECSIM::UpdateJMassMatrix(): J[n_corners][3], M[n_corners][243] on the unique corners, TotalParticleEnergy, cfl per species
  void UpdateJMassMatrix(double *J, double *M, double *energy, double *cfl) {
    check(amps_gpu_deposit_JM(ctx_, energy, cfl));
    check(amps_gpu_exchange_JM(ctx_));
    if (energy || cfl) check(amps_gpu_diagnostics(ctx_, energy, cfl));
    check(amps_gpu_JM_download(ctx_, J, M));
  }

  // one call for a whole particle phase when nothing on the host needs the intermediate state
  void MoveAndDeposit(double *J, double *M, int mover = AMPS_MOVER_LAPENTA2017) { check(amps_gpu_step_JM(ctx_, mover, J, M)); }

  // the same with the packed rows JM[n_corners][129] (J + the 14 independent neighbour blocks, amps_gpu_JM_packed_slots):
  // half the PCIe volume; the caller mirrors the other 13 blocks while scattering into the corner buffers
  void MoveAndDepositPacked(double *JM129, int mover = AMPS_MOVER_LAPENTA2017) { check(amps_gpu_step_JM_packed(ctx_, mover, JM129)); }

  // ECSIM::ComputeNetCharge(): rho_new on the unique centre nodes.  The reference fills its global StencilTable here, after which its
  // movers leave full B stencils un-normalised (pic_interpolation_routines.cpp:903): the library is put into the same state
  void ComputeNetCharge(double charge_conv, double *rho_center) {
    check(amps_gpu_net_charge(ctx_, charge_conv, rho_center));
    check(amps_gpu_global_stencil_set(ctx_, 1));
  }

  // the guiding-centre species of PIC::GYROKINETIC (cfg.gc_species_mask): PB::GetMagneticMoment / PB::GetVNormal by ParticleBuffer slot
  void SetGuidingCentreState(const double *mu_by_ptr, const double *vnormal_by_ptr, int64_t n) {
    if (mu_by_ptr) check(amps_gpu_magnetic_moment_upload(ctx_, mu_by_ptr, n));
    if (vnormal_by_ptr) check(amps_gpu_v_normal_upload(ctx_, vnormal_by_ptr, n));
  }
  // the current E on the unique corners: what the guiding-centre movers read with cfg.gc_fields_ecsim (ECSIM::GetElectricField)
  void SetCurrentE(const double *E_cur) { check(amps_gpu_E_upload(ctx_, E_cur)); }

  // ECSIM::CorrectParticleLocation() with phi of the Poisson solve on the unique centre nodes; returns {shifted, deleted}.
  // The species corner moments it reads are sampled here (UpdateJMassMatrix does that in the reference); the lists are rebuilt.
  struct ShiftCount {
    int64_t shifted, deleted;
  };
  ShiftCount CorrectParticleLocation(const double *phi_center, double charge_conv, double mass_conv) {
    ShiftCount c{0, 0};
    check(amps_gpu_species_moments(ctx_, nullptr));
    check(amps_gpu_phi_upload(ctx_, phi_center));
    check(amps_gpu_correct_particle_location(ctx_, charge_conv, mass_conv, &c.shifted, &c.deleted));
    check(amps_gpu_migrate(ctx_, nullptr, nullptr));  // PIC::Parallel::ExchangeParticleData()
    check(amps_gpu_sort(ctx_));                        // exchangeParticleLocal(): the cell lists
    return c;
  }

  // PIC::Sampling: one more sample of the resident particles in the collecting buffer on the device
  void Sampling() { check(amps_gpu_sample_cells(ctx_)); }
  // the collecting buffer sample[n_cells][n_species][13] and the particles sampled per species; clear starts a new period
  void SampledData(double *sample, int64_t *n_sampled, bool clear) { check(amps_gpu_sample_download(ctx_, sample, n_sampled, clear ? 1 : 0)); }

  // epoch end on several ranks / with deleting boundaries: first settle the books of the caller's ParticleBuffer.
  // getNewParticle() -> long int is PIC::ParticleBuffer::GetNewParticle (pic_pbuffer.cpp:371) for every arrival of the
  // exchange, deleteParticle(long int) is PIC::ParticleBuffer::DeleteParticle (pic_pbuffer.cpp:594) for every record whose
  // particle left this rank or was deleted by a mover / boundary; then the records and lists come back.
  template <class GetNewParticle, class DeleteParticle>
  int64_t DownloadParticles(ParticleBufferView &pb, long int *FirstCellParticleTable, GetNewParticle getNewParticle, DeleteParticle deleteParticle) {
    int64_t nNew = 0, nRel = 0;
    check(amps_gpu_particles_slot_delta(ctx_, &nNew, nullptr, 0, &nRel));
    std::vector<int64_t> rel((size_t)nRel + 1);
    check(amps_gpu_particles_slot_delta(ctx_, &nNew, rel.data(), nRel, &nRel));
    for (int64_t i = 0; i < nRel; i++) deleteParticle((long int)rel[i]);  // first: arrivals may reuse these records
    std::vector<int64_t> fresh((size_t)nNew);
    for (int64_t i = 0; i < nNew; i++) fresh[i] = (int64_t)getNewParticle();
    check(amps_gpu_particles_assign_slots(ctx_, fresh.data(), nNew));
    return DownloadParticles(pb, FirstCellParticleTable);
  }

  // epoch end: records and lists back into the caller's buffer (single rank, nothing deleted: every record keeps its slot)
  int64_t DownloadParticles(ParticleBufferView &pb, long int *FirstCellParticleTable) {
    std::vector<int64_t> first((size_t)n_cells_);
    int64_t n = 0;
    check(amps_gpu_particles_download_aos(ctx_, pb.ParticleDataBuffer, first.data(), pb.MaxNPart, &pb.layout, &n));
    for (int64_t c = 0; c < n_cells_; c++) FirstCellParticleTable[c] = (long int)first[c];
    return n;
  }

  int64_t n_cells() const { return n_cells_; }
  int n_corners() const { return n_corners_; }
  amps_gpu_ctx *handle() { return ctx_; }

 private:
  void check(int rc) {
    if (rc != AMPS_GPU_OK) throw std::runtime_error(std::string("amps_gpu: ") + amps_gpu_last_error(ctx_));
  }
  amps_gpu_ctx *ctx_ = nullptr;
  int64_t n_cells_ = 0;
  int n_corners_ = 0, n_species_ = 0;
};

}  // namespace amps_b200
