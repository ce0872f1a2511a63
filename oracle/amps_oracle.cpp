/*
 * amps_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
 * See amps_oracle.h for scope and the parity-pinning statement.
 *
 * Every routine restates the reference algorithm literally (same operation
 * order, same data structures) and cites the reference file:line.  Build the
 * parity variant with  -O2 -ffp-contract=off  (no FMA contraction, SSE2 fp64),
 * the timing variant with -O3 -march=native -fopenmp.
 *
 * Reference paths are relative to the AMPS source tree.
 */
#include "amps_oracle.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef unsigned char byte;

// ---------------------------------------------------------------------------------------------
// node-associated data layout, src/pic/pic_field_solver_ecsim.cpp:484-538 (doubles from
// ElectricField.RelativeOffset):  E[0:3] E_half[3:6] J[6:9] M[9:252]; centre: B_cur[0:3] B_prev[3:6]
// (CurrentBOffset / PrevBOffset swap every step, :5697-5700; fixed here)
// ---------------------------------------------------------------------------------------------
const int ExOffsetIndex = 0;
const int OffsetE_HalfTimeStep_d = 3;  // in doubles
const int JxOffsetIndex = 6;
const int MassMatrixOffsetIndex = 9;
const int CornerDataLength = 252;
const int OffsetB_corner_d = 252;  // _PIC_FIELD_SOLVER_B_CORNER_BASED_: 6 more doubles per corner, B_cur[0:3] B_prev[3:6] (:534-538)
const int CurrentBOffset_d = 0;
const int PrevBOffset_d = 3;
const int BackgroundE_d = 6;   // coupler table: DATAFILE::Offset::ElectricField
const int BackgroundB_d = 9;   // coupler table: DATAFILE::Offset::MagneticField
const int BackgroundGCA_d = 12; // 15 tabulated derivative variables of the relativistic GCA (pic_datafile.cpp:1164-1340)
const int BackgroundGradB_d = 27; // DATAFILE::Offset::MagneticFieldGradient, 9 values {d/dx,d/dy,d/dz} of Bx, By, Bz (pic.h:8434-8470)
const int netChargeNew_d = 36;  // centre buffer: rho_new of the div-E correction (netChargeNewIndex, pic_field_solver_ecsim.cpp:484-538)
const int CenterDataLength = 37;
const double SpeedOfLight_SI = 299792458.0;  // src/general/constants.h:40

// src/pic/pic_field_solver_ecsim.cpp:1377-1380
const int IndexMatrix[8][8] = {{0, 2, 8, 6, 18, 20, 26, 24},  {1, 0, 6, 7, 19, 18, 24, 25},
                               {4, 3, 0, 1, 22, 21, 18, 19},  {3, 5, 2, 0, 21, 23, 20, 18},
                               {9, 11, 17, 15, 0, 2, 8, 6},   {10, 9, 15, 16, 1, 0, 6, 7},
                               {13, 12, 9, 10, 4, 3, 0, 1},   {12, 14, 11, 9, 3, 5, 2, 0}};

// particle-mover return codes, src/pic/pic.h:5955-5960
const int _PARTICLE_LEFT_THE_DOMAIN_ = 2;
const int _PARTICLE_MOTION_FINISHED_ = 3;
const int _PARTICLE_IN_NOT_IN_USE_NODE_ = 4;
const int _ORACLE_ERROR_ = -1;

struct cCornerNode {
  double *data;  // associated data buffer (CornerDataLength doubles)
  std::atomic_flag lock_associated_data = ATOMIC_FLAG_INIT;
};
struct cCenterNode {
  double *data;  // CenterDataLength doubles
};

struct cTempList {
  long int first, last;
};

struct cBlock {
  long int *FirstCellParticleTable;         // src/pic/pic.h:4547
  long int *tempParticleMovingListTable;    // src/pic/pic.h:4582-4586
  cTempList *tempThreadTable;               // [thread][cell] (hybrid mode)
  cCornerNode **cornerNodes;                // [(Nx+2g+1)(Ny+2g+1)(Nz+2g+1)]
  cCenterNode **centerNodes;                // [(Nx+2g)(Ny+2g)(Nz+2g)]
};

// cTreeNodeAMR, src/meshAMR/meshAMRgeneric.h:825-838
struct cTreeNode {
  double xmin[3], xmax[3];
  int xMinGlobalIndex[3];
  int NodeGeometricSizeIndex;
  cTreeNode *downNode[8];
  cTreeNode *upNode;
  int RefinmentLevel;
  int minNeibRefinmentLevel, maxNeibRefinmentLevel;  // meshAMRgeneric.h:829, SetNeibRefinmentLevelLimits :1018-1048
  int Thread;
  bool IsUsedInCalculationFlag;
  bool IsGhostNodeFlag;
  cBlock *block;
  int leaf;
  int faceBoundary;
  int id;
};

// cStencilGeneric, src/pic/pic.h:7202-7293
const int nMaxStencilLength = 64;
struct cStencil {
  int Length;
  double Weight[nMaxStencilLength];
  int LocalCellID[nMaxStencilLength];
  cCenterNode *cell[nMaxStencilLength];  // centre stencils of the coupler (cells of several blocks on AMR meshes)
  bool overflow;                         // the reference exit()s when Length would exceed nMaxStencilLength
  cStencil() : Length(0), overflow(false) {}
  void flush() { Length = 0, overflow = false; }
  void MultiplyScalar(double a) { for (int i = 0; i < Length; i++) Weight[i] *= a; }
  void Normalize() {
    double norm = 0.0;
    int i;
    for (i = 0; i < Length; i++) norm += Weight[i];
    if (norm > 0.0)
      for (i = 0; i < Length; i++) Weight[i] /= norm;
  }
};

// cCellData, src/pic/pic.h (ECSIM::cCellData): per-corner J[3], M[243]
struct cCornerData {
  double *CornerMassMatrix_ptr, *CornerJ_ptr;
  double CornerMassMatrix[243];
  double CornerJ[3];
  cCornerNode *CornerNode;
};
struct cCellData {
  cCornerData CornerData[8];
  double ParticleEnergy;
  double cflCell[AMPS_GPU_MAX_SPECIES];
  void clean() {
    for (int ic = 0; ic < 8; ic++) {
      for (int i = 0; i < 243; i++) CornerData[ic].CornerMassMatrix[i] = 0.0;
      for (int i = 0; i < 3; i++) CornerData[ic].CornerJ[i] = 0.0;
    }
    ParticleEnergy = 0.0;
    for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++) cflCell[s] = 0.0;
  }
};

}  // namespace

struct oracle_ctx {
  amps_gpu_config cfg;
  std::string err;
  // compile-time macros of the reference
  int _BLOCK_CELLS_X_, _BLOCK_CELLS_Y_, _BLOCK_CELLS_Z_;
  int _GHOST_CELLS_X_, _GHOST_CELLS_Y_, _GHOST_CELLS_Z_;
  int _TOTAL_BLOCK_CELLS_X_, _TOTAL_BLOCK_CELLS_Y_, _TOTAL_BLOCK_CELLS_Z_;
  // cMeshAMRgeneric
  double xGlobalMin[3], xGlobalMax[3], dx_max_refinment[3], dxRootBlock[3], EPS;
  int nRoot[3], maxRefinementLevel;
  std::vector<cTreeNode> nodes;
  std::vector<cTreeNode *> rootTable;
  std::vector<cTreeNode *> BlockTable;  // DomainBlockDecomposition::BlockTable (leaf order)
  std::vector<cBlock> blocks;
  std::vector<int> leaf_real;
  int n_corners, n_centers;
  std::vector<cCornerNode> cornerPool;
  std::vector<cCenterNode> centerPool;
  std::vector<double> cornerData, centerData;
  std::vector<double> cornerSpec;  // [n_corners][10*n_species]: corner buffer from SpeciesDataIndex[0] on (:527-531)
  std::vector<double> phiCenter;   // [n_centers]: centre buffer, phiIndex (div-E correction potential)
  std::vector<double> cellSample;  // [n_leaves*cells][n_species][13]: the collecting sampling buffer of the cells (pic.h:4117-4130)
  std::vector<long long> sampledParticles;  // localSimulatedSpeciesParticleNumber
  std::vector<long int> listStorage;
  std::vector<cTempList> threadListStorage;
  int nThreadListTables;
  // PIC::ParticleBuffer (src/pic/pic_pbuffer.cpp)
  long int MaxNPart, ParticleDataLength, FirstPBufferParticle, NAllPart;
  long long nSubSteps = 0;  // Relativistic::Boris sub-steps of the current move (statistic)
  byte *ParticleDataBuffer;
  int globalStencilLength = 0;  // Length of the stencil last built through the reference's global StencilTable (see GetTriliniarInterpolationStencil)
  // per-thread E/B staging, src/pic/pic_mover.cpp:660-711
  std::vector<std::vector<double>> E_Corner, B_Center;

  int nCornerLocal() const { return (_TOTAL_BLOCK_CELLS_X_ + 1) * (_TOTAL_BLOCK_CELLS_Y_ + 1) * (_TOTAL_BLOCK_CELLS_Z_ + 1); }
  int nCenterLocal() const { return _TOTAL_BLOCK_CELLS_X_ * _TOTAL_BLOCK_CELLS_Y_ * _TOTAL_BLOCK_CELLS_Z_; }
  int nCellsBlock() const { return _BLOCK_CELLS_X_ * _BLOCK_CELLS_Y_ * _BLOCK_CELLS_Z_; }
  // src/meshAMR/meshAMRgeneric.h:74-75
  int _getCornerNodeLocalNumber(int i, int j, int k) const {
    return (i + _GHOST_CELLS_X_ + (1 + _TOTAL_BLOCK_CELLS_X_) * (j + _GHOST_CELLS_Y_ + (k + _GHOST_CELLS_Z_) * (1 + _TOTAL_BLOCK_CELLS_Y_)));
  }
  int _getCenterNodeLocalNumber(int i, int j, int k) const {
    return (i + _GHOST_CELLS_X_ + _TOTAL_BLOCK_CELLS_X_ * (j + _GHOST_CELLS_Y_ + (k + _GHOST_CELLS_Z_) * _TOTAL_BLOCK_CELLS_Y_));
  }

  // ---- PIC::ParticleBuffer accessors, packed layout picParticleDataMacro.h:55-81 ----
  // + _PIC_PARTICLE_DATA__MAGNETIC_MOMENT_OFFSET_ (picParticleDataMacro.h:178-187) right after the basic data
  // + _PIC_PARTICLE_DATA__V_PARALLEL_OFFSET_ of the gyrokinetic reduced state behind it
  // + _PIC_PARTICLE_DATA__V_NORMAL_OFFSET_ (read by ProcessCell for guiding-centre species, pic_field_solver_ecsim.cpp:2233)
  enum { OFF_NEXT = 0, OFF_PREV = 8, OFF_SPEC = 16, OFF_V = 17, OFF_X = 41, OFF_W = 65, OFF_MU = 73, OFF_VPAR = 81, OFF_VNORMAL = 89, BASIC_LEN = 97 };
  static double GetVNormal(const byte *p) { double m; memcpy(&m, p + OFF_VNORMAL, 8); return m; }
  static void SetVNormal(double m, byte *p) { memcpy(p + OFF_VNORMAL, &m, 8); }
  static double GetVParallel(const byte *p) { double m; memcpy(&m, p + OFF_VPAR, 8); return m; }
  static void SetVParallel(double m, byte *p) { memcpy(p + OFF_VPAR, &m, 8); }
  static double GetMagneticMoment(const byte *p) { double m; memcpy(&m, p + OFF_MU, 8); return m; }
  static void SetMagneticMoment(double m, byte *p) { memcpy(p + OFF_MU, &m, 8); }
  byte *GetParticleDataPointer(long int ptr) const { return ParticleDataBuffer + ptr * ParticleDataLength; }
  static long int GetNext(const byte *p) { long int t; memcpy(&t, p + OFF_NEXT, 8); return t; }
  static long int GetPrev(const byte *p) { long int t; memcpy(&t, p + OFF_PREV, 8); return t; }
  static void SetNext(long int v, byte *p) { memcpy(p + OFF_NEXT, &v, 8); }
  static void SetPrev(long int v, byte *p) { memcpy(p + OFF_PREV, &v, 8); }
  long int GetNext(long int ptr) const { return GetNext(GetParticleDataPointer(ptr)); }
  void SetNext(long int v, long int ptr) { SetNext(v, GetParticleDataPointer(ptr)); }
  void SetPrev(long int v, long int ptr) { SetPrev(v, GetParticleDataPointer(ptr)); }
  // species byte: bits 0-5 id, bit 7 "allocated", src/pic/pic.h:2808-2900
  static unsigned int GetI(const byte *p) { return (*(p + OFF_SPEC)) & 0x3f; }
  static void SetI(int spec, byte *p) { *(p + OFF_SPEC) = (byte)((spec & 0x3f) | ((*(p + OFF_SPEC)) & 0xc0)); }
  // InitFlag: bit 6 of the species byte (pic.h:3668-3690)
  static bool TestInitFlag(const byte *p) { return ((*(p + OFF_SPEC)) & 0x40) != 0; }
  static void SetInitFlag(bool t, byte *p) { if (t) *(p + OFF_SPEC) |= 0x40; else *(p + OFF_SPEC) &= 0xbf; }
  static bool IsParticleAllocated(const byte *p) { return ((*(p + OFF_SPEC)) & 0x80) != 0; }
  static void SetParticleDeleted(byte *p) { *(p + OFF_SPEC) &= 0x7f; }
  static void SetParticleAllocated(byte *p) { *(p + OFF_SPEC) |= 0x80; }
  static void GetV(double *v, const byte *p) { memcpy(v, p + OFF_V, 24); }
  static void SetV(const double *v, byte *p) { memcpy(p + OFF_V, v, 24); }
  static void GetX(double *x, const byte *p) { memcpy(x, p + OFF_X, 24); }
  static void SetX(const double *x, byte *p) { memcpy(p + OFF_X, x, 24); }
  static double GetIndividualStatWeightCorrection(const byte *p) { double w; memcpy(&w, p + OFF_W, 8); return w; }
  static void SetIndividualStatWeightCorrection(double w, byte *p) { memcpy(p + OFF_W, &w, 8); }

  // src/pic/pic_pbuffer.cpp:371-437 (MPI mode branch)
  long int GetNewParticle() {
    if (FirstPBufferParticle == -1) return -1;
    long int newptr = FirstPBufferParticle;
    byte *p = GetParticleDataPointer(newptr);
    FirstPBufferParticle = GetNext(p);
    NAllPart++;
    SetParticleAllocated(p);
    SetPrev(-1, p);
    SetNext(-1, p);
    return newptr;
  }
  // src/pic/pic_pbuffer.cpp:594-666 (the particle is already detached from its cell list by the caller)
  void DeleteParticle_withoutTrajectoryTermination(long int ptr) {
    byte *p = GetParticleDataPointer(ptr);
    SetParticleDeleted(p);
#pragma omp critical(oracle_pbuffer)
    {
      SetNext(FirstPBufferParticle, p);
      FirstPBufferParticle = ptr;
      NAllPart--;
    }
  }
  void DeleteParticle(long int ptr) { DeleteParticle_withoutTrajectoryTermination(ptr); }

  // ---- cMeshAMRgeneric::findTreeNode(int*), src/meshAMR/meshAMRgeneric.h:2793-2848.
  // Forest extension: when the walk leaves a root block it continues in the root grid.
  cTreeNode *findTreeNode(int *ix, cTreeNode *startNode) const {
    int i = 0, j = 0, k = 0, idim;
    bool inblock;
    if (startNode == NULL) startNode = rootLookup(ix);
    if (startNode == NULL) return NULL;
    while (true) {
      inblock = true;
      for (idim = 0; idim < 3; idim++) {
        if ((ix[idim] < startNode->xMinGlobalIndex[idim]) || (ix[idim] >= startNode->xMinGlobalIndex[idim] + startNode->NodeGeometricSizeIndex)) {
          inblock = false;
          break;
        }
      }
      if (inblock == true) {
        i = (ix[0] - startNode->xMinGlobalIndex[0] < startNode->NodeGeometricSizeIndex / 2) ? 0 : 1;
        j = (ix[1] - startNode->xMinGlobalIndex[1] < startNode->NodeGeometricSizeIndex / 2) ? 0 : 1;
        k = (ix[2] - startNode->xMinGlobalIndex[2] < startNode->NodeGeometricSizeIndex / 2) ? 0 : 1;
        cTreeNode *t = startNode->downNode[i + 2 * (j + 2 * k)];
        if (t != NULL) {
          startNode = t;
          continue;
        } else
          return startNode;
      } else {
        if (startNode->upNode != 0) {
          startNode = startNode->upNode;
          continue;
        } else {
          // single octree: return NULL (meshAMRgeneric.h:2840-2842); forest: look in the root grid
          startNode = rootLookup(ix);
          if (startNode == NULL) return NULL;
          continue;
        }
      }
    }
  }
  cTreeNode *rootLookup(const int *ix) const {
    int S = 1 << maxRefinementLevel, r[3];
    for (int d = 0; d < 3; d++) {
      if (ix[d] < 0) return NULL;
      r[d] = ix[d] / S;
      if (r[d] >= nRoot[d]) return NULL;
    }
    return rootTable[r[0] + nRoot[0] * (r[1] + nRoot[1] * r[2])];
  }
  // cMeshAMRgeneric::findTreeNode(double*), src/meshAMR/meshAMRgeneric.h:2851-2882
  cTreeNode *findTreeNode(const double *x, cTreeNode *startNode) const {
    int idim, ix[3];
    cTreeNode *res;
    bool flag;
    for (idim = 0; idim < 3; idim++) {
      ix[idim] = (int)floor((x[idim] - xGlobalMin[idim]) / dx_max_refinment[idim]);
    }
    res = findTreeNode(ix, startNode);
    flag = false;
    if (res != NULL)
      for (idim = 0; idim < 3; idim++) {
        if (x[idim] < res->xmin[idim]) ix[idim]--, flag = true;
        if (x[idim] >= res->xmax[idim]) ix[idim]++, flag = true;
      }
    if (flag == true) {
      res = findTreeNode(ix, res);
    }
    return res;
  }
  // cMeshAMRgeneric::FindCellIndex, src/meshAMR/meshAMRgeneric.h:2256-2323 (ExitFlag=false)
  long int FindCellIndex(const double *x, int &i, int &j, int &k, const cTreeNode *startNode) const {
    double dx;
    if ((x[0] < startNode->xmin[0]) || (startNode->xmax[0] < x[0])) return -1;
    dx = dxRootBlock[0] / (1 << startNode->RefinmentLevel) / double(_BLOCK_CELLS_X_);
    i = (int)((x[0] - startNode->xmin[0]) / dx);
    if (i == _BLOCK_CELLS_X_) i = _BLOCK_CELLS_X_ - 1;
    if ((x[1] < startNode->xmin[1]) || (startNode->xmax[1] < x[1])) return -1;
    dx = dxRootBlock[1] / (1 << startNode->RefinmentLevel) / double(_BLOCK_CELLS_Y_);
    j = (int)((x[1] - startNode->xmin[1]) / dx);
    if (j == _BLOCK_CELLS_Y_) j = _BLOCK_CELLS_Y_ - 1;
    if ((x[2] < startNode->xmin[2]) || (startNode->xmax[2] < x[2])) return -1;
    dx = dxRootBlock[2] / (1 << startNode->RefinmentLevel) / double(_BLOCK_CELLS_Z_);
    k = (int)((x[2] - startNode->xmin[2]) / dx);
    if (k == _BLOCK_CELLS_Z_) k = _BLOCK_CELLS_Z_ - 1;
    return _getCenterNodeLocalNumber(i, j, k);
  }

  // cStencilGeneric::AddCell for centre nodes, src/pic/pic.h:7228-7252: in non-periodic mode a
  // node whose centre lies outside the global box is not added.  The centre coordinate is
  // rebuilt from the block geometry (cDataCenterNode::GetX()).
  bool CenterOutsideDomain(const cTreeNode *node, int i, int j, int k) const {
    if (cfg.periodic) return false;
    const int N[3] = {_BLOCK_CELLS_X_, _BLOCK_CELLS_Y_, _BLOCK_CELLS_Z_};
    const int ijk[3] = {i, j, k};
    for (int d = 0; d < 3; ++d) {
      const double x = node->xmin[d] + (ijk[d] + 0.5) * ((node->xmax[d] - node->xmin[d]) / N[d]);
      if (x < xGlobalMin[d] || x > xGlobalMax[d]) return true;
    }
    return false;
  }

  // PIC::InterpolationRoutines::CornerBased::InitStencil, src/pic/pic_interpolation_routines.cpp:1074-1194
  // returns false where the reference exit()s ("the point is out of block")
  bool CornerBased_InitStencil(double *x, cTreeNode *node, cStencil &Stencil, double *InterpolationCoefficientTable) const {
    int iStencil, jStencil, kStencil, iX[3], nd, idim;
    double w, xLoc[3], dx[3], *xMinNode, *xMaxNode;
    cCornerNode *CornerNode;
    cBlock *block;

    xMinNode = node->xmin;
    xMaxNode = node->xmax;
    dx[0] = (xMaxNode[0] - xMinNode[0]) / _BLOCK_CELLS_X_;
    dx[1] = (xMaxNode[1] - xMinNode[1]) / _BLOCK_CELLS_Y_;
    dx[2] = (xMaxNode[2] - xMinNode[2]) / _BLOCK_CELLS_Z_;

    for (idim = 0; idim < 3; idim++) {
      if ((x[idim] < xMinNode[idim]) || (x[idim] > xMaxNode[idim])) return false;
      if (fabs(x[idim] - xMaxNode[idim]) < 1e-10 * dx[idim]) x[idim] = xMaxNode[idim] - 1e-10 * dx[idim];
      xLoc[idim] = (x[idim] - xMinNode[idim]) / dx[idim];
      iX[idim] = (int)(xLoc[idim]);
      xLoc[idim] -= iX[idim];
    }

    Stencil.flush();
    if ((block = node->block) == NULL) return true;

    for (iStencil = 0; iStencil < 2; iStencil++)
      for (jStencil = 0; jStencil < 2; jStencil++)
        for (kStencil = 0; kStencil < 2; kStencil++) {
          nd = _getCornerNodeLocalNumber(iStencil + iX[0], jStencil + iX[1], kStencil + iX[2]);
          CornerNode = block->cornerNodes[nd];
          // cell-corner order of the table: (0,0,0)(1,0,0)(1,1,0)(0,1,0)(0,0,1)(1,0,1)(1,1,1)(0,1,1)
          static const int slot[8] = {0, 1, 3, 2, 4, 5, 7, 6};
          int code = iStencil + 2 * jStencil + 4 * kStencil;
          if (CornerNode != NULL) {
            switch (code) {
              case 0: w = (1.0 - xLoc[0]) * (1.0 - xLoc[1]) * (1.0 - xLoc[2]); break;
              case 1: w = xLoc[0] * (1.0 - xLoc[1]) * (1.0 - xLoc[2]); break;
              case 2: w = (1.0 - xLoc[0]) * xLoc[1] * (1.0 - xLoc[2]); break;
              case 3: w = xLoc[0] * xLoc[1] * (1.0 - xLoc[2]); break;
              case 4: w = (1.0 - xLoc[0]) * (1.0 - xLoc[1]) * xLoc[2]; break;
              case 5: w = xLoc[0] * (1.0 - xLoc[1]) * xLoc[2]; break;
              case 6: w = (1.0 - xLoc[0]) * xLoc[1] * xLoc[2]; break;
              default: w = xLoc[0] * xLoc[1] * xLoc[2]; break;
            }
            InterpolationCoefficientTable[slot[code]] = w;
            // AddCell (pic.h:7228): in-block corners are inside the global box
            Stencil.Weight[Stencil.Length] = w;
            Stencil.LocalCellID[Stencil.Length] = nd;
            Stencil.Length++;
          } else {
            InterpolationCoefficientTable[slot[code]] = 0.0;
          }
        }
    Stencil.Normalize();
    return true;
  }

  // CellCentered::Linear::GetTriliniarInterpolationStencil, src/pic/pic_interpolation_routines.cpp:820-907
  // always_normalize: the reference tests the GLOBAL StencilTable->Length (:903) although the movers and ProcessCell pass their own
  // stencil object: Normalize() runs unless the global table holds an 8-cell stencil.  Only ComputeNetCharge fills the global table
  // (the StencilTable overload, :4783), so without the div-E correction Normalize() always runs; after a ComputeNetCharge whose last
  // particle had a full stencil it never does (globalStencilLength below; found by comparing two consecutive moves with the
  // reference-compiled code, tests/test_reference_gyrokinetic.py).
  void GetTriliniarInterpolationStencil(double iLoc, double jLoc, double kLoc, const double *x, cTreeNode *node, cStencil &Stencil, bool always_normalize) const {
    cCenterNode *cell;
    cBlock *block = node->block;
    Stencil.flush();
    double w[3], InterpolationWeight;
    int i, j, k, i0, j0, k0, nd;

    i0 = (iLoc < 0.5) ? -1 : (int)(iLoc - 0.50);
    j0 = (jLoc < 0.5) ? -1 : (int)(jLoc - 0.50);
    k0 = (kLoc < 0.5) ? -1 : (int)(kLoc - 0.50);

    w[0] = iLoc - (i0 + 0.5);
    w[1] = jLoc - (j0 + 0.5);
    w[2] = kLoc - (k0 + 0.5);

    for (i = 0; i < 2; i++)
      for (j = 0; j < 2; j++)
        for (k = 0; k < 2; k++) {
          nd = _getCenterNodeLocalNumber(i0 + i, j0 + j, k0 + k);
          switch (i + 2 * j + 4 * k) {
            case 0: InterpolationWeight = (1.0 - w[0]) * (1.0 - w[1]) * (1.0 - w[2]); break;
            case 1: InterpolationWeight = w[0] * (1.0 - w[1]) * (1.0 - w[2]); break;
            case 2: InterpolationWeight = (1.0 - w[0]) * w[1] * (1.0 - w[2]); break;
            case 3: InterpolationWeight = w[0] * w[1] * (1.0 - w[2]); break;
            case 4: InterpolationWeight = (1.0 - w[0]) * (1.0 - w[1]) * w[2]; break;
            case 5: InterpolationWeight = w[0] * (1.0 - w[1]) * w[2]; break;
            case 6: InterpolationWeight = (1.0 - w[0]) * w[1] * w[2]; break;
            default: InterpolationWeight = w[0] * w[1] * w[2]; break;
          }
          cell = (block == NULL) ? NULL : block->centerNodes[nd];
          if (cell != NULL) {
            if (CenterOutsideDomain(node, i0 + i, j0 + j, k0 + k)) continue;  // AddCell, pic.h:7235-7245
            Stencil.Weight[Stencil.Length] = InterpolationWeight;
            Stencil.LocalCellID[Stencil.Length] = nd;
            Stencil.Length++;
          }
        }

    if (Stencil.Length == 0) {
      // Constant::InitStencil fallback, pic_interpolation_routines.cpp:168-220
      int ci, cj, ck;
      long int cnd = FindCellIndex(x, ci, cj, ck, node);
      if (cnd >= 0) {
        Stencil.Weight[0] = 1.0;
        Stencil.LocalCellID[0] = (int)cnd;
        Stencil.Length = 1;
      }
      return;
    } else if (always_normalize || Stencil.Length != 8) {
      Stencil.Normalize();
    }
  }
  // CellCentered::Linear::InitStencil, src/pic/pic_interpolation_routines.cpp:224-330
  // (uniform mesh / all neighbours at the same level branch)
  void CellCentered_Linear_InitStencil(const double *XyzIn_D, cTreeNode *node, cStencil &Stencil, bool always_normalize) const {
    double iLoc, jLoc, kLoc;
    double *xmin = node->xmin, *xmax = node->xmax;
    iLoc = (XyzIn_D[0] - xmin[0]) / (xmax[0] - xmin[0]) * _BLOCK_CELLS_X_;
    jLoc = (XyzIn_D[1] - xmin[1]) / (xmax[1] - xmin[1]) * _BLOCK_CELLS_Y_;
    kLoc = (XyzIn_D[2] - xmin[2]) / (xmax[2] - xmin[2]) * _BLOCK_CELLS_Z_;
    GetTriliniarInterpolationStencil(iLoc, jLoc, kLoc, XyzIn_D, node, Stencil, always_normalize);
  }


  // ------------------------------------------------------------------------------------------
  // a4, AMR branch: neighbours by lattice probe (meshAMRgeneric.h:505-725), neighbour level limits (:1018-1048)
  // ------------------------------------------------------------------------------------------
  cTreeNode *neibNodeCorner(cTreeNode *n, int i) const {
    int ix[3];
    for (int d = 0; d < 3; d++) ix[d] = ((i >> d) & 1) ? n->xMinGlobalIndex[d] + n->NodeGeometricSizeIndex : n->xMinGlobalIndex[d] - 1;
    return findTreeNode(ix, n);
  }
  cTreeNode *neibNodeFace(cTreeNode *n, int i) const {
    int ix[3];
    for (int d = 0; d < 3; d++) ix[d] = n->xMinGlobalIndex[d];
    int nface = i / 4;
    i -= 4 * nface;
    int jFace = i / 2, iFace = i % 2;
    const int dn = nface / 2;                                // normal direction
    const int t0 = (dn == 0) ? 1 : 0, t1 = (dn == 2) ? 1 : 2;  // tangential directions in the reference's order
    ix[dn] -= 1;
    if ((iFace == 1) && (n->NodeGeometricSizeIndex > 1)) ix[t0] += n->NodeGeometricSizeIndex / 2;
    if ((jFace == 1) && (n->NodeGeometricSizeIndex > 1)) ix[t1] += n->NodeGeometricSizeIndex / 2;
    if (nface & 1) ix[dn] += 1 + n->NodeGeometricSizeIndex;
    return findTreeNode(ix, n);
  }
  cTreeNode *neibNodeEdge(cTreeNode *n, int i) const {
    static const int Increment[12][3] = {{0, -1, -1}, {0, 1, -1}, {0, 1, 1}, {0, -1, 1}, {-1, 0, -1}, {1, 0, -1},
                                         {1, 0, 1},   {-1, 0, 1}, {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0}};
    static const int Direction[12] = {0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2};
    int ix[3];
    int iedge = i / 2, isegment = i % 2;
    for (int idim = 0; idim < 3; idim++) {
      ix[idim] = n->xMinGlobalIndex[idim];
      if (Increment[iedge][idim] == -1) ix[idim] -= 1;
      else if (Increment[iedge][idim] == 1) ix[idim] += n->NodeGeometricSizeIndex;
    }
    if (isegment == 1)
      if (n->NodeGeometricSizeIndex > 1) ix[Direction[iedge]] += n->NodeGeometricSizeIndex / 2;
    return findTreeNode(ix, n);
  }
  cTreeNode *GetNeibFace(cTreeNode *n, int nface, int iFace, int jFace) const { return neibNodeFace(n, iFace + 2 * (jFace + 2 * nface)); }
  cTreeNode *GetNeibEdge(cTreeNode *n, int nedge, int iEdge) const { return neibNodeEdge(n, iEdge + 2 * nedge); }
  cTreeNode *GetNeibCorner(cTreeNode *n, int c) const { return neibNodeCorner(n, c); }
  void SetNeibRefinmentLevelLimits(cTreeNode *n) const {
    cTreeNode *node;
    int i;
    n->minNeibRefinmentLevel = -1, n->maxNeibRefinmentLevel = -1;
    for (i = 0; i < 6 * 4; i++)
      if ((node = neibNodeFace(n, i)) != NULL) {
        if ((n->minNeibRefinmentLevel == -1) || (n->minNeibRefinmentLevel > node->RefinmentLevel)) n->minNeibRefinmentLevel = node->RefinmentLevel;
        if (n->maxNeibRefinmentLevel < node->RefinmentLevel) n->maxNeibRefinmentLevel = node->RefinmentLevel;
      }
    for (i = 0; i < 8; i++)
      if ((node = neibNodeCorner(n, i)) != NULL) {
        if ((n->minNeibRefinmentLevel == -1) || (n->minNeibRefinmentLevel > node->RefinmentLevel)) n->minNeibRefinmentLevel = node->RefinmentLevel;
        if (n->maxNeibRefinmentLevel < node->RefinmentLevel) n->maxNeibRefinmentLevel = node->RefinmentLevel;
      }
    for (i = 0; i < 12 * 2; i++)
      if ((node = neibNodeEdge(n, i)) != NULL) {
        if ((n->minNeibRefinmentLevel == -1) || (n->minNeibRefinmentLevel > node->RefinmentLevel)) n->minNeibRefinmentLevel = node->RefinmentLevel;
        if (n->maxNeibRefinmentLevel < node->RefinmentLevel) n->maxNeibRefinmentLevel = node->RefinmentLevel;
      }
  }

  // cStencilGeneric::AddCell (pic.h:7228-7252) for the centre (i,j,k) of `node`
  void CplrAddCell(cStencil &S, double w, cTreeNode *node, int i, int j, int k, cCenterNode *c, int id) const {
    if (S.Length == nMaxStencilLength) {
      S.overflow = true;
      return;
    }
    if (CenterOutsideDomain(node, i, j, k)) return;
    S.Weight[S.Length] = w;
    S.cell[S.Length] = c;
    S.LocalCellID[S.Length] = id;
    S.Length++;
  }
  // AddPhysicalStencilCell, pic_interpolation_routines.cpp:124-134
  void AddPhysicalStencilCell(cStencil &S, double Weight, cTreeNode *node, int i, int j, int k, cCenterNode *cell, int LocalCellID) const {
    for (int iElement = 0; iElement < S.Length; iElement++)
      if (S.cell[iElement] == cell) {
        S.Weight[iElement] += Weight;
        return;
      }
    CplrAddCell(S, Weight, node, i, j, k, cell, LocalCellID);
  }
  // cStencilGeneric::Add, pic.h:7274-7292 (the position test of AddCell already passed when t was built)
  void StencilAdd(cStencil &S, const cStencil &t) const {
    for (int i = 0; i < t.Length; i++) {
      cCenterNode *el = t.cell[i];
      bool flag = false;
      for (int j = 0; j < S.Length; j++)
        if (S.cell[j] == el) {
          flag = true;
          S.Weight[j] += t.Weight[i];
          break;
        }
      if (flag == false) {
        if (S.Length == nMaxStencilLength) {
          S.overflow = true;
          return;
        }
        S.Weight[S.Length] = t.Weight[i], S.cell[S.Length] = el, S.LocalCellID[S.Length] = t.LocalCellID[i];
        S.Length++;
      }
    }
  }
  // Constant::InitStencil (:168-220); false where the reference exit()s
  bool CplrConstantStencil(const double *x, cTreeNode *node, cStencil &S) const {
    S.flush();
    if (node == NULL || node->block == NULL) return false;
    int i, j, k;
    long int nd = FindCellIndex(x, i, j, k, node);
    if (nd < 0) return false;
    cCenterNode *cell = node->block->centerNodes[nd];
    if (cell == NULL) return false;
    CplrAddCell(S, 1.0, node, i, j, k, cell, (int)nd);
    return true;
  }
  // GetTriliniarInterpolationStencil (:820-909) with cell pointers; StencilTable = the object the reference tests at :903
  bool CplrTrilinearStencil(double iLoc, double jLoc, double kLoc, const double *x, cTreeNode *node, cStencil &Stencil, const cStencil *StencilTable) const {
    cBlock *block = node->block;
    if (block == NULL) return false;
    Stencil.flush();
    double w[3], InterpolationWeight;
    int i, j, k, i0, j0, k0, nd;
    i0 = (iLoc < 0.5) ? -1 : (int)(iLoc - 0.50);
    j0 = (jLoc < 0.5) ? -1 : (int)(jLoc - 0.50);
    k0 = (kLoc < 0.5) ? -1 : (int)(kLoc - 0.50);
    // indices past the ghost layer are out-of-bounds reads in the reference
    if (i0 < -_GHOST_CELLS_X_ || i0 + 1 > _BLOCK_CELLS_X_ + _GHOST_CELLS_X_ - 1 || j0 < -_GHOST_CELLS_Y_ || j0 + 1 > _BLOCK_CELLS_Y_ + _GHOST_CELLS_Y_ - 1 ||
        k0 < -_GHOST_CELLS_Z_ || k0 + 1 > _BLOCK_CELLS_Z_ + _GHOST_CELLS_Z_ - 1)
      return false;
    w[0] = iLoc - (i0 + 0.5);
    w[1] = jLoc - (j0 + 0.5);
    w[2] = kLoc - (k0 + 0.5);
    for (i = 0; i < 2; i++)
      for (j = 0; j < 2; j++)
        for (k = 0; k < 2; k++) {
          nd = _getCenterNodeLocalNumber(i0 + i, j0 + j, k0 + k);
          switch (i + 2 * j + 4 * k) {
            case 0: InterpolationWeight = (1.0 - w[0]) * (1.0 - w[1]) * (1.0 - w[2]); break;
            case 1: InterpolationWeight = w[0] * (1.0 - w[1]) * (1.0 - w[2]); break;
            case 2: InterpolationWeight = (1.0 - w[0]) * w[1] * (1.0 - w[2]); break;
            case 3: InterpolationWeight = w[0] * w[1] * (1.0 - w[2]); break;
            case 4: InterpolationWeight = (1.0 - w[0]) * (1.0 - w[1]) * w[2]; break;
            case 5: InterpolationWeight = w[0] * (1.0 - w[1]) * w[2]; break;
            case 6: InterpolationWeight = (1.0 - w[0]) * w[1] * w[2]; break;
            default: InterpolationWeight = w[0] * w[1] * w[2]; break;
          }
          cCenterNode *cell = block->centerNodes[nd];
          if (cell != NULL) CplrAddCell(Stencil, InterpolationWeight, node, i0 + i, j0 + j, k0 + k, cell, nd);
        }
    if (Stencil.Length == 0) return CplrConstantStencil(x, node, Stencil);
    else if (StencilTable->Length != 8) Stencil.Normalize();
    return true;
  }
  // GetTriliniarInterpolationMutiBlockStencil (:912-1070)
  bool CplrMultiBlockStencil(const double *x, cTreeNode *node, cStencil &Stencil) const {
    long int nd;
    cCenterNode *cell;
    Stencil.flush();
    const int nCells[3] = {_BLOCK_CELLS_X_, _BLOCK_CELLS_Y_, _BLOCK_CELLS_Z_};
    const int nGhost[3] = {_GHOST_CELLS_X_, _GHOST_CELLS_Y_, _GHOST_CELLS_Z_};
    double dxCell[3];
    int idim;
    for (idim = 0; idim < 3; idim++) dxCell[idim] = (node->xmax[idim] - node->xmin[idim]) / nCells[idim];
    int ijkStencilMin[3];
    double xStencilLower[3], xLoc[3];
    for (idim = 0; idim < 3; idim++) {
      ijkStencilMin[idim] = (x[idim] - node->xmin[idim] < 0.5 * dxCell[idim]) ? -1 : (int)((x[idim] - node->xmin[idim] - 0.5 * dxCell[idim]) / dxCell[idim]);
      xStencilLower[idim] = node->xmin[idim] + (ijkStencilMin[idim] + 0.5) * dxCell[idim];
      xLoc[idim] = (x[idim] - xStencilLower[idim]) / dxCell[idim];
      if (xLoc[idim] < 0.0) xLoc[idim] = 0.0;
      if (xLoc[idim] > 1.0) xLoc[idim] = 1.0;
    }
    bool GeometryAvailable = true;
    bool PhysicalStencilAvailable = true;
    for (int di = 0; di < 2; di++)
      for (int dj = 0; dj < 2; dj++)
        for (int dk = 0; dk < 2; dk++) {
          int ijk[3] = {ijkStencilMin[0] + di, ijkStencilMin[1] + dj, ijkStencilMin[2] + dk};
          double xLogical[3];
          for (idim = 0; idim < 3; idim++) xLogical[idim] = node->xmin[idim] + (ijk[idim] + 0.5) * dxCell[idim];
          double StencilElementWeight;
          switch (di + 2 * dj + 4 * dk) {
            case 0: StencilElementWeight = (1.0 - xLoc[0]) * (1.0 - xLoc[1]) * (1.0 - xLoc[2]); break;
            case 1: StencilElementWeight = xLoc[0] * (1.0 - xLoc[1]) * (1.0 - xLoc[2]); break;
            case 2: StencilElementWeight = (1.0 - xLoc[0]) * xLoc[1] * (1.0 - xLoc[2]); break;
            case 3: StencilElementWeight = xLoc[0] * xLoc[1] * (1.0 - xLoc[2]); break;
            case 4: StencilElementWeight = (1.0 - xLoc[0]) * (1.0 - xLoc[1]) * xLoc[2]; break;
            case 5: StencilElementWeight = xLoc[0] * (1.0 - xLoc[1]) * xLoc[2]; break;
            case 6: StencilElementWeight = (1.0 - xLoc[0]) * xLoc[1] * xLoc[2]; break;
            default: StencilElementWeight = xLoc[0] * xLoc[1] * xLoc[2]; break;
          }
          cTreeNode *StencilNode = findTreeNode(xLogical, node);
          if ((StencilNode == NULL) || (StencilNode->IsUsedInCalculationFlag == false)) {
            GeometryAvailable = false;
            continue;
          }
          if (StencilNode->RefinmentLevel == node->RefinmentLevel) {
            bool inTile = true;
            for (idim = 0; idim < 3; idim++)
              if (ijk[idim] < -nGhost[idim] || ijk[idim] > nCells[idim] + nGhost[idim] - 1) inTile = false;
            nd = inTile ? _getCenterNodeLocalNumber(ijk[0], ijk[1], ijk[2]) : -1;
            cell = (node->block == NULL || !inTile) ? NULL : node->block->centerNodes[nd];
            if (cell != NULL) AddPhysicalStencilCell(Stencil, StencilElementWeight, node, ijk[0], ijk[1], ijk[2], cell, (int)nd);
            else PhysicalStencilAvailable = false;
          } else {
            int iNeib[3];
            for (idim = 0; idim < 3; idim++) iNeib[idim] = 2 * ((int)((xLogical[idim] - StencilNode->xmin[idim]) / dxCell[idim]));
            for (int ii = 0; ii < 2; ii++)
              for (int jj = 0; jj < 2; jj++)
                for (int kk = 0; kk < 2; kk++) {
                  const int iFine = iNeib[0] + ii, jFine = iNeib[1] + jj, kFine = iNeib[2] + kk;
                  const double FineCellWeight = (1.0 / 8.0) * StencilElementWeight;
                  const int f[3] = {iFine, jFine, kFine};
                  bool inTile = true;
                  for (idim = 0; idim < 3; idim++)
                    if (f[idim] < -nGhost[idim] || f[idim] > nCells[idim] + nGhost[idim] - 1) inTile = false;
                  nd = inTile ? _getCenterNodeLocalNumber(iFine, jFine, kFine) : -1;
                  cell = (StencilNode->block == NULL || !inTile) ? NULL : StencilNode->block->centerNodes[nd];
                  if (cell != NULL) AddPhysicalStencilCell(Stencil, FineCellWeight, StencilNode, iFine, jFine, kFine, cell, (int)nd);
                  else PhysicalStencilAvailable = false;
                }
          }
        }
    if (GeometryAvailable == false) {
      cTreeNode *InterpolationNode = findTreeNode(x, node);
      if ((InterpolationNode != NULL) && (InterpolationNode->block != NULL)) return CplrConstantStencil(x, InterpolationNode, Stencil);
      Stencil.flush();
      return true;
    }
    if (PhysicalStencilAvailable == false) {
      cTreeNode *InterpolationNode = findTreeNode(x, node);
      if ((InterpolationNode != NULL) && (InterpolationNode->block != NULL)) return CplrConstantStencil(x, InterpolationNode, Stencil);
      Stencil.flush();
    }
    return true;
  }
  // CellCentered::Linear::InitStencil (:224-706), _PIC_CELL_CENTERED_LINEAR_INTERPOLATION_ROUTINE__AMPS_, non-uniform mesh type.
  // Fills Stencil.cell[]; false where the reference exit()s / reads out of bounds.
  bool CplrLinearStencil(const double *XyzIn_D, cTreeNode *node, cStencil &Stencil) const {
    Stencil.flush();
    if (node == NULL || node->block == NULL) return false;
    double iLoc, jLoc, kLoc;
    double *xmin = node->xmin, *xmax = node->xmax;
    iLoc = (XyzIn_D[0] - xmin[0]) / (xmax[0] - xmin[0]) * _BLOCK_CELLS_X_;
    jLoc = (XyzIn_D[1] - xmin[1]) / (xmax[1] - xmin[1]) * _BLOCK_CELLS_Y_;
    kLoc = (XyzIn_D[2] - xmin[2]) / (xmax[2] - xmin[2]) * _BLOCK_CELLS_Z_;
    if (!(iLoc >= -1.0e9 && iLoc <= 1.0e9 && jLoc >= -1.0e9 && jLoc <= 1.0e9 && kLoc >= -1.0e9 && kLoc <= 1.0e9)) return false;
    if ((node->RefinmentLevel == node->minNeibRefinmentLevel) && (node->RefinmentLevel == node->maxNeibRefinmentLevel)) {
      return CplrTrilinearStencil(iLoc, jLoc, kLoc, XyzIn_D, node, Stencil, &Stencil);
    } else if ((1.0 < iLoc) && (iLoc < _BLOCK_CELLS_X_ - 1) && (1.0 < jLoc) && (jLoc < _BLOCK_CELLS_Y_ - 1) && (1.0 < kLoc) && (kLoc < _BLOCK_CELLS_Z_ - 1)) {
      return CplrTrilinearStencil(iLoc, jLoc, kLoc, XyzIn_D, node, Stencil, &Stencil);
    } else if (node->RefinmentLevel == node->minNeibRefinmentLevel) {
      if ((0.5 < iLoc) && (iLoc < _BLOCK_CELLS_X_ - 0.5) && (0.5 < jLoc) && (jLoc < _BLOCK_CELLS_Y_ - 0.5) && (0.5 < kLoc) && (kLoc < _BLOCK_CELLS_Z_ - 0.5)) {
        return CplrTrilinearStencil(iLoc, jLoc, kLoc, XyzIn_D, node, Stencil, &Stencil);
      } else {
        return CplrMultiBlockStencil(XyzIn_D, node, Stencil);
      }
    } else {
      cTreeNode *CoarserBlock = NULL;
      cTreeNode *NeibNode;
      int idim, iFace = 0;
      double dxCell[3];
      dxCell[0] = (xmax[0] - xmin[0]) / _BLOCK_CELLS_X_;
      dxCell[1] = (xmax[1] - xmin[1]) / _BLOCK_CELLS_Y_;
      dxCell[2] = (xmax[2] - xmin[2]) / _BLOCK_CELLS_Z_;
      int nBlockCells[3] = {_BLOCK_CELLS_X_, _BLOCK_CELLS_Y_, _BLOCK_CELLS_Z_};
      double dmin = 10.0 * _BLOCK_CELLS_X_ * _BLOCK_CELLS_Y_ * _BLOCK_CELLS_Z_;
      bool CornerTestFlagTable[8] = {false, false, false, false, false, false, false, false};
      bool EdgeTestFlagTable[12] = {false, false, false, false, false, false, false, false, false, false, false, false};
      double xLoc[3] = {iLoc, jLoc, kLoc};
      // distance of the point to the block boundary across direction d: low side -> xLoc, high side -> N - xLoc
      auto usable = [&](cTreeNode *nb) -> bool {
        if (nb == NULL) return false;
        if (!((nb->RefinmentLevel < node->RefinmentLevel) && (nb->IsUsedInCalculationFlag == true))) return false;
        int cnt = 0;
        for (int ii = 0; ii < 3; ii++)
          if ((nb->xmin[ii] - dxCell[ii] <= XyzIn_D[ii]) && (nb->xmax[ii] + dxCell[ii] >= XyzIn_D[ii])) cnt++;
        return cnt == 3;
      };
      for (idim = 0; idim < 3; idim++) {
        if (xLoc[idim] <= 1.0) iFace = 2 * idim;
        else if (xLoc[idim] >= nBlockCells[idim] - 1.0) iFace = 2 * idim + 1;
        else continue;

        NeibNode = GetNeibFace(node, iFace, 0, 0);
        if (usable(NeibNode)) {
          const int d = iFace / 2;
          if ((iFace & 1) == 0) {
            if ((xLoc[d] < 1.0) && (xLoc[d] < dmin)) dmin = xLoc[d], CoarserBlock = NeibNode;
          } else {
            if ((xLoc[d] > nBlockCells[d] - 1) && (nBlockCells[d] - xLoc[d] < dmin)) dmin = nBlockCells[d] - xLoc[d], CoarserBlock = NeibNode;
          }
        }

        static const int faceEdges[6][4] = {{4, 11, 7, 8}, {5, 10, 6, 9}, {0, 9, 3, 8}, {1, 10, 2, 11}, {0, 5, 1, 4}, {3, 6, 2, 7}};
        // edge -> the two transverse directions and their sides (0 low, 1 high), in the order the reference tests them
        static const int edgeDir[12][2] = {{1, 2}, {1, 2}, {1, 2}, {1, 2}, {0, 2}, {0, 2}, {0, 2}, {0, 2}, {0, 1}, {0, 1}, {0, 1}, {0, 1}};
        static const int edgeSide[12][2] = {{0, 0}, {0, 1}, {1, 1}, {1, 0}, {0, 0}, {1, 0}, {1, 1}, {0, 1}, {0, 0}, {1, 0}, {1, 1}, {0, 1}};
        for (int iEdge = 0; iEdge < 4; iEdge++)
          if (EdgeTestFlagTable[faceEdges[iFace][iEdge]] == false) {
            const int e = faceEdges[iFace][iEdge];
            EdgeTestFlagTable[e] = true;
            NeibNode = GetNeibEdge(node, e, 0);
            if (usable(NeibNode)) {
              bool in = true;
              double dist[2];
              for (int q = 0; q < 2; q++) {
                const int d = edgeDir[e][q];
                if (edgeSide[e][q] == 0) {
                  in = in && (xLoc[d] < 1.0);
                  dist[q] = xLoc[d];
                } else {
                  in = in && (xLoc[d] > nBlockCells[d] - 1);
                  dist[q] = nBlockCells[d] - xLoc[d];
                }
              }
              if (in)
                for (int q = 0; q < 2; q++)
                  if (dist[q] < dmin) dmin = dist[q], CoarserBlock = NeibNode;
            }
          }

        static const int FaceNodeMap[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
        for (int iCorner = 0; iCorner < 4; iCorner++)
          if (CornerTestFlagTable[FaceNodeMap[iFace][iCorner]] == false) {
            const int c = FaceNodeMap[iFace][iCorner];
            CornerTestFlagTable[c] = true;
            NeibNode = GetNeibCorner(node, c);
            if (usable(NeibNode)) {
              bool in = true;
              double dist[3];
              for (int d = 0; d < 3; d++) {
                if (((c >> d) & 1) == 0) {
                  in = in && (xLoc[d] < 1.0);
                  dist[d] = xLoc[d];
                } else {
                  in = in && (xLoc[d] > nBlockCells[d] - 1);
                  dist[d] = nBlockCells[d] - xLoc[d];
                }
              }
              if (in)
                for (int d = 0; d < 3; d++)
                  if (dist[d] < dmin) dmin = dist[d], CoarserBlock = NeibNode;
            }
          }
      }

      if (CoarserBlock != NULL) {
        if (!CplrMultiBlockStencil(XyzIn_D, CoarserBlock, Stencil)) return false;
        if ((0.5 < dmin) && (dmin <= 1.0)) {
          cStencil FineStencil;
          // the fine stencil is normalised iff the *outer* stencil's Length != 8 (the reference tests StencilTable, :903)
          if (!CplrTrilinearStencil(iLoc, jLoc, kLoc, XyzIn_D, node, FineStencil, &Stencil)) return false;
          Stencil.MultiplyScalar(1.0 - (dmin - 0.5) / 0.5);
          FineStencil.MultiplyScalar((dmin - 0.5) / 0.5);
          StencilAdd(Stencil, FineStencil);
        }
        return !Stencil.overflow;
      }
      return CplrTrilinearStencil(iLoc, jLoc, kLoc, XyzIn_D, node, Stencil, &Stencil);
    }
  }
  // PIC::CPLR::InitInterpolationStencil, pic_swmf.cpp:76-90
  bool CplrInitStencil(const double *x, cTreeNode *node, cStencil &Stencil) const {
    bool ok = (cfg.coupler_interpolation == AMPS_CPLR_CELL_CENTERED_LINEAR) ? CplrLinearStencil(x, node, Stencil) : CplrConstantStencil(x, node, Stencil);
    return ok && !Stencil.overflow && Stencil.Length > 0;
  }
  void CplrGather(const cStencil &Stencil, int offset, int nVars, double *out) const {
    for (int i = 0; i < nVars; i++) out[i] = 0.0;
    for (int iStencil = 0; iStencil < Stencil.Length; iStencil++) {
      const double *t = Stencil.cell[iStencil]->data + offset;
      for (int i = 0; i < nVars; i++) out[i] += Stencil.Weight[iStencil] * t[i];
    }
  }

  // PIC::Mover::SetBlock_E / SetBlock_B, src/pic/pic_mover.cpp:86-166
  void SetBlock_E(double *E_Corner, cTreeNode *node) const {
    if (!node->block) return;
    for (int k = -_GHOST_CELLS_Z_; k <= _BLOCK_CELLS_Z_ + _GHOST_CELLS_Z_; k++)
      for (int j = -_GHOST_CELLS_Y_; j <= _BLOCK_CELLS_Y_ + _GHOST_CELLS_Y_; j++)
        for (int i = -_GHOST_CELLS_X_; i <= _BLOCK_CELLS_X_ + _GHOST_CELLS_X_; i++) {
          int LocalCornerId = _getCornerNodeLocalNumber(i, j, k);
          if (!node->block->cornerNodes[LocalCornerId]) continue;
          double *ptr = node->block->cornerNodes[LocalCornerId]->data + OffsetE_HalfTimeStep_d;
          memcpy(&E_Corner[LocalCornerId * 3], ptr, 3 * sizeof(double));
        }
  }
  void SetBlock_B(double *B_C, cTreeNode *node) const {
    if (!node->block) return;
    if (cfg.b_mode == AMPS_B_CORNER_BASED) {  // pic_mover.cpp:106-118: block corners only, no ghost layer
      for (int k = 0; k <= _BLOCK_CELLS_Z_; k++)
        for (int j = 0; j <= _BLOCK_CELLS_Y_; j++)
          for (int i = 0; i <= _BLOCK_CELLS_X_; i++) {
            int LocalCornerId = _getCornerNodeLocalNumber(i, j, k);
            if (!node->block->cornerNodes[LocalCornerId]) continue;
            double *ptr = node->block->cornerNodes[LocalCornerId]->data + OffsetB_corner_d + PrevBOffset_d;
            memcpy(&B_C[LocalCornerId * 3], ptr, 3 * sizeof(double));
          }
      return;
    }
    for (int k = -_GHOST_CELLS_Z_; k < _BLOCK_CELLS_Z_ + _GHOST_CELLS_Z_; k++)
      for (int j = -_GHOST_CELLS_Y_; j < _BLOCK_CELLS_Y_ + _GHOST_CELLS_Y_; j++)
        for (int i = -_GHOST_CELLS_X_; i < _BLOCK_CELLS_X_ + _GHOST_CELLS_X_; i++) {
          int LocalCenterId = _getCenterNodeLocalNumber(i, j, k);
          if (!node->block->centerNodes[LocalCenterId]) continue;
          double *ptr = node->block->centerNodes[LocalCenterId]->data + PrevBOffset_d;
          memcpy(&B_C[LocalCenterId * 3], ptr, 3 * sizeof(double));
        }
  }

  // attach to the temp moving list, src/pic/pic_mover_boris.cpp:1333-1365
  void AttachToTempList(long int ptr, byte *ParticleData, cBlock *block, int i, int j, int k, int nThreads, int thread) {
    int cell = i + _BLOCK_CELLS_X_ * (j + _BLOCK_CELLS_Y_ * k);
    if (nThreads <= 1) {  // _COMPILATION_MODE__MPI_
      long int tempFirstCellParticle, *tempFirstCellParticlePtr;
      tempFirstCellParticlePtr = block->tempParticleMovingListTable + cell;
      tempFirstCellParticle = (*tempFirstCellParticlePtr);
      SetNext(tempFirstCellParticle, ParticleData);
      SetPrev(-1, ParticleData);
      if (tempFirstCellParticle != -1) SetPrev(ptr, tempFirstCellParticle);
      *tempFirstCellParticlePtr = ptr;
    } else {  // _COMPILATION_MODE__HYBRID_, per-thread lists
      cTempList *t = block->tempThreadTable + (size_t)thread * nCellsBlock() + cell;
      SetNext(t->first, ParticleData);
      SetPrev(-1, ParticleData);
      if (t->last == -1) t->last = ptr;
      if (t->first != -1) SetPrev(ptr, t->first);
      t->first = ptr;
    }
  }

  // ------------------------------------------------------------------------------------------
  // PIC::Mover::Lapenta2017, src/pic/pic_mover_boris.cpp:876-1393 (scalar, non-AVX branch,
  // _PIC_FIELD_SOLVER_MODE__ELECTROMAGNETIC__ECSIM_, B centre based, no internal sphere)
  // ------------------------------------------------------------------------------------------
  int Lapenta2017(byte *ParticleData, long int ptr, cTreeNode *startNode, const double *E_Corner, const double *B_C, int nThreads, int thread,
                  cTreeNode **newNodeOut) {
    cTreeNode *newNode = NULL;
    int idim, i, j, k, spec;
    double dtTotal;
    double vInit[3], xInit[3], vFinal[3], xFinal[3];

    GetV(vInit, ParticleData);
    GetX(xInit, ParticleData);
    spec = GetI(ParticleData);

    switch (cfg.time_step_mode) {
      case AMPS_DT_SPECIES_GLOBAL: dtTotal = cfg.time_step[spec]; break;
      default: dtTotal = cfg.time_step[0];
    }

    cStencil ElectricFieldStencil, MagneticFieldStencil;
    double E[4] = {0.0, 0.0, 0.0, 0.0}, B[4] = {0.0, 0.0, 0.0, 0.0};
    int *LocalCellID, Length;
    double *Weight;
    double Wtab[8];

    if (!CornerBased_InitStencil(xInit, startNode, ElectricFieldStencil, Wtab)) return _ORACLE_ERROR_;
    Length = ElectricFieldStencil.Length;
    LocalCellID = ElectricFieldStencil.LocalCellID;
    Weight = ElectricFieldStencil.Weight;

    for (int iStencil = 0; iStencil < Length; iStencil++) {
      const double *tempE1 = E_Corner + 3 * LocalCellID[iStencil];
      const double *tempB1 = (cfg.b_mode == AMPS_B_CORNER_BASED) ? B_C + 3 * LocalCellID[iStencil] : NULL;
      for (idim = 0; idim < 3; idim++) {
        E[idim] += Weight[iStencil] * tempE1[idim];
        if (cfg.b_mode == AMPS_B_CORNER_BASED) B[idim] += Weight[iStencil] * tempB1[idim];
      }
    }

    if (cfg.b_mode == AMPS_B_CENTER_BASED) {
      CellCentered_Linear_InitStencil(xInit, startNode, MagneticFieldStencil, globalStencilLength != 8);
      Length = MagneticFieldStencil.Length;
      LocalCellID = MagneticFieldStencil.LocalCellID;
      Weight = MagneticFieldStencil.Weight;
      for (int iStencil = 0; iStencil < Length; iStencil++) {
        const double *tempB1 = B_C + 3 * LocalCellID[iStencil];
        for (idim = 0; idim < 3; idim++) B[idim] += Weight[iStencil] * tempB1[idim];
      }
    }

    double QdT_over_m, QdT_over_2m, alpha[3][3];
    double c0, QdT_over_2m_squared, mass, chargeQ;

    chargeQ = cfg.charge[spec];  // picunits::si2no_q applied by the host once per species
    mass = cfg.mass[spec];

    QdT_over_m = chargeQ * dtTotal / mass;
    QdT_over_2m = 0.5 * QdT_over_m;
    QdT_over_2m_squared = QdT_over_2m * QdT_over_2m;

    double BB[3][3], P[3];
    for (i = 0; i < 3; i++) {
      P[i] = -QdT_over_2m * B[i];
      for (j = 0; j <= i; j++) {
        BB[i][j] = QdT_over_2m_squared * B[i] * B[j];
        BB[j][i] = BB[i][j];
      }
    }
    c0 = 1.0 / (1.0 + QdT_over_2m_squared * (B[0] * B[0] + B[1] * B[1] + B[2] * B[2]));

    alpha[0][0] = c0 * (1.0 + BB[0][0]);
    alpha[0][1] = c0 * (-P[2] + BB[0][1]);
    alpha[0][2] = c0 * (P[1] + BB[0][2]);
    alpha[1][0] = c0 * (P[2] + BB[1][0]);
    alpha[1][1] = c0 * (1.0 + BB[1][1]);
    alpha[1][2] = c0 * (-P[0] + BB[1][2]);
    alpha[2][0] = c0 * (-P[1] + BB[2][0]);
    alpha[2][1] = c0 * (P[0] + BB[2][1]);
    alpha[2][2] = c0 * (1.0 + BB[2][2]);

    for (idim = 0; idim < 3; idim++) {
      double vp = 0.0;
      for (j = 0; j < 3; j++) vp += alpha[idim][j] * (vInit[j] + QdT_over_2m * E[j]);
      vFinal[idim] = 2.0 * vp - vInit[idim];
    }
    for (idim = 0; idim < 3; idim++) xFinal[idim] = xInit[idim] + dtTotal * vFinal[idim];

    newNode = findTreeNode(xFinal, startNode);

    if (newNode == NULL) {
      // the particle left the computational domain, pic_mover_boris.cpp:1159-1265
      int code = 0;  // _PARTICLE_DELETED_ON_THE_FACE_
      if (cfg.boundary_mode != AMPS_BOUNDARY_DELETE) {
        int nface, nIntersectionFace = -1;
        double tVelocityIncrement, cx, cv, r0[3], dt, vMiddle[3] = {0.5 * (vInit[0] + vFinal[0]), 0.5 * (vInit[1] + vFinal[1]), 0.5 * (vInit[2] + vFinal[2])}, c, dtIntersection = -1.0;
        for (nface = 0; nface < 6; nface++) {
          for (idim = 0, cx = 0.0, cv = 0.0; idim < 3; idim++) {
            r0[idim] = xInit[idim] - FaceTable[nface].x0[idim];
            cx += r0[idim] * FaceTable[nface].norm[idim];
            cv += vMiddle[idim] * FaceTable[nface].norm[idim];
          }
          if (cv > 0.0) {
            dt = -cx / cv;
            if ((dtIntersection < 0.0) || ((dt < dtIntersection) && (dt > 0.0))) {
              double cE0 = 0.0, cE1 = 0.0;
              for (idim = 0; idim < 3; idim++) {
                c = r0[idim] + dt * vMiddle[idim];
                cE0 += c * FaceTable[nface].e0[idim], cE1 += c * FaceTable[nface].e1[idim];
              }
              if ((cE0 < -EPS) || (cE0 > FaceTable[nface].lE0 + EPS) || (cE1 < -EPS) || (cE1 > FaceTable[nface].lE1 + EPS)) continue;
              nIntersectionFace = nface, dtIntersection = dt;
            }
          }
        }
        if (nIntersectionFace == -1) return _ORACLE_ERROR_;
        for (idim = 0, tVelocityIncrement = ((dtIntersection / dtTotal < 1) ? dtIntersection / dtTotal : 1); idim < 3; idim++) {
          xInit[idim] += dtIntersection * vMiddle[idim] - FaceTable[nIntersectionFace].norm[idim] * EPS;
          vInit[idim] += tVelocityIncrement * (vFinal[idim] - vInit[idim]);
        }
        newNode = findTreeNode(xInit, startNode);
        if (newNode == NULL) {
          double xmin[3], xmax[3];
          memcpy(xmin, xGlobalMin, 3 * sizeof(double));
          memcpy(xmax, xGlobalMax, 3 * sizeof(double));
          for (int ii = 0; ii < 3; ii++) {
            if (xmin[ii] >= xInit[ii]) xInit[ii] = xmin[ii] + EPS;
            if (xmax[ii] <= xInit[ii]) xInit[ii] = xmax[ii] - EPS;
          }
          newNode = findTreeNode(xInit, startNode);
          if (newNode == NULL) return _ORACLE_ERROR_;
        }
        switch (cfg.boundary_mode) {
          case AMPS_BOUNDARY_SPECULAR_REFLECTION: {
            double cc = 0.0;
            for (int d = 0; d < 3; d++) cc += FaceTable[nIntersectionFace].norm[d] * vInit[d];
            for (int d = 0; d < 3; d++) vInit[d] -= 2.0 * cc * FaceTable[nIntersectionFace].norm[d];
            code = 1;  // _PARTICLE_REJECTED_ON_THE_FACE_
          } break;
          default:
            // user function: recorded for the host (CutoffRigidity::ProcessOutsideDomainParticles always deletes)
            AddExitRecord(ptr, spec, nIntersectionFace, newNode, xInit, vInit);
            code = 0;
        }
        memcpy(vFinal, vInit, 3 * sizeof(double));
        memcpy(xFinal, xInit, 3 * sizeof(double));
      }
      switch (code) {
        case 0:
          DeleteParticle(ptr);
          return _PARTICLE_LEFT_THE_DOMAIN_;
        default:
          // the reference exit()s here ("not implemented", pic_mover_boris.cpp:1262-1263)
          return _ORACLE_ERROR_;
      }
    } else {
      if (newNode->IsUsedInCalculationFlag == false) {
        DeleteParticle(ptr);
        return _PARTICLE_IN_NOT_IN_USE_NODE_;
      }
    }

    cBlock *block;
    if (FindCellIndex(xFinal, i, j, k, newNode) == -1) return _ORACLE_ERROR_;
    if ((block = newNode->block) == NULL) {
      DeleteParticle(ptr);
      return _PARTICLE_LEFT_THE_DOMAIN_;
    }

    AttachToTempList(ptr, ParticleData, block, i, j, k, nThreads, thread);
    SetV(vFinal, ParticleData);
    SetX(xFinal, ParticleData);
    *newNodeOut = newNode;
    return _PARTICLE_MOTION_FINISHED_;
  }


  // exit records handed to fProcessOutsideDomainParticles / ParticleSphereInteraction
  std::vector<amps_gpu_exit_record> exitRecords;
  void AddExitRecord(long int ptr, int spec, int face, cTreeNode *node, const double *x, const double *v) {
#pragma omp critical(oracle_exit)
    {
      amps_gpu_exit_record r;
      r.ptr = (int)ptr, r.species = spec, r.face = face, r.leaf = node ? node->leaf : -1;
      for (int d = 0; d < 3; d++) r.x[d] = x[d], r.v[d] = v[d];
      exitRecords.push_back(r);
    }
  }

  // PIC::CPLR::InitInterpolationStencil (pic_swmf.cpp:76-90) + GetBackgroundElectricField/MagneticField
  // (pic.h:8338-8425) on the coupler's centre-node table.  Here the stencil IS the global StencilTable, so the
  // "Length != 8 -> Normalize" test of GetTriliniarInterpolationStencil (:903) applies as written.
  bool GetBackgroundFields(const double *x, cTreeNode *node, double *E, double *B) const {
    cStencil Stencil;
    if (!CplrInitStencil(x, node, Stencil)) return false;
    CplrGather(Stencil, BackgroundE_d, 3, E);
    CplrGather(Stencil, BackgroundB_d, 3, B);
    return true;
  }

  // domain-exit search shared by the test-particle movers, pic_mover_relativistic_boris.cpp:320-446.
  // NOTE (reference defects, restated as intended): nIntersectionFace is read uninitialised when no face is found
  // (:381 tests it against -1) -> initialised to -1 here; the position shift at :387/:393 indexes the face table with
  // the loop variable `nface` (== 6 after the loop, out of bounds) -> the intersection face is used.
  // returns the reference's `code`: 0 = _PARTICLE_DELETED_ON_THE_FACE_, -1 = would exit()
  int ProcessDomainExit(long int ptr, int spec, double *xInit, double *vInit, double *xFinal, double *vFinal, cTreeNode *startNode, cTreeNode **newNodeOut) {
    int idim, nface, nIntersectionFace = -1;
    double cx, cv, r0[3], dtEffective, vEffective[3], c, dtIntersection = -1.0;
    const bool backward = cfg.backward_time_integration != 0;
    if (backward) for (idim = 0; idim < 3; idim++) vEffective[idim] = xInit[idim] - xFinal[idim];
    else for (idim = 0; idim < 3; idim++) vEffective[idim] = xFinal[idim] - xInit[idim];
    for (nface = 0; nface < 6; nface++) {
      if (backward) {
        for (idim = 0, cx = 0.0, cv = 0.0; idim < 3; idim++) {
          r0[idim] = xFinal[idim] - FaceTable[nface].x0[idim];
          cx += r0[idim] * FaceTable[nface].norm[idim];
          cv += vEffective[idim] * FaceTable[nface].norm[idim];
        }
        dtEffective = (cv < 0.0) ? -cx / cv : -1.0;
      } else {
        for (idim = 0, cx = 0.0, cv = 0.0; idim < 3; idim++) {
          r0[idim] = xInit[idim] - FaceTable[nface].x0[idim];
          cx += r0[idim] * FaceTable[nface].norm[idim];
          cv += vEffective[idim] * FaceTable[nface].norm[idim];
        }
        dtEffective = (cv > 0.0) ? -cx / cv : -1.0;
      }
      if (dtEffective > 0.0) {
        if ((dtIntersection < 0.0) || ((dtEffective < dtIntersection) && (dtEffective > 0.0))) {
          double cE0 = 0.0, cE1 = 0.0;
          for (idim = 0; idim < 3; idim++) {
            c = r0[idim] + dtEffective * vEffective[idim];
            cE0 += c * FaceTable[nface].e0[idim], cE1 += c * FaceTable[nface].e1[idim];
          }
          if ((cE0 < -EPS) || (cE0 > FaceTable[nface].lE0 + EPS) || (cE1 < -EPS) || (cE1 > FaceTable[nface].lE1 + EPS)) continue;
          nIntersectionFace = nface, dtIntersection = dtEffective;
        }
      }
    }
    if (nIntersectionFace == -1) return -1;
    if (backward) {
      for (idim = 0; idim < 3; idim++) {
        xInit[idim] = xFinal[idim] + dtIntersection * (xInit[idim] - xFinal[idim]) - FaceTable[nIntersectionFace].norm[idim] * EPS;
        vInit[idim] = vFinal[idim] + dtIntersection * (vInit[idim] - vFinal[idim]);
      }
    } else {
      for (idim = 0; idim < 3; idim++) {
        xInit[idim] += dtIntersection * (xFinal[idim] - xInit[idim]) - FaceTable[nIntersectionFace].norm[idim] * EPS;
        vInit[idim] += dtIntersection * (vFinal[idim] - vInit[idim]);
      }
    }
    cTreeNode *newNode = findTreeNode(xInit, startNode);
    if (newNode == NULL) {
      for (int ii = 0; ii < 3; ii++) {
        if (xGlobalMin[ii] >= xInit[ii]) xInit[ii] = xGlobalMin[ii] + EPS;
        if (xGlobalMax[ii] <= xInit[ii]) xInit[ii] = xGlobalMax[ii] - EPS;
      }
      newNode = findTreeNode(xInit, startNode);
      if (newNode == NULL) return -1;
    }
    int code;
    switch (cfg.boundary_mode) {
      case AMPS_BOUNDARY_USER_FUNCTION:
        // ProcessOutsideDomainParticles(ptr,xInit,vInit,nIntersectionFace,newNode): recorded for the host; the
        // cutoff-rigidity callback always deletes (srcEarth/CutoffRigidity.cpp:129-230)
        AddExitRecord(ptr, spec, nIntersectionFace, newNode, xInit, vInit);
        code = 0;
        break;
      case AMPS_BOUNDARY_SPECULAR_REFLECTION: {
        double cc = 0.0;
        for (int d = 0; d < 3; d++) cc += FaceTable[nIntersectionFace].norm[d] * vInit[d];
        for (int d = 0; d < 3; d++) vInit[d] -= 2.0 * cc * FaceTable[nIntersectionFace].norm[d];
        code = 1;  // _PARTICLE_REJECTED_ON_THE_FACE_ -> the reference exit()s ("not implemented", :452-453)
      } break;
      default:
        return -1;
    }
    memcpy(vFinal, vInit, 3 * sizeof(double));
    memcpy(xFinal, xInit, 3 * sizeof(double));
    *newNodeOut = newNode;
    return code;
  }

  // ------------------------------------------------------------------------------------------
  // PIC::Mover::Relativistic::Boris, src/pic/pic_mover_relativistic_boris.cpp:16-583 (scalar branch)
  // ------------------------------------------------------------------------------------------
  int Relativistic_Boris(byte *ParticleData, long int ptr, double dtTotalIn, cTreeNode *startNode, int nThreads, int thread, cTreeNode **newNodeOut) {
    cTreeNode *newNode = NULL;
    double gamma;
    double mass, QdT_over_twoM, ElectricCharge;
    int idim, i, j, k, spec;
    double uMinus[3], E[3], B[3];
    double vInit[3], xInit[3], xFinal[3], vFinal[4];
    const double SpeedOfLight = cfg.speed_of_light;
    const bool backward = cfg.backward_time_integration != 0;

    GetV(vInit, ParticleData);
    GetX(xInit, ParticleData);
    spec = GetI(ParticleData);
    ElectricCharge = cfg.charge[spec];
    mass = cfg.mass[spec];

    if (dtTotalIn == 0.0) {
      memcpy(xFinal, xInit, 3 * sizeof(double));
      memcpy(vFinal, vInit, 3 * sizeof(double));
      newNode = startNode;
    } else
      while (dtTotalIn > 0.0) {
#pragma omp atomic
        nSubSteps++;
        gamma = 1.0 / sqrt(1.0 - (vInit[0] * vInit[0] + vInit[1] * vInit[1] + vInit[2] * vInit[2]) / (SpeedOfLight * SpeedOfLight));
        if (!GetBackgroundFields(xInit, startNode, E, B)) return _ORACLE_ERROR_;

        double dt, GyroFreq, dtMax;
        if (sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]) > 1.0E-25) {
          GyroFreq = RelGyroFrequency(vInit, mass, ElectricCharge, B, SpeedOfLight);
          dtMax = 1.0 / GyroFreq;
          dt = (dtMax < dtTotalIn) ? dtMax : dtTotalIn;
        } else
          dt = dtTotalIn;
        dtTotalIn -= dt;

        if (backward)
          for (idim = 0; idim < 3; idim++) vInit[idim] = -vInit[idim], B[idim] = -B[idim];

        QdT_over_twoM = ElectricCharge * dt / (2.0 * mass);
        for (idim = 0; idim < 3; idim++) uMinus[idim] = gamma * vInit[idim] + QdT_over_twoM * E[idim];

        double t[3], s[3], uPrime[3], uPlus[3], l = 0.0;
        gamma = sqrt(1.0 + (uMinus[0] * uMinus[0] + uMinus[1] * uMinus[1] + uMinus[2] * uMinus[2]) / (SpeedOfLight * SpeedOfLight));
        for (idim = 0; idim < 3; idim++) {
          t[idim] = QdT_over_twoM / gamma * B[idim];
          l += t[idim] * t[idim];  // pow(t,2)
        }
        // Vector3D::CrossProduct(uPrime,uMinus,t)
        uPrime[0] = uMinus[1] * t[2] - uMinus[2] * t[1];
        uPrime[1] = uMinus[2] * t[0] - uMinus[0] * t[2];
        uPrime[2] = uMinus[0] * t[1] - uMinus[1] * t[0];
        for (idim = 0; idim < 3; idim++) uPrime[idim] += uMinus[idim];
        for (idim = 0; idim < 3; idim++) s[idim] = 2.0 * t[idim] / (1.0 + l);
        uPlus[0] = uPrime[1] * s[2] - uPrime[2] * s[1];
        uPlus[1] = uPrime[2] * s[0] - uPrime[0] * s[2];
        uPlus[2] = uPrime[0] * s[1] - uPrime[1] * s[0];
        for (idim = 0; idim < 3; idim++) uPlus[idim] += uMinus[idim];

        double uFinal[3];
        for (idim = 0; idim < 3; idim++) uFinal[idim] = uPlus[idim] + QdT_over_twoM * E[idim];
        gamma = sqrt(1.0 + (uFinal[0] * uFinal[0] + uFinal[1] * uFinal[1] + uFinal[2] * uFinal[2]) / (SpeedOfLight * SpeedOfLight));
        for (idim = 0; idim < 3; idim++) {
          vFinal[idim] = uFinal[idim] / gamma;
          xFinal[idim] = xInit[idim] + vFinal[idim] * dt;
        }
        if (backward)
          for (idim = 0; idim < 3; idim++) vFinal[idim] = -vFinal[idim], vInit[idim] = -vInit[idim];

        // internal sphere (:270-302)
        if (cfg.internal_sphere_radius > 0.0) {
          const double rSphere = cfg.internal_sphere_radius;
          double rFinal2;
          if ((rFinal2 = xFinal[0] * xFinal[0] + xFinal[1] * xFinal[1] + xFinal[2] * xFinal[2]) < rSphere * rSphere) {
            double r = sqrt(rFinal2);
            for (idim = 0; idim < 3; idim++) xFinal[idim] *= rSphere / r;
            newNode = findTreeNode(xFinal, startNode);
            // ParticleSphereInteraction(spec,ptr,xFinal,vFinal,dt,newNode,...): recorded; the cutoff-rigidity model deletes
            AddExitRecord(ptr, spec, AMPS_EXIT_SPHERE, newNode, xFinal, vFinal);
            DeleteParticle(ptr);
            return _PARTICLE_LEFT_THE_DOMAIN_;
          } else
            newNode = findTreeNode(xFinal, startNode);
        } else
          newNode = findTreeNode(xFinal, startNode);

        if (newNode == NULL) {
          int code = 0;
          if (cfg.boundary_mode != AMPS_BOUNDARY_DELETE) {
            code = ProcessDomainExit(ptr, spec, xInit, vInit, xFinal, vFinal, startNode, &newNode);
          }
          switch (code) {
            case 0:
              DeleteParticle(ptr);
              return _PARTICLE_LEFT_THE_DOMAIN_;
            default:
              return _ORACLE_ERROR_;
          }
        } else {
          if (newNode->IsUsedInCalculationFlag == false) {
            DeleteParticle(ptr);
            return _PARTICLE_IN_NOT_IN_USE_NODE_;
          }
        }
        if (newNode->block == NULL) return _ORACLE_ERROR_;  // fields of a block that is not allocated here
        startNode = newNode;
        memcpy(xInit, xFinal, 3 * sizeof(double));
        memcpy(vInit, vFinal, 3 * sizeof(double));
      }

    cBlock *block;
    if (FindCellIndex(xFinal, i, j, k, newNode) == -1) return _ORACLE_ERROR_;
    if ((block = newNode->block) == NULL) return _ORACLE_ERROR_;
    AttachToTempList(ptr, ParticleData, block, i, j, k, nThreads, thread);
    SetV(vFinal, ParticleData);
    SetX(xFinal, ParticleData);
    *newNodeOut = newNode;
    return _PARTICLE_MOTION_FINISHED_;
  }


  // domain exit of the single-step movers (Boris :362-462, Lapenta2017 :1159-1265): mid-velocity ray against the six
  // faces.  returns code 0 = _PARTICLE_DELETED_ON_THE_FACE_, 1 = _PARTICLE_REJECTED_ON_THE_FACE_, -1 = exit()
  int ProcessDomainExit_vMiddle(long int ptr, int spec, double dtTotal, double *xInit, double *vInit, double *xFinal, double *vFinal, cTreeNode *startNode,
                                cTreeNode **newNodeOut) {
    int idim, nface, nIntersectionFace = -1;
    double tVelocityIncrement, cx, cv, r0[3], dt, vMiddle[3] = {0.5 * (vInit[0] + vFinal[0]), 0.5 * (vInit[1] + vFinal[1]), 0.5 * (vInit[2] + vFinal[2])}, c,
                                                  dtIntersection = -1.0;
    for (nface = 0; nface < 6; nface++) {
      for (idim = 0, cx = 0.0, cv = 0.0; idim < 3; idim++) {
        r0[idim] = xInit[idim] - FaceTable[nface].x0[idim];
        cx += r0[idim] * FaceTable[nface].norm[idim];
        cv += vMiddle[idim] * FaceTable[nface].norm[idim];
      }
      if (cv > 0.0) {
        dt = -cx / cv;
        if ((dtIntersection < 0.0) || ((dt < dtIntersection) && (dt > 0.0))) {
          double cE0 = 0.0, cE1 = 0.0;
          for (idim = 0; idim < 3; idim++) {
            c = r0[idim] + dt * vMiddle[idim];
            cE0 += c * FaceTable[nface].e0[idim], cE1 += c * FaceTable[nface].e1[idim];
          }
          if ((cE0 < -EPS) || (cE0 > FaceTable[nface].lE0 + EPS) || (cE1 < -EPS) || (cE1 > FaceTable[nface].lE1 + EPS)) continue;
          nIntersectionFace = nface, dtIntersection = dt;
        }
      }
    }
    if (nIntersectionFace == -1) return -1;
    for (idim = 0, tVelocityIncrement = ((dtIntersection / dtTotal < 1) ? dtIntersection / dtTotal : 1); idim < 3; idim++) {
      xInit[idim] += dtIntersection * vMiddle[idim] - FaceTable[nIntersectionFace].norm[idim] * EPS;
      vInit[idim] += tVelocityIncrement * (vFinal[idim] - vInit[idim]);
    }
    cTreeNode *newNode = findTreeNode(xInit, startNode);
    if (newNode == NULL) {
      for (int ii = 0; ii < 3; ii++) {
        if (xGlobalMin[ii] >= xInit[ii]) xInit[ii] = xGlobalMin[ii] + EPS;
        if (xGlobalMax[ii] <= xInit[ii]) xInit[ii] = xGlobalMax[ii] - EPS;
      }
      newNode = findTreeNode(xInit, startNode);
      if (newNode == NULL) return -1;
    }
    int code;
    switch (cfg.boundary_mode) {
      case AMPS_BOUNDARY_USER_FUNCTION:
        AddExitRecord(ptr, spec, nIntersectionFace, newNode, xInit, vInit);
        code = 0;
        break;
      case AMPS_BOUNDARY_SPECULAR_REFLECTION: {
        double cc = 0.0;
        for (int d = 0; d < 3; d++) cc += FaceTable[nIntersectionFace].norm[d] * vInit[d];
        for (int d = 0; d < 3; d++) vInit[d] -= 2.0 * cc * FaceTable[nIntersectionFace].norm[d];
        code = 1;
      } break;
      default:
        return -1;
    }
    memcpy(vFinal, vInit, 3 * sizeof(double));
    memcpy(xFinal, xInit, 3 * sizeof(double));
    *newNodeOut = newNode;
    return code;
  }

  // ------------------------------------------------------------------------------------------
  // PIC::Mover::Boris + BorisSplitAcceleration_default, src/pic/pic_mover_boris.cpp:126-553, :22-123
  // (scalar branch, planar symmetry, Lorentz force from the coupler table [+ central gravity])
  // ------------------------------------------------------------------------------------------
  int Boris(byte *ParticleData, long int ptr, double dtTotal, cTreeNode *startNode, int nThreads, int thread, cTreeNode **newNodeOut) {
    cTreeNode *newNode = NULL;
    int idim, i, j, k, spec;
    double vInit[3], xInit[3] = {0.0, 0.0, 0.0}, vFinal[3], xFinal[3];
    double u[3] = {0.0, 0.0, 0.0}, U[3] = {0.0, 0.0, 0.0};
    double acclInit[3], rotInit[3];
    GetV(vInit, ParticleData);
    GetX(xInit, ParticleData);
    spec = GetI(ParticleData);

    {  // BorisSplitAcceleration_default
      double accl_LOCAL[3] = {0.0, 0.0, 0.0}, rotation_LOCAL[3] = {0.0, 0.0, 0.0};
      double E[3] = {0.0, 0.0, 0.0}, B[3] = {0.0, 0.0, 0.0};
      cTreeNode *fieldNode = startNode;
      if (fieldNode->block == NULL) return _ORACLE_ERROR_;
      long int nd = FindCellIndex(xInit, i, j, k, fieldNode);
      if (nd == -1) {
        fieldNode = findTreeNode(xInit, fieldNode);
        if (fieldNode == NULL || fieldNode->block == NULL) return _ORACLE_ERROR_;
        nd = FindCellIndex(xInit, i, j, k, fieldNode);
        if (nd == -1) return _ORACLE_ERROR_;
      }
      if (!GetBackgroundFields(xInit, fieldNode, E, B)) return _ORACLE_ERROR_;
      double ElectricCharge = cfg.charge[spec], mass = cfg.mass[spec], Charge2Mass;
      if (ElectricCharge != 0.0) {
        Charge2Mass = ElectricCharge / mass;
        for (idim = 0; idim < 3; idim++) {
          accl_LOCAL[idim] += Charge2Mass * E[idim];
          rotation_LOCAL[idim] -= Charge2Mass * B[idim];
        }
      }
      if (cfg.gravity_gm != 0.0) {  // GravityConstant*_MASS_(_TARGET_)
        double r2 = xInit[0] * xInit[0] + xInit[1] * xInit[1] + xInit[2] * xInit[2];
        double r = sqrt(r2);
        for (idim = 0; idim < 3; idim++) accl_LOCAL[idim] -= cfg.gravity_gm / r2 * xInit[idim] / r;
      }
      memcpy(acclInit, accl_LOCAL, 3 * sizeof(double));
      memcpy(rotInit, rotation_LOCAL, 3 * sizeof(double));
    }

    double dtTempOverTwo, dtTemp;
    if (cfg.backward_time_integration) dtTemp = -dtTotal, dtTempOverTwo = -dtTotal / 2.0;
    else dtTemp = dtTotal, dtTempOverTwo = dtTotal / 2.0;

    u[0] = vInit[0] + dtTempOverTwo * acclInit[0];
    u[1] = vInit[1] + dtTempOverTwo * acclInit[1];
    u[2] = vInit[2] + dtTempOverTwo * acclInit[2];
    double h[3];
    h[0] = -dtTempOverTwo * rotInit[0];
    h[1] = -dtTempOverTwo * rotInit[1];
    h[2] = -dtTempOverTwo * rotInit[2];
    double h2 = h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
    double uh = u[0] * h[0] + u[1] * h[1] + u[2] * h[2];
    U[0] = ((1 - h2) * u[0] + 2 * (u[1] * h[2] - h[1] * u[2] + uh * h[0])) / (1 + h2);
    U[1] = ((1 - h2) * u[1] + 2 * (u[2] * h[0] - h[2] * u[0] + uh * h[1])) / (1 + h2);
    U[2] = ((1 - h2) * u[2] + 2 * (u[0] * h[1] - h[0] * u[1] + uh * h[2])) / (1 + h2);
    vFinal[0] = U[0] + dtTempOverTwo * acclInit[0];
    vFinal[1] = U[1] + dtTempOverTwo * acclInit[1];
    vFinal[2] = U[2] + dtTempOverTwo * acclInit[2];
    xFinal[0] = xInit[0] + dtTemp * vFinal[0];
    xFinal[1] = xInit[1] + dtTemp * vFinal[1];
    xFinal[2] = xInit[2] + dtTemp * vFinal[2];

    // internal sphere centred at the origin (:296-358)
    if (cfg.internal_sphere_radius > 0.0) {
      const double R = cfg.internal_sphere_radius;
      const double dx0 = xFinal[0] - 0.0, dx1 = xFinal[1] - 0.0, dx2 = xFinal[2] - 0.0;
      const double r2 = dx0 * dx0 + dx1 * dx1 + dx2 * dx2;
      if (r2 < R * R) {
        double r = sqrt(r2);
        if (r <= 0.0) r = 1.0;
        xFinal[0] = 0.0 + dx0 * (R / r);
        xFinal[1] = 0.0 + dx1 * (R / r);
        xFinal[2] = 0.0 + dx2 * (R / r);
        newNode = findTreeNode(xFinal, startNode);
        AddExitRecord(ptr, spec, AMPS_EXIT_SPHERE, newNode, xFinal, vFinal);
        DeleteParticle(ptr);
        return _PARTICLE_LEFT_THE_DOMAIN_;
      } else
        newNode = findTreeNode(xFinal, startNode);
    } else
      newNode = findTreeNode(xFinal, startNode);

    if (newNode == NULL) {
      int code = 0;
      if (cfg.boundary_mode != AMPS_BOUNDARY_DELETE) code = ProcessDomainExit_vMiddle(ptr, spec, dtTotal, xInit, vInit, xFinal, vFinal, startNode, &newNode);
      switch (code) {
        case 0:
          DeleteParticle(ptr);
          return _PARTICLE_LEFT_THE_DOMAIN_;
        default:
          return _ORACLE_ERROR_;  // exit("not implemented") :461
      }
    } else if (newNode->IsUsedInCalculationFlag == false) {
      DeleteParticle(ptr);
      return _PARTICLE_IN_NOT_IN_USE_NODE_;
    }
    cBlock *block;
    if (FindCellIndex(xFinal, i, j, k, newNode) == -1) return _ORACLE_ERROR_;
    if ((block = newNode->block) == NULL) {
      DeleteParticle(ptr);
      return _PARTICLE_LEFT_THE_DOMAIN_;
    }
    AttachToTempList(ptr, ParticleData, block, i, j, k, nThreads, thread);
    SetV(vFinal, ParticleData);
    SetX(xFinal, ParticleData);
    *newNodeOut = newNode;
    return _PARTICLE_MOTION_FINISHED_;
  }


  // PIC::Mover::Markidis2010, src/pic/pic_mover_boris.cpp:557-835 (Markidis et al. 2010, eqs 22-23; fields from the coupler).
  // The exit search is the vMiddle one shared with Boris (nIntersectionFace is read uninitialised at :722 when no face is
  // found: restated as intended, see ProcessDomainExit_vMiddle).
  int Markidis2010(byte *ParticleData, long int ptr, double dtTotal, cTreeNode *startNode, int nThreads, int thread, cTreeNode **newNodeOut) {
    cTreeNode *newNode = NULL;
    double vInit[3], xInit[3] = {0.0, 0.0, 0.0}, vFinal[3], xFinal[3], B[3], E[3];
    int idim, i, j, k, spec;
    GetV(vInit, ParticleData);
    GetX(xInit, ParticleData);
    spec = GetI(ParticleData);
    if (!GetBackgroundFields(xInit, startNode, E, B)) return _ORACLE_ERROR_;
    double v_prime[3], QdT_over_m, QdT_over_2m;
    QdT_over_m = cfg.charge[spec] * dtTotal / cfg.mass[spec];
    QdT_over_2m = 0.5 * QdT_over_m;
    for (idim = 0; idim < 3; idim++) v_prime[idim] = vInit[idim] + QdT_over_m * E[idim];
    double Denominator = 1.0 / (1.0 + QdT_over_2m * QdT_over_2m * (B[0] * B[0] + B[1] * B[1] + B[2] * B[2]));
    double n1[3], n2;
    n1[0] = v_prime[1] * B[2] - v_prime[2] * B[1];  // Vector3D::CrossProduct, specfunc.h:780-806
    n1[1] = v_prime[2] * B[0] - v_prime[0] * B[2];
    n1[2] = v_prime[0] * B[1] - v_prime[1] * B[0];
    n2 = QdT_over_2m * QdT_over_2m * (v_prime[0] * B[0] + v_prime[1] * B[1] + v_prime[2] * B[2]);
    for (idim = 0; idim < 3; idim++) {
      vFinal[idim] = Denominator * (v_prime[idim] + QdT_over_2m * n1[idim] + n2 * B[idim]);
      xFinal[idim] = xInit[idim] + dtTotal * vFinal[idim];
    }
    if (cfg.internal_sphere_radius > 0.0) {
      const double R = cfg.internal_sphere_radius;
      double rFinal2;
      if ((rFinal2 = xFinal[0] * xFinal[0] + xFinal[1] * xFinal[1] + xFinal[2] * xFinal[2]) < R * R) {
        double r = sqrt(rFinal2);
        for (idim = 0; idim < 3; idim++) xFinal[idim] *= R / r;
        newNode = findTreeNode(xFinal, startNode);
        AddExitRecord(ptr, spec, AMPS_EXIT_SPHERE, newNode, xFinal, vFinal);  // ParticleSphereInteraction -> deleted
        DeleteParticle(ptr);
        return _PARTICLE_LEFT_THE_DOMAIN_;
      } else
        newNode = findTreeNode(xFinal, startNode);
    } else
      newNode = findTreeNode(xFinal, startNode);
    if (newNode == NULL) {
      int code = 0;
      if (cfg.boundary_mode != AMPS_BOUNDARY_DELETE) code = ProcessDomainExit_vMiddle(ptr, spec, dtTotal, xInit, vInit, xFinal, vFinal, startNode, &newNode);
      if (code != 0) return _ORACLE_ERROR_;  // exit("not implemented") :787
      DeleteParticle(ptr);
      return _PARTICLE_LEFT_THE_DOMAIN_;
    }
    cBlock *block;
    if (FindCellIndex(xFinal, i, j, k, newNode) == -1) return _ORACLE_ERROR_;
    if ((block = newNode->block) == NULL) return _ORACLE_ERROR_;
    AttachToTempList(ptr, ParticleData, block, i, j, k, nThreads, thread);
    SetV(vFinal, ParticleData);
    SetX(xFinal, ParticleData);
    *newNodeOut = newNode;
    return _PARTICLE_MOTION_FINISHED_;
  }

  // fields + the 15 GCA variables through the coupler stencil (pic.h:8338-8425, 8643-8680)
  bool GetBackgroundFieldsGCA(const double *x, cTreeNode *node, double *E, double *B, double *v15) const {
    cStencil Stencil;
    if (!CplrInitStencil(x, node, Stencil)) return false;
    CplrGather(Stencil, BackgroundB_d, 3, B);
    CplrGather(Stencil, BackgroundE_d, 3, E);
    if (v15) CplrGather(Stencil, BackgroundGCA_d, 15, v15);
    return true;
  }

  // PIC::Mover::Relativistic::GuidingCenter::InitiateMagneticMoment, pic_mover_relativistic_guiding_center.cpp:19-93
  // NOTE (reference defect, restated as intended): :45 reads |vE| before vE is computed (uninitialised stack); here
  // vE = E x B / B^2 is formed first and vE_norm is its length.
  bool RelGCA_InitiateMagneticMoment(int spec, const double *x, const double *v, byte *ParticleData, cTreeNode *node) {
    double B[3] = {0.0, 0.0, 0.0}, AbsB = 0.0, E[3] = {0.0, 0.0, 0.0};
    if (!GetBackgroundFieldsGCA(x, node, E, B, NULL)) return false;
    AbsB = sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]) + 1E-15;
    double vE[3];
    vE[0] = E[1] * B[2] - E[2] * B[1];
    vE[1] = E[2] * B[0] - E[0] * B[2];
    vE[2] = E[0] * B[1] - E[1] * B[0];
    if (AbsB > 0.0)
      for (int idim = 0; idim < 3; idim++) vE[idim] /= AbsB * AbsB;
    double vE_norm = sqrt(vE[0] * vE[0] + vE[1] * vE[1] + vE[2] * vE[2]);
    double v_norm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double kappa, gamma, gamma_star, c2 = cfg.speed_of_light * cfg.speed_of_light;
    kappa = 1 / sqrt(1 - vE_norm * vE_norm / (c2));
    gamma = 1 / sqrt(1 - v_norm * v_norm / (c2));
    double m0, mu = 0.0;
    double B_star[3], v_star[3];
    gamma_star = gamma / kappa;
    double vE_cross_E[3], B_dot_vE = B[0] * vE[0] + B[1] * vE[1] + B[2] * vE[2];
    vE_cross_E[0] = vE[1] * E[2] - vE[2] * E[1];
    vE_cross_E[1] = vE[2] * E[0] - vE[0] * E[2];
    vE_cross_E[2] = vE[0] * E[1] - vE[1] * E[0];
    for (int idim = 0; idim < 3; idim++) {
      B_star[idim] = kappa * (B[idim] - vE_cross_E[idim] / c2);
      if (vE_norm > 0) B_star[idim] -= (kappa - 1) * B_dot_vE * vE[idim] / (vE_norm * vE_norm);
      v_star[idim] = v[idim] - vE[idim];
    }
    double Bstar_norm = sqrt(B_star[0] * B_star[0] + B_star[1] * B_star[1] + B_star[2] * B_star[2]);
    if (Bstar_norm > 0.0) {
      double vstar_par;
      double vstar_norm = sqrt(v_star[0] * v_star[0] + v_star[1] * v_star[1] + v_star[2] * v_star[2]);
      vstar_par = (v_star[0] * B_star[0] + v_star[1] * B_star[1] + v_star[2] * B_star[2]) / Bstar_norm;
      m0 = cfg.mass[spec];
      mu = 0.5 * (gamma_star * gamma_star) * m0 * (vstar_norm * vstar_norm - vstar_par * vstar_par) / Bstar_norm;
    }
    SetMagneticMoment(mu, ParticleData);
    return true;
  }

  // PIC::Mover::Relativistic::GuidingCenter::Mover_FirstOrder, pic_mover_relativistic_guiding_center.cpp:96-409
  // (DELETE boundary; the reference's USER_FUNCTION branch exit()s and its SPECULAR branch uses an undefined face)
  int RelGCA_Mover_FirstOrder(byte *ParticleData, long int ptr, double dtTotal, cTreeNode *startNode, int nThreads, int thread, cTreeNode **newNodeOut) {
    cTreeNode *newNode = NULL;
    double mass, ElectricCharge;
    int i, j, k, spec;
    double vInit[3], xInit[3], xFinal[3], vFinal[3] = {0.0, 0.0, 0.0};
    double var15[15];
    double *b_dot_grad_b = var15, *vE_dot_grad_b = var15 + 3, *b_dot_grad_vE = var15 + 6, *vE_dot_grad_vE = var15 + 9, *grad_kappaB = var15 + 12;
    double B[3], E[3];
    GetV(vInit, ParticleData);
    GetX(xInit, ParticleData);
    spec = GetI(ParticleData);
    ElectricCharge = cfg.charge[spec];
    mass = cfg.mass[spec];
    double vNorm = sqrt(vInit[0] * vInit[0] + vInit[1] * vInit[1] + vInit[2] * vInit[2]);
    double c2 = cfg.speed_of_light * cfg.speed_of_light;
    double lfac = 1 / sqrt(1.0 - vNorm * vNorm / c2);
    if (!GetBackgroundFieldsGCA(xInit, startNode, E, B, var15)) return _ORACLE_ERROR_;
    double bNorm = sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
    double bHat[3] = {0.0, 0.0, 0.0};
    double ePar = 0.0;
    if (bNorm > 0.0) {
      for (int idim = 0; idim < 3; idim++) {
        bHat[idim] = B[idim] / bNorm;
        ePar += E[idim] * bHat[idim];
      }
    }
    double vPar = vInit[0] * bHat[0] + vInit[1] * bHat[1] + vInit[2] * bHat[2];
    double vPerp = sqrt(vNorm * vNorm - vPar * vPar);
    (void)vPerp;
    double uPar = lfac * vPar;
    double Mr = GetMagneticMoment(ParticleData);
    double vE[3], vENorm = 0.0;
    vE[0] = E[1] * bHat[2] - E[2] * bHat[1];
    vE[1] = E[2] * bHat[0] - E[0] * bHat[2];
    vE[2] = E[0] * bHat[1] - E[1] * bHat[0];
    if (bNorm > 0.0) {
      for (int idim = 0; idim < 3; idim++) {
        vE[idim] = vE[idim] / bNorm;
        vENorm += vE[idim] * vE[idim];
      }
    }
    vENorm = sqrt(vENorm);
    double kappa, gamma;
    kappa = 1 / sqrt(1 - vENorm * vENorm / c2);
    gamma = sqrt(1.0 + (uPar * uPar + 2.0 * Mr * bNorm / mass) / c2) * kappa;
    double utmp1[3] = {0.0, 0.0, 0.0}, utmp2[3], utmp3[3];
    double temp = bNorm / (kappa * kappa);
    if (bNorm > 0.0)
      for (int idim = 0; idim < 3; idim++) utmp1[idim] = bHat[idim] / temp;
    for (int idim = 0; idim < 3; idim++) {
      utmp2[idim] = Mr / (gamma * ElectricCharge) * grad_kappaB[idim] +
                    mass / ElectricCharge * (uPar * uPar / gamma * b_dot_grad_b[idim] + uPar * vE_dot_grad_b[idim] + uPar * b_dot_grad_vE[idim] + gamma * vE_dot_grad_vE[idim]);
    }
    for (int idim = 0; idim < 3; idim++) utmp2[idim] = utmp2[idim] + uPar * ePar / (gamma)*vE[idim];
    double u[3];
    utmp3[0] = utmp1[1] * utmp2[2] - utmp1[2] * utmp2[1];
    utmp3[1] = utmp1[2] * utmp2[0] - utmp1[0] * utmp2[2];
    utmp3[2] = utmp1[0] * utmp2[1] - utmp1[1] * utmp2[0];
    for (int idim = 0; idim < 3; idim++) {
      u[idim] = vE[idim] + utmp3[idim];
      u[idim] += uPar / gamma * bHat[idim];
    }
    double dupardt;
    dupardt = ElectricCharge / mass * ePar;
    temp = Mr / (mass * gamma);
    for (int idim = 0; idim < 3; idim++) {
      dupardt += -temp * bHat[idim] * grad_kappaB[idim] + vE[idim] * (uPar * b_dot_grad_b[idim] + gamma * vE_dot_grad_b[idim]);
      xFinal[idim] = xInit[idim] + dtTotal * u[idim];
    }
    uPar += dupardt * dtTotal;

    if (cfg.internal_sphere_radius > 0.0) {
      double rFinal = sqrt(xFinal[0] * xFinal[0] + xFinal[1] * xFinal[1] + xFinal[2] * xFinal[2]);
      if (rFinal < cfg.internal_sphere_radius) {
        AddExitRecord(ptr, spec, AMPS_EXIT_SPHERE, startNode, xInit, vInit);
        DeleteParticle(ptr);
        return _PARTICLE_LEFT_THE_DOMAIN_;
      }
    }
    newNode = findTreeNode(xFinal, startNode);
    if (newNode == NULL) {
      if (cfg.boundary_mode != AMPS_BOUNDARY_DELETE) return _ORACLE_ERROR_;
      DeleteParticle(ptr);
      return _PARTICLE_LEFT_THE_DOMAIN_;
    } else {
      if (newNode->block == NULL) return _ORACLE_ERROR_;
      if (!GetBackgroundFieldsGCA(xFinal, newNode, E, B, NULL)) return _ORACLE_ERROR_;
      bNorm = sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
      if (bNorm > 0.0)
        for (int idim = 0; idim < 3; idim++) bHat[idim] = B[idim] / bNorm;
      vE[0] = E[1] * bHat[2] - E[2] * bHat[1];
      vE[1] = E[2] * bHat[0] - E[0] * bHat[2];
      vE[2] = E[0] * bHat[1] - E[1] * bHat[0];
      vENorm = 0.0;
      if (bNorm > 0.0) {
        for (int idim = 0; idim < 3; idim++) {
          vE[idim] = vE[idim] / bNorm;
          vENorm += vE[idim] * vE[idim];
        }
      }
      vENorm = sqrt(vENorm);
      kappa = 1 / sqrt(1 - vENorm * vENorm / c2);
      gamma = sqrt(1.0 + (uPar * uPar + 2.0 * Mr * bNorm / mass) / c2) * kappa;
      vPar = uPar / gamma;
      double res = (1 - 1 / (gamma * gamma)) * c2 - vPar * vPar;
      if (bNorm == 0 && res < 0) res = 0.0;
      if (bNorm > 0 && res < 0) {
        DeleteParticle(ptr);
        return _PARTICLE_LEFT_THE_DOMAIN_;
      }
      vPerp = sqrt(res);
      double e0[3] = {1.0, 0.0, 0.0}, e1[3] = {0.0, 1.0, 0.0};
      double ePerp[3] = {1.0, 0.0, 0.0};
      double diff = (e0[0] - bHat[0]) * (e0[0] - bHat[0]) + (e0[1] - bHat[1]) * (e0[1] - bHat[1]) + (e0[2] - bHat[2]) * (e0[2] - bHat[2]);
      const double *ee = (diff > 0.0) ? e0 : e1;
      ePerp[0] = ee[1] * bHat[2] - ee[2] * bHat[1];
      ePerp[1] = ee[2] * bHat[0] - ee[0] * bHat[2];
      ePerp[2] = ee[0] * bHat[1] - ee[1] * bHat[0];
      for (int idim = 0; idim < 3; idim++) vFinal[idim] += vPerp * ePerp[idim] + vPar * bHat[idim];
    }
    cBlock *block;
    if (FindCellIndex(xFinal, i, j, k, newNode) == -1) return _ORACLE_ERROR_;
    if ((block = newNode->block) == NULL) return _ORACLE_ERROR_;
    AttachToTempList(ptr, ParticleData, block, i, j, k, nThreads, thread);
    SetV(vFinal, ParticleData);
    SetX(xFinal, ParticleData);
    *newNodeOut = newNode;
    return _PARTICLE_MOTION_FINISHED_;
  }


  // ------------------------------------------------------------------------------------------
  // a8: PIC::Mover::GuidingCenter, src/pic/pic_mover_guiding_center.cpp  (coupler mode, relativity off)
  // ------------------------------------------------------------------------------------------
  // PIC::CPLR::InitInterpolationStencil(x,node) for a point that may lie OUTSIDE `node` (Mover_FirstOrder :713 builds the
  // stencil of the new position in the start block).  Indices beyond the block's ghost layer are out-of-bounds reads in
  // the reference -> reported as an error here (returns false), like every place where the reference exit()s.
  // ---- the ECSIM field getters the guiding-centre movers use when _PIC_FIELD_SOLVER_MODE_ is ECSIM (cfg.gc_fields_ecsim) ----
  // ECSIM::GetElectricField, pic_field_solver_ecsim.cpp:7440-7453: corner stencil on E (slot 0 of the corner data = the current E).
  // (The reference hands the mover's own x to CornerBased::InitStencil, which snaps a point closer than 1e-10 dx to the block's upper
  // face; here the snap stays local to the stencil.)
  bool ECSIM_GetElectricField(double *E, const double *x, cTreeNode *node) const {
    cStencil Stencil;
    double xx[3] = {x[0], x[1], x[2]}, Wdummy[8];
    for (int idim = 0; idim < 3; idim++) E[idim] = 0.0;
    if (node == NULL || node->block == NULL) return false;
    if (!CornerBased_InitStencil(xx, node, Stencil, Wdummy)) return false;
    for (int iCornerNode = 0; iCornerNode < Stencil.Length; iCornerNode++) {
      const double *t = node->block->cornerNodes[Stencil.LocalCellID[iCornerNode]]->data + ExOffsetIndex;
      const double w = Stencil.Weight[iCornerNode];
      for (int idim = 0; idim < 3; idim++) E[idim] += w * t[idim];
    }
    return true;
  }
  // ECSIM::GetMagneticField, :7456-7469: centre stencil on slot 0 of the centre data (the reference's CurrentBOffset / PrevBOffset swap
  // every step, :5697-5700, while this getter always reads slot 0; here slot 0 is B_cur)
  bool ECSIM_GetMagneticField(double *B, const double *x, cTreeNode *node) const {
    cStencil Stencil;
    for (int idim = 0; idim < 3; idim++) B[idim] = 0.0;
    if (node == NULL || node->block == NULL) return false;
    CellCentered_Linear_InitStencil(x, node, Stencil, globalStencilLength != 8);
    for (int iCenterNode = 0; iCenterNode < Stencil.Length; iCenterNode++) {
      const double *t = node->block->centerNodes[Stencil.LocalCellID[iCenterNode]]->data + CurrentBOffset_d;
      const double w = Stencil.Weight[iCenterNode];
      for (int idim = 0; idim < 3; idim++) B[idim] += w * t[idim];
    }
    return true;
  }
  // ECSIM::GetMagneticFieldGradient, :7473-7547: central differences over half a cell, one-sided next to the domain boundary
  bool ECSIM_GetMagneticFieldGradient(double *gradB, const double *x, cTreeNode *node) const {
    double x_plus[3], x_minus[3], dx;
    double B_plus[3], B_minus[3], B0[3];
    if (!ECSIM_GetMagneticField(B0, x, node)) return false;
    for (int idim = 0; idim < 3; idim++) {
      memcpy(x_plus, x, 3 * sizeof(double));
      memcpy(x_minus, x, 3 * sizeof(double));
      if (idim == 0) dx = 0.5 * (node->xmax[0] - node->xmin[0]) / _BLOCK_CELLS_X_;
      else if (idim == 1) dx = 0.5 * (node->xmax[1] - node->xmin[1]) / _BLOCK_CELLS_Y_;
      else dx = 0.5 * (node->xmax[2] - node->xmin[2]) / _BLOCK_CELLS_Z_;
      x_plus[idim] += dx;
      x_minus[idim] -= dx;
      cTreeNode *node_plus = findTreeNode(x_plus, node), *node_minus = findTreeNode(x_minus, node);
      const bool has_plus = (node_plus != NULL), has_minus = (node_minus != NULL);
      if (has_plus && !ECSIM_GetMagneticField(B_plus, x_plus, node_plus)) return false;
      if (has_minus && !ECSIM_GetMagneticField(B_minus, x_minus, node_minus)) return false;
      if (has_plus && has_minus) {
        gradB[0 + idim] = (B_plus[0] - B_minus[0]) / (2.0 * dx);
        gradB[3 + idim] = (B_plus[1] - B_minus[1]) / (2.0 * dx);
        gradB[6 + idim] = (B_plus[2] - B_minus[2]) / (2.0 * dx);
      } else if (has_plus) {
        gradB[0 + idim] = (B_plus[0] - B0[0]) / dx;
        gradB[3 + idim] = (B_plus[1] - B0[1]) / dx;
        gradB[6 + idim] = (B_plus[2] - B0[2]) / dx;
      } else if (has_minus) {
        gradB[0 + idim] = (B0[0] - B_minus[0]) / dx;
        gradB[3 + idim] = (B0[1] - B_minus[1]) / dx;
        gradB[6 + idim] = (B0[2] - B_minus[2]) / dx;
      } else {
        gradB[0 + idim] = 0.0;
        gradB[3 + idim] = 0.0;
        gradB[6 + idim] = 0.0;
      }
    }
    return true;
  }

  bool GC_InitStencil(const double *x, cTreeNode *node, cStencil &Stencil) const { return CplrInitStencil(x, node, Stencil); }
  void GC_Gather(const cStencil &Stencil, cTreeNode *node, int offset, int nVars, double *out) const {
    (void)node;
    CplrGather(Stencil, offset, nVars, out);
  }

  // InitiateMagneticMoment, :85-144: mu from the perpendicular speed, then v is ALIGNED with B
  bool GC_InitiateMagneticMoment(int spec, const double *x, double *v, byte *ParticleData, cTreeNode *node) {
    double B[3] = {0.0, 0.0, 0.0}, AbsB = 0.0;
    if (cfg.gc_fields_ecsim) {  // :103-104
      if (!ECSIM_GetMagneticField(B, x, node)) return false;
    } else {
      cStencil Stencil;
      if (!GC_InitStencil(x, node, Stencil)) return false;
      GC_Gather(Stencil, node, BackgroundB_d, 3, B);
    }
    AbsB = sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]) + 1E-15;
    double v_par = 0.0, v2, gamma2, m0, mu = 0.0;
    double b[3] = {B[0] / AbsB, B[1] / AbsB, B[2] / AbsB};
    if (AbsB > 0.0) {
      v_par = v[0] * b[0] + v[1] * b[1] + v[2] * b[2];
      v2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
      gamma2 = 1.0;
      m0 = cfg.mass[spec];
      mu = 0.5 * gamma2 * m0 * (v2 - v_par * v_par) / AbsB;
    }
    v[0] = v_par * b[0];
    v[1] = v_par * b[1];
    v[2] = v_par * b[2];
    SetMagneticMoment(mu, ParticleData);
    return true;
  }

  // GuidingCenterMotion_default, :146-289
  bool GC_GuidingCenterMotion(double *Vguide_perp, double &ForceParal, double &BAbsoluteValue, double *BDirection, const double *PParal, int spec,
                              double mu, const double *x, const double *v, cTreeNode *startNode) const {
    double Vguide_perp_LOC[3] = {0.0, 0.0, 0.0}, ForceParal_LOC = 0.0;
    double E[3], gradB[9], gradAbsB[3], AbsB = 0.0;
    double b[3], B[3];
    if (cfg.gc_fields_ecsim) {  // :179-184
      if (!ECSIM_GetMagneticField(B, x, startNode)) return false;
      if (!ECSIM_GetElectricField(E, x, startNode)) return false;
      if (!ECSIM_GetMagneticFieldGradient(gradB, x, startNode)) return false;
    } else {
      cStencil Stencil;
      if (!GC_InitStencil(x, startNode, Stencil)) return false;
      GC_Gather(Stencil, startNode, BackgroundE_d, 3, E);
      GC_Gather(Stencil, startNode, BackgroundB_d, 3, B);
      GC_Gather(Stencil, startNode, BackgroundGradB_d, 9, gradB);
    }
    AbsB = pow(B[0] * B[0] + B[1] * B[1] + B[2] * B[2], 0.5) + 1E-15;
    b[0] = B[0] / AbsB;
    b[1] = B[1] / AbsB;
    b[2] = B[2] / AbsB;
    gradAbsB[0] = b[0] * gradB[0] + b[1] * gradB[3] + b[2] * gradB[6];
    gradAbsB[1] = b[0] * gradB[1] + b[1] * gradB[4] + b[2] * gradB[7];
    gradAbsB[2] = b[0] * gradB[2] + b[1] * gradB[5] + b[2] * gradB[8];
    double q = cfg.charge[spec];
    double m0 = cfg.mass[spec];
    double gamma = 1.0;
    double p_par;
    p_par = (PParal == NULL) ? gamma * m0 * (v[0] * b[0] + v[1] * b[1] + v[2] * b[2]) : *PParal;
    double msc, vec[3] = {0.0, 0.0, 0.0};
    Vguide_perp_LOC[0] += (E[1] * b[2] - E[2] * b[1]) / AbsB;
    Vguide_perp_LOC[1] += (E[2] * b[0] - E[0] * b[2]) / AbsB;
    Vguide_perp_LOC[2] += (E[0] * b[1] - E[1] * b[0]) / AbsB;
    msc = mu / (q * gamma) / AbsB;
    Vguide_perp_LOC[0] += msc * (b[1] * gradAbsB[2] - b[2] * gradAbsB[1]);
    Vguide_perp_LOC[1] += msc * (b[2] * gradAbsB[0] - b[0] * gradAbsB[2]);
    Vguide_perp_LOC[2] += msc * (b[0] * gradAbsB[1] - b[1] * gradAbsB[0]);
    msc = p_par * p_par / (q * gamma * m0) / AbsB / AbsB;
    vec[0] = b[0] * gradB[0] + b[1] * gradB[1] + b[2] * gradB[2];
    vec[1] = b[0] * gradB[3] + b[1] * gradB[4] + b[2] * gradB[5];
    vec[2] = b[0] * gradB[6] + b[1] * gradB[7] + b[2] * gradB[8];
    Vguide_perp_LOC[0] += msc * (b[1] * vec[2] - b[2] * vec[1]);
    Vguide_perp_LOC[1] += msc * (b[2] * vec[0] - b[0] * vec[2]);
    Vguide_perp_LOC[2] += msc * (b[0] * vec[1] - b[1] * vec[0]);
    if (cfg.ideal_mhd) {  // _PIC__IDEAL_MHD_MODE_ (picGlobal.dfn:339, default ON): E.b = 0
      ForceParal_LOC = -mu / gamma * (gradAbsB[0] * b[0] + gradAbsB[1] * b[1] + gradAbsB[2] * b[2]);
    } else {
      ForceParal_LOC = q * (E[0] * b[0] + E[1] * b[1] + E[2] * b[2]) - mu / gamma * (gradAbsB[0] * b[0] + gradAbsB[1] * b[1] + gradAbsB[2] * b[2]);
    }
    memcpy(Vguide_perp, Vguide_perp_LOC, 3 * sizeof(double));
    memcpy(BDirection, b, 3 * sizeof(double));
    ForceParal = ForceParal_LOC;
    BAbsoluteValue = AbsB;
    return true;
  }

  // Mover_FirstOrder, :622-849
  int GC_Mover_FirstOrder(byte *ParticleData, long int ptr, double dtTotal, cTreeNode *startNode, int nThreads, int thread, cTreeNode **newNodeOut) {
    cTreeNode *newNode = NULL;
    double AbsBInit = 0.0, bInit[3] = {0.0, 0.0, 0.0};
    double v[3], p = 0.0, x[3];
    int idim, i, j, k, spec;
    double misc, mu;
    GetV(v, ParticleData);
    GetX(x, ParticleData);
    spec = GetI(ParticleData);
    double m0 = cfg.mass[spec];
    if (TestInitFlag(ParticleData) == false) {
      SetInitFlag(true, ParticleData);
      if (!GC_InitiateMagneticMoment(spec, x, v, ParticleData, startNode)) return _ORACLE_ERROR_;
    }
    mu = GetMagneticMoment(ParticleData);
    double Vguide_perpInit[3] = {0.0, 0.0, 0.0}, ForceParalInit = 0.0;
    if (!GC_GuidingCenterMotion(Vguide_perpInit, ForceParalInit, AbsBInit, bInit, NULL, spec, mu, x, v, startNode)) return _ORACLE_ERROR_;
    misc = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (v[0] * bInit[0] + v[1] * bInit[1] + v[2] * bInit[2] < 0.0) misc *= -1.0;
    for (idim = 0; idim < 3; idim++) v[idim] = misc * bInit[idim];
    p = m0 * (v[0] * bInit[0] + v[1] * bInit[1] + v[2] * bInit[2]);
    for (idim = 0; idim < 3; idim++) x[idim] += dtTotal * (Vguide_perpInit[idim] + v[idim]);
    p += dtTotal * ForceParalInit;
    newNode = findTreeNode(x, NULL);  // PIC::Mesh::Search::FindBlock
    if (newNode == NULL) {
      DeleteParticle(ptr);
      return _PARTICLE_LEFT_THE_DOMAIN_;
    }
    double bFinal[3];
    if (cfg.gc_fields_ecsim) {  // the NEW node in this branch (:727-729)
      if (!ECSIM_GetMagneticField(bFinal, x, newNode)) return _ORACLE_ERROR_;
    } else {
      cStencil Stencil;
      if (!GC_InitStencil(x, startNode, Stencil)) return _ORACLE_ERROR_;  // the START node, as written (:713)
      GC_Gather(Stencil, startNode, BackgroundB_d, 3, bFinal);
    }
    Vector3D_Normalize(bFinal);
    misc = p / m0;
    for (idim = 0; idim < 3; idim++) v[idim] = misc * bFinal[idim];
    if (cfg.internal_sphere_radius > 0.0) {
      double rFinal2;
      if ((rFinal2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2]) < cfg.internal_sphere_radius * cfg.internal_sphere_radius) {
        DeleteParticle(ptr);  // the sphere callback is commented out in this mover (:748-758)
        return _PARTICLE_LEFT_THE_DOMAIN_;
      } else
        newNode = findTreeNode(x, startNode);
    } else
      newNode = findTreeNode(x, startNode);
    if (newNode == NULL) return _ORACLE_ERROR_;
    if (FindCellIndex(x, i, j, k, newNode) == -1) return _ORACLE_ERROR_;
    cBlock *block;
    if ((block = newNode->block) == NULL) return _ORACLE_ERROR_;
    AttachToTempList(ptr, ParticleData, block, i, j, k, nThreads, thread);
    SetV(v, ParticleData);
    SetX(x, ParticleData);
    *newNodeOut = newNode;
    return _PARTICLE_MOTION_FINISHED_;
  }

  // ------------------------------------------------------------------------------------------
  // f2: PIC::GYROKINETIC, src/pic/gyro/gyro_mover.cpp (coupler fields; the reduced state is (x, v_parallel, mu))
  // ------------------------------------------------------------------------------------------
  // EvalRHS, :245-336 (GetEBandGradB :196-218 coupler branch, GetAbsBAndUnitB :220-232, GetGradAbsB :234-243)
  bool Gyro_EvalRHS(const double *x, cTreeNode *node, double vpar, double mu, int spec, double &absB, double *b, double *vdrift, double &dvpar_dt) const {
    absB = 0.0;
    b[0] = 0.0, b[1] = 0.0, b[2] = 0.0;
    vdrift[0] = 0.0, vdrift[1] = 0.0, vdrift[2] = 0.0;
    dvpar_dt = 0.0;
    double E[3], B[3], gradB[9];
    if (node == NULL) return true;
    cStencil Stencil;
    if (!GC_InitStencil(x, node, Stencil)) return false;  // (out-of-bounds read in the reference)
    GC_Gather(Stencil, node, BackgroundE_d, 3, E);
    GC_Gather(Stencil, node, BackgroundB_d, 3, B);
    GC_Gather(Stencil, node, BackgroundGradB_d, 9, gradB);
    const double absB2_ = B[0] * B[0] + B[1] * B[1] + B[2] * B[2];
    if (absB2_ <= 0.0) return true;
    absB = sqrt(absB2_);
    const double inv = 1.0 / absB;
    b[0] = B[0] * inv, b[1] = B[1] * inv, b[2] = B[2] * inv;
    const double m = cfg.mass[spec], q = cfg.charge[spec];
    const double absB2 = absB * absB;
    const double invAbsB2 = 1.0 / absB2;
    double gradAbsB[3];
    {
      const double invAbsB = 1.0 / absB;
      for (int j = 0; j < 3; j++) gradAbsB[j] = (B[0] * gradB[0 * 3 + j] + B[1] * gradB[1 * 3 + j] + B[2] * gradB[2 * 3 + j]) * invAbsB;
    }
    const double Epar = E[0] * b[0] + E[1] * b[1] + E[2] * b[2];
    const double bDotGradAbsB = b[0] * gradAbsB[0] + b[1] * gradAbsB[1] + b[2] * gradAbsB[2];
    dvpar_dt = (q / m) * Epar - (mu / m) * bDotGradAbsB;
    double ExB[3] = {E[1] * B[2] - E[2] * B[1], E[2] * B[0] - E[0] * B[2], E[0] * B[1] - E[1] * B[0]};
    vdrift[0] = ExB[0] * invAbsB2;
    vdrift[1] = ExB[1] * invAbsB2;
    vdrift[2] = ExB[2] * invAbsB2;
    if (q != 0.0 && mu != 0.0) {
      double BxGradAbsB[3] = {B[1] * gradAbsB[2] - B[2] * gradAbsB[1], B[2] * gradAbsB[0] - B[0] * gradAbsB[2], B[0] * gradAbsB[1] - B[1] * gradAbsB[0]};
      const double c = (mu / q) * invAbsB2;
      vdrift[0] += c * BxGradAbsB[0];
      vdrift[1] += c * BxGradAbsB[1];
      vdrift[2] += c * BxGradAbsB[2];
    }
    double BB[3];
    BB[0] = B[0] * gradB[0 * 3 + 0] + B[1] * gradB[0 * 3 + 1] + B[2] * gradB[0 * 3 + 2];
    BB[1] = B[0] * gradB[1 * 3 + 0] + B[1] * gradB[1 * 3 + 1] + B[2] * gradB[1 * 3 + 2];
    BB[2] = B[0] * gradB[2 * 3 + 0] + B[1] * gradB[2 * 3 + 1] + B[2] * gradB[2 * 3 + 2];
    if (q != 0.0 && vpar != 0.0) {
      double BxBB[3] = {B[1] * BB[2] - B[2] * BB[1], B[2] * BB[0] - B[0] * BB[2], B[0] * BB[1] - B[1] * BB[0]};
      const double invAbsB4 = 1.0 / (absB2 * absB2);
      const double c = (m * vpar * vpar / q) * invAbsB4;
      vdrift[0] += c * BxBB[0];
      vdrift[1] += c * BxBB[1];
      vdrift[2] += c * BxBB[2];
    }
    if (!std::isfinite(vdrift[0]) || !std::isfinite(vdrift[1]) || !std::isfinite(vdrift[2]) || !std::isfinite(dvpar_dt)) {
      vdrift[0] = 0.0, vdrift[1] = 0.0, vdrift[2] = 0.0;
      dvpar_dt = 0.0;
    }
    return true;
  }
  // CommitReducedStateAndVelocity, :338-381 (v_normal and the stored drift velocity are functions of the state: not kept here)
  bool Gyro_Commit(byte *ParticleData, int spec, const double *x, cTreeNode *node, double vpar, double mu, const double *vdrift_to_store) {
    double absB = 0.0, b[3] = {0.0, 0.0, 0.0};
    double vdrift_dummy[3] = {0.0, 0.0, 0.0};
    double dvpar_dt_dummy = 0.0;
    if (!Gyro_EvalRHS(x, node, vpar, mu, spec, absB, b, vdrift_dummy, dvpar_dt_dummy)) return false;
    SetVParallel(vpar, ParticleData);
    double v[3];
    v[0] = b[0] * vpar + vdrift_to_store[0];
    v[1] = b[1] * vpar + vdrift_to_store[1];
    v[2] = b[2] * vpar + vdrift_to_store[2];
    SetV(v, ParticleData);
    return true;
  }
  // the tail both movers share: internal sphere, findTreeNode, FindCellIndex, temp list (:451-541, :630-719)
  int Gyro_File(byte *ParticleData, long int ptr, const double *x, cTreeNode *startNode, int nThreads, int thread, cTreeNode **newNodeOut) {
    cTreeNode *newNode;
    if (cfg.internal_sphere_radius > 0.0) {
      double r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
      if (r2 < cfg.internal_sphere_radius * cfg.internal_sphere_radius) {
        DeleteParticle(ptr);
        return _PARTICLE_LEFT_THE_DOMAIN_;
      } else
        newNode = findTreeNode(x, startNode);
    } else
      newNode = findTreeNode(x, startNode);
    int i, j, k;
    cBlock *block;
    if (newNode == NULL) return _ORACLE_ERROR_;
    if (FindCellIndex(x, i, j, k, newNode) == -1) return _ORACLE_ERROR_;  // exit("cannot find cell index for moved particle")
    if ((block = newNode->block) == NULL) return _ORACLE_ERROR_;          // exit("destination block is empty")
    AttachToTempList(ptr, ParticleData, block, i, j, k, nThreads, thread);
    *newNodeOut = newNode;
    return _PARTICLE_MOTION_FINISHED_;
  }
  // Mover_FirstOrder, :383-544
  int Gyro_Mover_FirstOrder(byte *ParticleData, long int ptr, double dtTotal, cTreeNode *startNode, int nThreads, int thread, cTreeNode **newNodeOut) {
    double x[3];
    GetX(x, ParticleData);
    int spec = GetI(ParticleData);
    const double mu = GetMagneticMoment(ParticleData);
    double vpar = GetVParallel(ParticleData);
    double absB = 0.0, b[3] = {0.0, 0.0, 0.0};
    double vdrift[3] = {0.0, 0.0, 0.0};
    double dvpar_dt = 0.0;
    if (!Gyro_EvalRHS(x, startNode, vpar, mu, spec, absB, b, vdrift, dvpar_dt)) return _ORACLE_ERROR_;
    x[0] += dtTotal * (vdrift[0] + b[0] * vpar);
    x[1] += dtTotal * (vdrift[1] + b[1] * vpar);
    x[2] += dtTotal * (vdrift[2] + b[2] * vpar);
    SetX(x, ParticleData);  // (x is a pointer into the record in the reference)
    vpar += dtTotal * dvpar_dt;
    cTreeNode *newNode = findTreeNode(x, NULL);  // PIC::Mesh::Search::FindBlock
    if (newNode == NULL) {
      DeleteParticle(ptr);
      return _PARTICLE_LEFT_THE_DOMAIN_;
    }
    double absB1 = 0.0, b1[3] = {0.0, 0.0, 0.0};
    double vdrift1[3] = {0.0, 0.0, 0.0};
    double dvpar_dt_dummy = 0.0;
    if (!Gyro_EvalRHS(x, newNode, vpar, mu, spec, absB1, b1, vdrift1, dvpar_dt_dummy)) return _ORACLE_ERROR_;
    if (!Gyro_Commit(ParticleData, spec, x, newNode, vpar, mu, vdrift1)) return _ORACLE_ERROR_;
    return Gyro_File(ParticleData, ptr, x, startNode, nThreads, thread, newNodeOut);
  }
  // Mover_SecondOrder, :544-720 (midpoint)
  int Gyro_Mover_SecondOrder(byte *ParticleData, long int ptr, double dtTotal, cTreeNode *startNode, int nThreads, int thread, cTreeNode **newNodeOut) {
    double x[3];
    GetX(x, ParticleData);
    int spec = GetI(ParticleData);
    const double mu = GetMagneticMoment(ParticleData);
    const double vpar0 = GetVParallel(ParticleData);
    const double x0[3] = {x[0], x[1], x[2]};
    double absB0 = 0.0, b0[3] = {0.0, 0.0, 0.0};
    double vdrift0[3] = {0.0, 0.0, 0.0};
    double dvpar_dt0 = 0.0;
    if (!Gyro_EvalRHS(x0, startNode, vpar0, mu, spec, absB0, b0, vdrift0, dvpar_dt0)) return _ORACLE_ERROR_;
    double xHalf[3];
    xHalf[0] = x0[0] + 0.5 * dtTotal * (vdrift0[0] + b0[0] * vpar0);
    xHalf[1] = x0[1] + 0.5 * dtTotal * (vdrift0[1] + b0[1] * vpar0);
    xHalf[2] = x0[2] + 0.5 * dtTotal * (vdrift0[2] + b0[2] * vpar0);
    const double vparHalf = vpar0 + 0.5 * dtTotal * dvpar_dt0;
    cTreeNode *nodeHalf = findTreeNode(xHalf, NULL);  // FindBlock
    if (nodeHalf == NULL) {
      DeleteParticle(ptr);
      return _PARTICLE_LEFT_THE_DOMAIN_;
    }
    double absBH = 0.0, bH[3] = {0.0, 0.0, 0.0};
    double vdriftH[3] = {0.0, 0.0, 0.0};
    double dvpar_dtH = 0.0;
    if (!Gyro_EvalRHS(xHalf, nodeHalf, vparHalf, mu, spec, absBH, bH, vdriftH, dvpar_dtH)) return _ORACLE_ERROR_;
    x[0] = x0[0] + dtTotal * (vdriftH[0] + bH[0] * vparHalf);
    x[1] = x0[1] + dtTotal * (vdriftH[1] + bH[1] * vparHalf);
    x[2] = x0[2] + dtTotal * (vdriftH[2] + bH[2] * vparHalf);
    SetX(x, ParticleData);
    const double vpar1 = vpar0 + dtTotal * dvpar_dtH;
    cTreeNode *newNode = findTreeNode(x, NULL);  // FindBlock
    if (newNode == NULL) {
      DeleteParticle(ptr);
      return _PARTICLE_LEFT_THE_DOMAIN_;
    }
    double absB1 = 0.0, b1[3] = {0.0, 0.0, 0.0};
    double vdrift1[3] = {0.0, 0.0, 0.0};
    double dvpar_dt_dummy = 0.0;
    if (!Gyro_EvalRHS(x, newNode, vpar1, mu, spec, absB1, b1, vdrift1, dvpar_dt_dummy)) return _ORACLE_ERROR_;
    if (!Gyro_Commit(ParticleData, spec, x, newNode, vpar1, mu, vdrift1)) return _ORACLE_ERROR_;
    return Gyro_File(ParticleData, ptr, x, startNode, nThreads, thread, newNodeOut);
  }

  // Mover_SecondOrder, :292-619 (predictor-corrector)
  int GC_Mover_SecondOrder(byte *ParticleData, long int ptr, double dtTotal, cTreeNode *startNode, int nThreads, int thread, cTreeNode **newNodeOut) {
    cTreeNode *newNode = NULL;
    double dtTemp;
    double AbsBInit = 0.0, bInit[3] = {0.0, 0.0, 0.0};
    double vInit[3] = {0.0, 0.0, 0.0}, pInit = 0.0, xInit[3] = {0.0, 0.0, 0.0};
    double AbsBMiddle = 0.0, bMiddle[3] = {0.0, 0.0, 0.0};
    double vMiddle[3] = {0.0, 0.0, 0.0}, pMiddle = 0.0, xMiddle[3] = {0.0, 0.0, 0.0};
    double vFinal[3] = {0.0, 0.0, 0.0}, pFinal = 0.0, xFinal[3] = {0.0, 0.0, 0.0};
    int i, j, k, spec;
    double misc;
    GetV(vInit, ParticleData);
    GetX(xInit, ParticleData);
    spec = GetI(ParticleData);
    double m0 = cfg.mass[spec];
    double mu = GetMagneticMoment(ParticleData);
    double Vguide_perpInit[3] = {0.0, 0.0, 0.0}, ForceParalInit = 0.0;
    if (!GC_GuidingCenterMotion(Vguide_perpInit, ForceParalInit, AbsBInit, bInit, NULL, spec, mu, xInit, vInit, startNode)) return _ORACLE_ERROR_;
    misc = pow(vInit[0] * vInit[0] + vInit[1] * vInit[1] + vInit[2] * vInit[2], 0.5);
    if (vInit[0] * bInit[0] + vInit[1] * bInit[1] + vInit[2] * bInit[2] < 0) misc *= -1.0;
    vInit[0] = misc * bInit[0];
    vInit[1] = misc * bInit[1];
    vInit[2] = misc * bInit[2];
    pInit = m0 * (vInit[0] * bInit[0] + vInit[1] * bInit[1] + vInit[2] * bInit[2]);
    dtTemp = dtTotal / 2.0;
    xMiddle[0] = xInit[0] + dtTemp * (Vguide_perpInit[0] + vInit[0]);
    xMiddle[1] = xInit[1] + dtTemp * (Vguide_perpInit[1] + vInit[1]);
    xMiddle[2] = xInit[2] + dtTemp * (Vguide_perpInit[2] + vInit[2]);
    pMiddle = pInit + dtTemp * ForceParalInit;
    newNode = findTreeNode(xMiddle, NULL);  // FindBlock
    if (newNode == NULL) {
      DeleteParticle(ptr);
      return _PARTICLE_LEFT_THE_DOMAIN_;
    }
    double Vguide_perpMiddle[3] = {0.0, 0.0, 0.0}, ForceParalMiddle = 0.0;
    if (!GC_GuidingCenterMotion(Vguide_perpMiddle, ForceParalMiddle, AbsBMiddle, bMiddle, &pMiddle, spec, mu, xMiddle, vMiddle, newNode)) return _ORACLE_ERROR_;
    misc = pMiddle / m0;
    vMiddle[0] = misc * bMiddle[0];
    vMiddle[1] = misc * bMiddle[1];
    vMiddle[2] = misc * bMiddle[2];
    xFinal[0] = xInit[0] + dtTotal * (Vguide_perpMiddle[0] + vMiddle[0]);
    xFinal[1] = xInit[1] + dtTotal * (Vguide_perpMiddle[1] + vMiddle[1]);
    xFinal[2] = xInit[2] + dtTotal * (Vguide_perpMiddle[2] + vMiddle[2]);
    pFinal = pInit + dtTotal * ForceParalMiddle;
    misc = pFinal / m0;
    vFinal[0] = misc * bMiddle[0];
    vFinal[1] = misc * bMiddle[1];
    vFinal[2] = misc * bMiddle[2];
    if (cfg.internal_sphere_radius > 0.0) {
      double rFinal2;
      const double R = cfg.internal_sphere_radius;
      if ((rFinal2 = xFinal[0] * xFinal[0] + xFinal[1] * xFinal[1] + xFinal[2] * xFinal[2]) < R * R) {
        double r = sqrt(rFinal2);
        for (int idim = 0; idim < 3; idim++) xFinal[idim] *= R / r;
        newNode = findTreeNode(xFinal, startNode);
        // ParticleSphereInteraction -> _PARTICLE_DELETED_ON_THE_FACE_: handed to the host as an exit record
        AddExitRecord(ptr, spec, AMPS_EXIT_SPHERE, newNode, xFinal, vFinal);
        DeleteParticle(ptr);
        return _PARTICLE_LEFT_THE_DOMAIN_;
      } else {
        newNode = findTreeNode(xFinal, startNode);
      }
    } else
      newNode = findTreeNode(xFinal, startNode);
    if (newNode == NULL) {
      DeleteParticle(ptr);
      return _PARTICLE_LEFT_THE_DOMAIN_;
    }
    if (FindCellIndex(xFinal, i, j, k, newNode) == -1) return _ORACLE_ERROR_;
    cBlock *block;
    if ((block = newNode->block) == NULL) return _ORACLE_ERROR_;
    AttachToTempList(ptr, ParticleData, block, i, j, k, nThreads, thread);
    SetV(vFinal, ParticleData);
    SetX(xFinal, ParticleData);
    *newNodeOut = newNode;
    return _PARTICLE_MOTION_FINISHED_;
  }

  // f3: PIC::Sampling::SamplingManager + ProcessCell, src/pic/pic.cpp:1045-1082, :705-990 (velocity tensor on, parallel/tangential
  // temperature, internal degrees of freedom, dust and user sampling off).  Per cell and species, added to the collecting buffer:
  //  [0] DatumParticleWeight  [1] DatumParticleNumber  [2] DatumNumberDensity (w / cell->Measure, Measure = cell volume)
  //  [3..5] DatumParticleVelocity (w v)  [6..8] DatumParticleVelocity2 (w v_i^2)  [9] DatumParticleSpeed (w |v|)
  //  [10..12] DatumParticleVelocity2Tensor (w v_i v_{(i+1)%3})
  int SampleCells() {
    const int nS = cfg.n_species, nC = nCellsBlock();
    if (cellSample.size() != blocks.size() * (size_t)nC * nS * 13) cellSample.assign(blocks.size() * (size_t)nC * nS * 13, 0.0);
    if (sampledParticles.size() != (size_t)nS) sampledParticles.assign(nS, 0);
    for (size_t nLocalNode = 0; nLocalNode < blocks.size(); nLocalNode++) {
      cTreeNode *node = BlockTable[nLocalNode];
      cBlock *block = node->block;
      if (!block) continue;
      double Measure = 1.0;
      Measure *= (node->xmax[0] - node->xmin[0]) / _BLOCK_CELLS_X_;
      Measure *= (node->xmax[1] - node->xmin[1]) / _BLOCK_CELLS_Y_;
      Measure *= (node->xmax[2] - node->xmin[2]) / _BLOCK_CELLS_Z_;
      for (int c = 0; c < nC; c++) {
        long int ptr = block->FirstCellParticleTable[c];
        double *cell = cellSample.data() + (nLocalNode * nC + c) * (size_t)nS * 13;
        while (ptr != -1) {
          byte *ParticleData = GetParticleDataPointer(ptr);
          double v[3], Speed2 = 0.0, miscv2[3], v2tensor[3];
          int s = GetI(ParticleData);
          GetV(v, ParticleData);
          sampledParticles[s]++;
          double LocalParticleWeight = cfg.species_weight[s];
          LocalParticleWeight *= GetIndividualStatWeightCorrection(ParticleData);
          double *d = cell + (size_t)s * 13;
          d[0] += LocalParticleWeight * 1.0;
          d[1] += 1.0 * 1.0;
          d[2] += (LocalParticleWeight / Measure) * 1.0;
          for (int idim = 0; idim < 3; idim++) {
            double v2 = v[idim] * v[idim];
            Speed2 += v2;
            miscv2[idim] = v2;
          }
          for (int i = 0; i < 3; i++) d[3 + i] += v[i] * LocalParticleWeight;
          for (int i = 0; i < 3; i++) d[6 + i] += miscv2[i] * LocalParticleWeight;
          d[9] += sqrt(Speed2) * LocalParticleWeight;
          for (int idim = 0; idim < 3; idim++) v2tensor[idim] = v[idim] * v[(idim + 1) % 3];
          for (int i = 0; i < 3; i++) d[10 + i] += v2tensor[i] * LocalParticleWeight;
          ptr = GetNext(ParticleData);
        }
      }
    }
    return AMPS_GPU_OK;
  }

  // The _PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_ part of UpdateJMassMatrix / ProcessCell, src/pic/pic_field_solver_ecsim.cpp:
  // zero :3269-3271, per particle :2270-2300, per cell :2384-2392, flush :3874-3879.  Restated as its own pass over the same
  // cells in the same order (the sums per corner see the same terms in the same order as inside ProcessCell).
  // Rho_=0 RhoUx_..RhoUz_=1..3 RhoUxUx_ RhoUyUy_ RhoUzUz_ = 4..6 RhoUxUy_ RhoUyUz_ RhoUxUz_ = 7..9 (:138-147)
  int ComputeSpeciesMoments() {
    const int nS = cfg.n_species;
    cornerSpec.assign((size_t)n_corners * 10 * nS, 0.0);
    const double length_conv = cfg.ecsim_length_conv;
    static const int cornerOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    std::vector<double> SpeciesData_GI((size_t)8 * 10 * nS), SpecData((size_t)8 * 10 * nS);
    for (size_t nLocalNode = 0; nLocalNode < blocks.size(); nLocalNode++) {
      cTreeNode *node = BlockTable[nLocalNode];
      if (node->block == NULL) continue;
      if (cfg.periodic && node->faceBoundary != 0) continue;  // boundary "ghost" block, :3815-3825
      cBlock *block = node->block;
      int nCell[3] = {_BLOCK_CELLS_X_, _BLOCK_CELLS_Y_, _BLOCK_CELLS_Z_};
      double CellVolume = 1, dx[3];
      for (int iDim = 0; iDim < 3; iDim++) dx[iDim] = (node->xmax[iDim] - node->xmin[iDim]) / nCell[iDim] * length_conv;
      for (int iDim = 0; iDim < 3; iDim++) CellVolume *= dx[iDim];
      for (int k = 0; k < _BLOCK_CELLS_Z_; k++)
        for (int j = 0; j < _BLOCK_CELLS_Y_; j++)
          for (int i = 0; i < _BLOCK_CELLS_X_; i++) {
            long int ptr = block->FirstCellParticleTable[i + _BLOCK_CELLS_X_ * (j + _BLOCK_CELLS_Y_ * k)];
            if (ptr == -1) continue;
            for (size_t q = 0; q < SpeciesData_GI.size(); q++) SpeciesData_GI[q] = 0.0, SpecData[q] = 0.0;
            cStencil CornerBasedStencil;
            while (ptr != -1) {
              byte *ParticleData = GetParticleDataPointer(ptr);
              double vInit[3], xInit[3], WeightPG[8];
              int spec = GetI(ParticleData);
              GetV(vInit, ParticleData);
              GetX(xInit, ParticleData);
              double LocalParticleWeight = cfg.species_weight[spec];
              LocalParticleWeight *= GetIndividualStatWeightCorrection(ParticleData);
              for (int idim = 0; idim < 3; idim++) vInit[idim] *= length_conv;
              double mass = cfg.mass[spec] * LocalParticleWeight;
              CornerBased_InitStencil(xInit, node, CornerBasedStencil, WeightPG);
              for (int ii = 0; ii < 8; ii++) {
                double *G = SpeciesData_GI.data() + (size_t)ii * 10 * nS + 10 * spec;
                double t = mass * WeightPG[ii];
                double t0 = t * vInit[0];
                double t1 = t * vInit[1];
                double t2 = t * vInit[2];
                G[0] += t;
                G[1] += t0;
                G[2] += t1;
                G[3] += t2;
                G[4] += t0 * vInit[0];
                G[5] += t1 * vInit[1];
                G[6] += t2 * vInit[2];
                G[7] += t0 * vInit[1];
                G[8] += t1 * vInit[2];
                G[9] += t0 * vInit[2];
              }
              ptr = GetNext(ParticleData);
            }
            for (int iCorner = 0; iCorner < 8; iCorner++)
              for (int ii = 0; ii < 10 * nS; ii++) SpecData[(size_t)iCorner * 10 * nS + ii] += SpeciesData_GI[(size_t)iCorner * 10 * nS + ii] / CellVolume;
            for (int icor = 0; icor < 8; icor++) {
              cCornerNode *cn = block->cornerNodes[_getCornerNodeLocalNumber(i + cornerOff[icor][0], j + cornerOff[icor][1], k + cornerOff[icor][2])];
              double *target = cornerSpec.data() + (size_t)(cn - cornerPool.data()) * 10 * nS;
              for (int ii = 0; ii < 10 * nS; ii++) target[ii] += SpecData[(size_t)icor * 10 * nS + ii];
            }
          }
    }
    return AMPS_GPU_OK;
  }

  // ::Relativistic::GetGyroFrequency with GetGamma and Vector3D::Length, src/general/specfunc.h:1290, :1214-1216, :763-765
  // (pinned bit for bit against the reference's header by tests/test_reference_mesh.py through oracle/_ref/libref_mesh.so)
  static double RelGyroFrequency(const double *v, double ParticleRestMass, double ElectricCharge, const double *B, double SpeedOfLight) {
    const double PiTimes2 = 6.28318530717958647692;
    return fabs(ElectricCharge) * sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]) /
           (PiTimes2 * ParticleRestMass * (1.0 / sqrt(1.0 - (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / (SpeedOfLight * SpeedOfLight))));
  }
  // Vector3D::Normalize, src/general/specfunc.h:969-981
  static void Vector3D_Normalize(double *x) {
    double l, l0 = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    if (l0 > 0.0) {
      l = 1.0 / l0;
      for (int idim = 0; idim < 3; idim++) x[idim] *= l;
    }
  }

  // ECSIM::isBoundaryCell != 0, src/pic/pic_field_solver_ecsim.cpp:6963-6999 with isFaceBoundary / isEdgeBoundary / isCornerBoundary
  // (:6487-6961), for a cell centre x inside `node`: the case tables reduce to "one of the face, edge or corner neighbours across
  // the block sides this cell touches is missing or not used in the calculation" (edge ids :6550-6790, corner id = ix+2iy+4iz).
  bool isBoundaryCellNonZero(const double *x, const double *dx, cTreeNode *node) const {
    int lo[3], hi[3];
    for (int idim = 0; idim < 3; idim++) {
      lo[idim] = fabs(x[idim] - 0.5 * dx[idim] - node->xmin[idim]) < EPS;
      hi[idim] = fabs(x[idim] + 0.5 * dx[idim] - node->xmax[idim]) < EPS;
    }
    static const int edgeId[3][2][2] = {
        // edges along x: (y side, z side) -> 0:(ymin,zmin) 1:(ymax,zmin) 2:(ymax,zmax) 3:(ymin,zmax)
        {{0, 3}, {1, 2}},
        // edges along y: (x side, z side) -> 4:(xmin,zmin) 5:(xmax,zmin) 6:(xmax,zmax) 7:(xmin,zmax)
        {{4, 7}, {5, 6}},
        // edges along z: (x side, y side) -> 8:(xmin,ymin) 9:(xmax,ymin) 10:(xmax,ymax) 11:(xmin,ymax)
        {{8, 11}, {9, 10}}};
    auto bad = [](cTreeNode *nb) { return nb == NULL || nb->IsUsedInCalculationFlag == false; };
    // side[d]: -1 none, 0 min, 1 max (a one-cell-wide block would touch both; the reference's switch matches neither pattern then)
    for (int sx = -1; sx <= 1; sx++)
      for (int sy = -1; sy <= 1; sy++)
        for (int sz = -1; sz <= 1; sz++) {
          const int s[3] = {sx, sy, sz};
          int n = 0;
          bool touched = true;
          for (int d = 0; d < 3; d++)
            if (s[d] >= 0) {
              n++;
              if (!(s[d] == 0 ? lo[d] : hi[d])) touched = false;
            }
          if (n == 0 || !touched) continue;
          cTreeNode *nb;
          if (n == 1) {
            int d = (sx >= 0) ? 0 : (sy >= 0) ? 1 : 2;
            nb = GetNeibFace(node, 2 * d + s[d], 0, 0);
          } else if (n == 2) {
            if (sx < 0) nb = GetNeibEdge(node, edgeId[0][sy][sz], 0);
            else if (sy < 0) nb = GetNeibEdge(node, edgeId[1][sx][sz], 0);
            else nb = GetNeibEdge(node, edgeId[2][sx][sy], 0);
          } else {
            nb = GetNeibCorner(node, sx + 2 * sy + 4 * sz);
          }
          if (bad(nb)) return true;
        }
    return false;
  }

  // ECSIM::CorrectParticleLocation, src/pic/pic_field_solver_ecsim.cpp:4440-4688 (+ exchangeParticleLocal :4366-4438, MPI mode):
  // species 0 is displaced along -grad(phi) / (4 pi rho_e) (limited to 0.1 dx), every particle is re-filed under its cell.
  static double interp2D(double vmm, double vpm, double vpp, double vmp, double dx, double dy) {  // :216-232
    return vmm * (1 - dx) * (1 - dy) + vpm * dx * (1 - dy) + vpp * dx * dy + vmp * (1 - dx) * dy;
  }
  int CorrectParticleLocation(double charge_conv, double mass_conv, long long *nDisplaced, long long *nDeleted) {
    const double Pi = 3.14159265358979323846264338327950288419716939937510582;  // src/general/constants.h
    const double length_conv = cfg.ecsim_length_conv;
    const int nS = cfg.n_species;
    double qom[AMPS_GPU_MAX_SPECIES];
    for (int iSp = 0; iSp < nS; iSp++) qom[iSp] = (cfg.charge[iSp] * charge_conv) / (cfg.mass[iSp] * mass_conv);
    if (cornerSpec.size() != (size_t)n_corners * 10 * nS || phiCenter.size() != (size_t)n_centers) return AMPS_GPU_ERR_STATE;
    *nDisplaced = 0, *nDeleted = 0;
    const int NX = _BLOCK_CELLS_X_, NY = _BLOCK_CELLS_Y_, NZ = _BLOCK_CELLS_Z_;
    std::vector<double> Phi((size_t)(NX + 2) * (NY + 2) * (NZ + 2));
    auto PHI = [&](int a, int b, int c) -> double & { return Phi[((size_t)a * (NY + 2) + b) * (NZ + 2) + c]; };
    for (size_t nLocalNode = 0; nLocalNode < blocks.size(); nLocalNode++) {
      cTreeNode *node = BlockTable[nLocalNode];
      if (node->block == NULL) continue;
      if (cfg.periodic && node->faceBoundary != 0) continue;  // 'ghost' block of the periodic boundary, :4459-4470
      int nCell[3] = {NX, NY, NZ};
      cBlock *block = node->block;
      long int *FirstCellParticleTable = block->FirstCellParticleTable;
      double dx[3];
      for (int iDim = 0; iDim < 3; iDim++) dx[iDim] = (node->xmax[iDim] - node->xmin[iDim]) / nCell[iDim] * length_conv;
      for (int k = -1; k < NZ + 1; k++)
        for (int j = -1; j < NY + 1; j++)
          for (int i = -1; i < NX + 1; i++) {
            int LocalCenterId = _getCenterNodeLocalNumber(i, j, k);
            PHI(i + 1, j + 1, k + 1) = 0.0;  // (left uninitialised by the reference when the node does not exist)
            if (!block->centerNodes[LocalCenterId]) continue;
            PHI(i + 1, j + 1, k + 1) = phiCenter[block->centerNodes[LocalCenterId] - centerPool.data()];
          }
      for (int k = 0; k < NZ; k++)
        for (int j = 0; j < NY; j++)
          for (int i = 0; i < NX; i++) {
            long int ptr = FirstCellParticleTable[i + NX * (j + NY * k)];
            if (ptr == -1) continue;
            double xInit[3] = {0.0, 0.0, 0.0};
            int spec;
            double xNode[3], xCell[3];
            int index[3] = {i, j, k};
            for (int iDim = 0; iDim < 3; iDim++) {
              xNode[iDim] = node->xmin[iDim] + dx[iDim] * index[iDim];
              xCell[iDim] = node->xmin[iDim] + dx[iDim] * (index[iDim] + 0.5);
            }
            bool atBoundary = isBoundaryCellNonZero(xCell, dx, node);
            long int ptrNext = ptr;
            while (ptrNext != -1) {
              ptr = ptrNext;
              byte *ParticleData = GetParticleDataPointer(ptr);
              spec = GetI(ParticleData);
              GetX(xInit, ParticleData);
              ptrNext = GetNext(ParticleData);
              double xFinal[3];
              if (spec != 0 || atBoundary) {
                for (int idim = 0; idim < 3; idim++) xFinal[idim] = xInit[idim];
              } else {
                double xRel[3];
                for (int iDim = 0; iDim < 3; iDim++) xRel[iDim] = (xInit[iDim] - xNode[iDim]) / dx[iDim];
                int iClosestNode[3];
                for (int iDim = 0; iDim < 3; iDim++) iClosestNode[iDim] = int(index[iDim] + round(xRel[iDim]));
                int ix = iClosestNode[0] + 1, iy = iClosestNode[1] + 1, iz = iClosestNode[2] + 1;
                for (int iDim = 0; iDim < 3; iDim++) xRel[iDim] = xRel[iDim] >= 0.5 ? xRel[iDim] - 0.5 : xRel[iDim] + 0.5;
                double GradPhi[3];
                GradPhi[0] = interp2D(PHI(ix, iy - 1, iz - 1) - PHI(ix - 1, iy - 1, iz - 1), PHI(ix, iy, iz - 1) - PHI(ix - 1, iy, iz - 1),
                                      PHI(ix, iy, iz) - PHI(ix - 1, iy, iz), PHI(ix, iy - 1, iz) - PHI(ix - 1, iy - 1, iz), xRel[1], xRel[2]);
                GradPhi[1] = interp2D(PHI(ix - 1, iy, iz - 1) - PHI(ix - 1, iy - 1, iz - 1), PHI(ix, iy, iz - 1) - PHI(ix, iy - 1, iz - 1),
                                      PHI(ix, iy, iz) - PHI(ix, iy - 1, iz), PHI(ix - 1, iy, iz) - PHI(ix - 1, iy - 1, iz), xRel[0], xRel[2]);
                GradPhi[2] = interp2D(PHI(ix - 1, iy - 1, iz) - PHI(ix - 1, iy - 1, iz - 1), PHI(ix, iy - 1, iz) - PHI(ix, iy - 1, iz - 1),
                                      PHI(ix, iy, iz) - PHI(ix, iy, iz - 1), PHI(ix - 1, iy, iz) - PHI(ix - 1, iy, iz - 1), xRel[0], xRel[1]);
                for (int iDim = 0; iDim < 3; iDim++) GradPhi[iDim] /= dx[iDim];
                double eChargeDens, eps = 0.9;
                int localCornerId = _getCornerNodeLocalNumber(iClosestNode[0], iClosestNode[1], iClosestNode[2]);
                eChargeDens = cornerSpec[(size_t)(block->cornerNodes[localCornerId] - cornerPool.data()) * 10 * nS + 0] * qom[0];
                double displacement[3], temp;
                if (eChargeDens != 0) temp = 1. / (4. * Pi * eChargeDens);
                else temp = 0;
                for (int iDim = 0; iDim < 3; iDim++) displacement[iDim] = -eps * GradPhi[iDim] * temp;
                double epsLimit = 0.1;
                if (fabs(displacement[0] / dx[0]) > epsLimit || fabs(displacement[1] / dx[1]) > epsLimit || fabs(displacement[2] / dx[2]) > epsLimit) {
                  double dl = sqrt(pow(displacement[0], 2) + pow(displacement[1], 2) + pow(displacement[2], 2));
                  for (int iDim = 0; iDim < 3; iDim++) displacement[iDim] *= epsLimit * dx[0] / dl;
                }
                for (int iDim = 0; iDim < 3; iDim++) xFinal[iDim] = xInit[iDim] + displacement[iDim];
                (*nDisplaced)++;
              }
              int ip, jp, kp;
              cTreeNode *newNode = findTreeNode(xFinal, node);
              if (newNode == NULL) {
                DeleteParticle(ptr);
                (*nDeleted)++;
              } else if (newNode->block == NULL) {
                DeleteParticle(ptr);
                (*nDeleted)++;
              } else {
                if (FindCellIndex(xFinal, ip, jp, kp, newNode) == -1) return AMPS_GPU_ERR_PARTICLE;
                AttachToTempList(ptr, ParticleData, newNode->block, ip, jp, kp, 1, 0);
                SetX(xFinal, ParticleData);
              }
            }
          }
    }
    // exchangeParticleLocal (:4366-4438, _COMPILATION_MODE__MPI_): temp lists become the cell lists of EVERY block
    const int nC = nCellsBlock();
    for (size_t l = 0; l < blocks.size(); l++) {
      cBlock *block = &blocks[l];
      // (a periodic ghost block's own list was not walked above; the reference overwrites it with the temp list all the same, so
      //  a particle still filed there - none after a MoveParticles - is no longer on any list)
      memcpy(block->FirstCellParticleTable, block->tempParticleMovingListTable, nC * sizeof(long int));
      for (int c = 0; c < nC; c++) block->tempParticleMovingListTable[c] = -1;
    }
    return AMPS_GPU_OK;
  }

  // ECSIM::ComputeNetCharge, src/pic/pic_field_solver_ecsim.cpp:4690-4828 (without UpdateOldNetCharge): the charge of every particle
  // goes to the 8 cell centres of its trilinear stencil, block by block through a local q_Center array; the sum of the ghost
  // copies into the real nodes (ProcessBlockBoundaryNodes / ProcessNetCharge :1417-1425) is implicit in the unique centre nodes.
  int ComputeNetCharge(double charge_conv) {
    for (int i = 0; i < n_centers; i++) centerPool[i].data[netChargeNew_d] = 0.0;  // SetCenterNodeAssociatedDataValue
    double q_I[AMPS_GPU_MAX_SPECIES];
    for (int iSp = 0; iSp < cfg.n_species; iSp++) q_I[iSp] = cfg.charge[iSp] * charge_conv;
    std::vector<double> q_Center((size_t)nCenterLocal());
    for (size_t nLocalNode = 0; nLocalNode < blocks.size(); nLocalNode++) {
      cTreeNode *node = BlockTable[nLocalNode];
      if (node->block == NULL) continue;
      if (cfg.periodic && node->faceBoundary != 0) continue;  // boundary "ghost" block
      cBlock *block = node->block;
      int nCell[3] = {_BLOCK_CELLS_X_, _BLOCK_CELLS_Y_, _BLOCK_CELLS_Z_};
      long int *FirstCellParticleTable = block->FirstCellParticleTable;
      double CellVolume = 1;
      double dx[3];
      for (int iDim = 0; iDim < 3; iDim++) dx[iDim] = (node->xmax[iDim] - node->xmin[iDim]) / nCell[iDim];
      for (int iDim = 0; iDim < 3; iDim++) CellVolume *= dx[iDim];
      for (int k = -1; k < _BLOCK_CELLS_Z_ + 1; k++)
        for (int j = -1; j < _BLOCK_CELLS_Y_ + 1; j++)
          for (int i = -1; i < _BLOCK_CELLS_X_ + 1; i++) {
            int LocalCenterId = _getCenterNodeLocalNumber(i, j, k);
            if (!block->centerNodes[LocalCenterId]) continue;
            q_Center[LocalCenterId] = 0.0;
          }
      for (int k = 0; k < _BLOCK_CELLS_Z_; k++)
        for (int j = 0; j < _BLOCK_CELLS_Y_; j++)
          for (int i = 0; i < _BLOCK_CELLS_X_; i++) {
            long int ptr = FirstCellParticleTable[i + _BLOCK_CELLS_X_ * (j + _BLOCK_CELLS_Y_ * k)];
            while (ptr != -1) {
              byte *ParticleData = GetParticleDataPointer(ptr);
              double xInit[3];
              int spec = GetI(ParticleData);
              GetX(xInit, ParticleData);
              double LocalParticleWeight = cfg.species_weight[spec];
              LocalParticleWeight *= GetIndividualStatWeightCorrection(ParticleData);
              double chargeQ = q_I[spec] * LocalParticleWeight;
              cStencil NetChargeStencil;
              CellCentered_Linear_InitStencil(xInit, node, NetChargeStencil, false);  // the StencilTable overload: Normalize() only if Length != 8
              globalStencilLength = NetChargeStencil.Length;  // PIC::InterpolationRoutines::CellCentered::StencilTable[thread] keeps it
              for (int iStencil = 0; iStencil < NetChargeStencil.Length; iStencil++)
                q_Center[NetChargeStencil.LocalCellID[iStencil]] += NetChargeStencil.Weight[iStencil] * chargeQ;
              ptr = GetNext(ParticleData);
            }
          }
      for (int k = -1; k < _BLOCK_CELLS_Z_ + 1; k++)
        for (int j = -1; j < _BLOCK_CELLS_Y_ + 1; j++)
          for (int i = -1; i < _BLOCK_CELLS_X_ + 1; i++) {
            int LocalCenterId = _getCenterNodeLocalNumber(i, j, k);
            if (!block->centerNodes[LocalCenterId]) continue;
            block->centerNodes[LocalCenterId]->data[netChargeNew_d] += q_Center[LocalCenterId] / CellVolume;
          }
    }
    return AMPS_GPU_OK;
  }

  // PIC::Mover::cExternalBoundaryFace + Init, src/pic/pic_mover.cpp:24-28,48-75
  struct cExternalBoundaryFace {
    double norm[3];
    int nX0[3];
    double e0[3], e1[3], x0[3];
    double lE0, lE1;
  } FaceTable[6];
  void InitFaceTable() {
    static const cExternalBoundaryFace init[6] = {
        {{-1.0, 0.0, 0.0}, {0, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}, 0.0, 0.0}, {{1.0, 0.0, 0.0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}, 0.0, 0.0},
        {{0.0, -1.0, 0.0}, {0, 0, 0}, {1, 0, 0}, {0, 0, 1}, {0, 0, 0}, 0.0, 0.0}, {{0.0, 1.0, 0.0}, {0, 1, 0}, {1, 0, 0}, {0, 0, 1}, {0, 0, 0}, 0.0, 0.0},
        {{0.0, 0.0, -1.0}, {0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 0}, 0.0, 0.0}, {{0.0, 0.0, 1.0}, {0, 0, 1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 0}, 0.0, 0.0}};
    for (int nface = 0; nface < 6; nface++) {
      FaceTable[nface] = init[nface];
      double cE0 = 0.0, cE1 = 0.0;
      for (int idim = 0; idim < 3; idim++) {
        FaceTable[nface].x0[idim] = (FaceTable[nface].nX0[idim] == 0) ? xGlobalMin[idim] : xGlobalMax[idim];
        cE0 += pow(((FaceTable[nface].e0[idim] + FaceTable[nface].nX0[idim] < 0.5) ? xGlobalMin[idim] : xGlobalMax[idim]) - FaceTable[nface].x0[idim], 2);
        cE1 += pow(((FaceTable[nface].e1[idim] + FaceTable[nface].nX0[idim] < 0.5) ? xGlobalMin[idim] : xGlobalMax[idim]) - FaceTable[nface].x0[idim], 2);
      }
      FaceTable[nface].lE0 = sqrt(cE0);
      FaceTable[nface].lE1 = sqrt(cE1);
    }
  }

  // PIC::BC::ExternalBoundary::Periodic::ExchangeParticlesLocal, src/pic/pic_bc_periodic.cpp:86-178
  long int ExchangeParticlesLocal(cTreeNode *RealBlock, cTreeNode *GhostBlock) {
    int i, j, k, idim;
    long int ptr, NextPtr, nMoved = 0;
    double dx[3];
    for (i = 0; i < 3; i++) dx[i] = RealBlock->xmin[i] - GhostBlock->xmin[i];
    for (int icell = 0; icell < nCellsBlock(); icell++) {
      int t, ii = icell;
      double x[3];
      t = _BLOCK_CELLS_X_ * _BLOCK_CELLS_Y_;
      k = ii / t;
      ii = ii % t;
      j = ii / _BLOCK_CELLS_X_;
      i = ii % _BLOCK_CELLS_X_;
      int c = i + _BLOCK_CELLS_X_ * (j + _BLOCK_CELLS_Y_ * k);
      if ((ptr = GhostBlock->block->FirstCellParticleTable[c]) != -1) {
        NextPtr = GetNext(ptr);
        byte *p = GetParticleDataPointer(ptr);
        GetX(x, p);
        for (idim = 0; idim < 3; idim++) {
          x[idim] += dx[idim];
          if (x[idim] < RealBlock->xmin[idim]) x[idim] = RealBlock->xmin[idim];
          if (x[idim] >= RealBlock->xmax[idim]) x[idim] = RealBlock->xmax[idim] - 1.0E-10 * (RealBlock->xmax[idim] - RealBlock->xmin[idim]);
        }
        SetX(x, p);
        nMoved++;
        if (NextPtr != -1) {
          do {
            ptr = NextPtr;
            p = GetParticleDataPointer(ptr);
            GetX(x, p);
            for (idim = 0; idim < 3; idim++) {
              x[idim] += dx[idim];
              if (x[idim] < RealBlock->xmin[idim]) x[idim] = RealBlock->xmin[idim];
              if (x[idim] >= RealBlock->xmax[idim]) x[idim] = RealBlock->xmax[idim] - 1.0E-10 * (RealBlock->xmax[idim] - RealBlock->xmin[idim]);
            }
            SetX(x, p);
            nMoved++;
            NextPtr = GetNext(ptr);
          } while (NextPtr != -1);
        }
        SetNext(RealBlock->block->FirstCellParticleTable[c], ptr);
        if (RealBlock->block->FirstCellParticleTable[c] != -1) {
          SetPrev(ptr, RealBlock->block->FirstCellParticleTable[c]);
        }
        RealBlock->block->FirstCellParticleTable[c] = GhostBlock->block->FirstCellParticleTable[c];
        GhostBlock->block->FirstCellParticleTable[c] = -1;
      }
    }
    return nMoved;
  }

  // ------------------------------------------------------------------------------------------
  // ECSIM_AddGuidingCenterMagnetizationCurrentToCorners, src/pic/pic_field_solver_ecsim.cpp:1828-1875: J_c += sum_a grad N_a(x_c^-) x M_a
  // with M_a = MCornerSum[a] / CellVolume and the one-sided gradients of the trilinear basis at the corner
  static void AddGuidingCenterMagnetizationCurrentToCorners(cCellData *CellData, const double MCornerSum[8][3], double CellVolume, const double dx[3]) {
    const int CornerBits[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    double Mnode[8][3];
    for (int iCorner = 0; iCorner < 8; iCorner++)
      for (int d = 0; d < 3; d++) Mnode[iCorner][d] = MCornerSum[iCorner][d] / CellVolume;
    for (int iTargetCorner = 0; iTargetCorner < 8; iTargetCorner++) {
      const int uc = CornerBits[iTargetCorner][0], vc = CornerBits[iTargetCorner][1], wc = CornerBits[iTargetCorner][2];
      double Jcorner[3] = {0.0, 0.0, 0.0};
      for (int iSourceCorner = 0; iSourceCorner < 8; iSourceCorner++) {
        const int ua = CornerBits[iSourceCorner][0], va = CornerBits[iSourceCorner][1], wa = CornerBits[iSourceCorner][2];
        double dNdx = 0.0, dNdy = 0.0, dNdz = 0.0;
        if ((va == vc) && (wa == wc)) dNdx = ((ua == 1) ? 1.0 : -1.0) / dx[0];
        if ((ua == uc) && (wa == wc)) dNdy = ((va == 1) ? 1.0 : -1.0) / dx[1];
        if ((ua == uc) && (va == vc)) dNdz = ((wa == 1) ? 1.0 : -1.0) / dx[2];
        const double *M = Mnode[iSourceCorner];
        Jcorner[0] += dNdy * M[2] - dNdz * M[1];
        Jcorner[1] += dNdz * M[0] - dNdx * M[2];
        Jcorner[2] += dNdx * M[1] - dNdy * M[0];
      }
      double *CornerJ = CellData->CornerData[iTargetCorner].CornerJ;
      CornerJ[0] += Jcorner[0];
      CornerJ[1] += Jcorner[1];
      CornerJ[2] += Jcorner[2];
    }
  }

  // ECSIM::ProcessCell, src/pic/pic_field_solver_ecsim.cpp:1881-2438 (scalar branch, B centre or
  // corner based, guiding-centre species of PIC::GYROKINETIC per cfg.gc_species_mask (:2084, :2205-2256, :2310, :2376), no
  // per-species corner sampling)
  // ------------------------------------------------------------------------------------------
  bool ProcessCell(int iCellIn, int jCellIn, int kCellIn, cTreeNode *node, cCellData *CellData, double *MassTable, double *ChargeTable) {
    std::vector<double *> &B_Center = tlsBCenter();
    bool res = false;
    const double dtTotal = cfg.ecsim_dt_total, B_conv = cfg.ecsim_B_conv, length_conv = cfg.ecsim_length_conv, LightSpeed = cfg.ecsim_light_speed;
    const int nSpecies = cfg.n_species;

    B_Center.assign((cfg.b_mode == AMPS_B_CENTER_BASED) ? nCenterLocal() : nCornerLocal(), NULL);
    if (cfg.b_mode == AMPS_B_CENTER_BASED) {
      for (int k = kCellIn - 1; k <= kCellIn + 1; k++)
        for (int j = jCellIn - 1; j <= jCellIn + 1; j++)
          for (int i = iCellIn - 1; i <= iCellIn + 1; i++) {
            int LocalCenterId = _getCenterNodeLocalNumber(i, j, k);
            if (!node->block->centerNodes[LocalCenterId]) continue;
            B_Center[LocalCenterId] = node->block->centerNodes[LocalCenterId]->data + CurrentBOffset_d;
          }
    }

    int nCell[3] = {_BLOCK_CELLS_X_, _BLOCK_CELLS_Y_, _BLOCK_CELLS_Z_};
    cBlock *block = node->block;
    long int *FirstCellParticleTable = block->FirstCellParticleTable;
    double CellVolume = 1;
    double dx[3];
    double GlobalTimeStep = cfg.time_step[0];

    for (int iDim = 0; iDim < 3; iDim++) dx[iDim] = (node->xmax[iDim] - node->xmin[iDim]) / nCell[iDim] * length_conv;
    for (int iDim = 0; iDim < 3; iDim++) CellVolume *= dx[iDim];

    long int ptr = FirstCellParticleTable[iCellIn + _BLOCK_CELLS_X_ * (jCellIn + _BLOCK_CELLS_Y_ * kCellIn)];
    double ParticleEnergyCell = 0, vmean_cell[AMPS_GPU_MAX_SPECIES];
    for (int iSp = 0; iSp < nSpecies; iSp++) vmean_cell[iSp] = 0.0;

    if (ptr != -1) {
      res = true;
      double vInit[3] = {0.0, 0.0, 0.0}, xInit[3] = {0.0, 0.0, 0.0};
      int spec;
      double Jg[8][3];
      for (int ii = 0; ii < 8; ii++)
        for (int jj = 0; jj < 3; jj++) Jg[ii][jj] = 0.0;

      double MCornerSum[8][3];  // guiding-centre magnetisation closure (:1988-1996)
      bool cell_has_gc = false;
      for (int ii = 0; ii < 8; ii++)
        for (int jj = 0; jj < 3; jj++) MCornerSum[ii][jj] = 0.0;

      double MassMatrix_GGD[8][8][9];
      for (int iCorner = 0; iCorner < 8; iCorner++)
        for (int jCorner = 0; jCorner < 8; jCorner++)
          for (int idim = 0; idim < 9; idim++) MassMatrix_GGD[iCorner][jCorner][idim] = 0.0;

      long int ptrNext = ptr;
      byte *ParticleData, *ParticleDataNext;
      ParticleDataNext = GetParticleDataPointer(ptr);

      static const int cornerOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
      for (int ii = 0; ii < 8; ii++) {
        cCornerNode *cn = block->cornerNodes[_getCornerNodeLocalNumber(iCellIn + cornerOff[ii][0], jCellIn + cornerOff[ii][1], kCellIn + cornerOff[ii][2])];
        CellData->CornerData[ii].CornerNode = cn;
        CellData->CornerData[ii].CornerMassMatrix_ptr = cn->data + MassMatrixOffsetIndex;
        CellData->CornerData[ii].CornerJ_ptr = cn->data + JxOffsetIndex;
      }

      int cnt = 0, particleNumber[AMPS_GPU_MAX_SPECIES];
      for (int iSp = 0; iSp < nSpecies; iSp++) particleNumber[iSp] = 0;

      cStencil MagneticFieldStencil, CornerBasedStencil;

      while (ptrNext != -1) {
        double LocalParticleWeight;
        ptr = ptrNext;
        ParticleData = ParticleDataNext;

        spec = GetI(ParticleData);
        GetV(vInit, ParticleData);
        GetX(xInit, ParticleData);
        LocalParticleWeight = cfg.species_weight[spec];
        LocalParticleWeight *= GetIndividualStatWeightCorrection(ParticleData);
        const bool use_gc_species = ((cfg.gc_species_mask >> spec) & 1) != 0;  // _PIC_GYROKINETIC_MODEL_MODE_ on && IsGuidingCenterSpecies (:2084)

        ptrNext = GetNext(ParticleData);
        if (ptrNext != -1) ParticleDataNext = GetParticleDataPointer(ptrNext);

        {
          double B[3] = {0.0, 0.0, 0.0};
          double Wdummy[8];
          if (cfg.b_mode == AMPS_B_CENTER_BASED)
            CellCentered_Linear_InitStencil(xInit, node, MagneticFieldStencil, globalStencilLength != 8);
          else
            CornerBased_InitStencil(xInit, node, MagneticFieldStencil, Wdummy);

          int Length = MagneticFieldStencil.Length;
          double *Weight_table = MagneticFieldStencil.Weight;
          int *LocalCellID_table = MagneticFieldStencil.LocalCellID;

          for (int iStencil = 0; iStencil < Length; iStencil++) {
            double *B_temp, Weight = Weight_table[iStencil];
            int LocalCellID = LocalCellID_table[iStencil];
            if (cfg.b_mode == AMPS_B_CENTER_BASED)
              B_temp = B_Center[LocalCellID];
            else
              B_temp = block->cornerNodes[LocalCellID]->data + OffsetB_corner_d + CurrentBOffset_d;  // B_corner[LocalCellID] (:1932-1946)
            for (int idim = 0; idim < 3; idim++) B[idim] += Weight * B_temp[idim];
          }

          for (int idim = 0; idim < 3; idim++) {
            B[idim] *= B_conv;
            vInit[idim] *= length_conv;
          }
          const double Braw[3] = {B[0], B[1], B[2]};  // before the /LightSpeed scaling (:2141)

          double QdT_over_m, QdT_over_2m, alpha[9], chargeQ;
          double WeightPG[8];
          double c0, QdT_over_2m_squared;
          double mass;

          chargeQ = ChargeTable[spec];
          mass = MassTable[spec];
          chargeQ *= LocalParticleWeight;
          mass *= LocalParticleWeight;

          QdT_over_m = chargeQ * dtTotal / mass;
          QdT_over_2m = 0.5 * QdT_over_m;
          QdT_over_2m_squared = QdT_over_2m * QdT_over_2m;

          for (int idim = 0; idim < 3; idim++) B[idim] /= LightSpeed;

          double BB[3][3], P[3];
          for (int ii = 0; ii < 3; ii++) {
            P[ii] = -QdT_over_2m * B[ii];
            for (int jj = 0; jj <= ii; jj++) {
              BB[ii][jj] = QdT_over_2m_squared * B[ii] * B[jj];
              BB[jj][ii] = BB[ii][jj];
            }
          }
          c0 = 1.0 / (1.0 + QdT_over_2m_squared * (B[0] * B[0] + B[1] * B[1] + B[2] * B[2]));

          alpha[0] = c0 * (1.0 + BB[0][0]);
          alpha[1] = c0 * (-P[2] + BB[0][1]);
          alpha[2] = c0 * (P[1] + BB[0][2]);
          alpha[3] = c0 * (P[2] + BB[1][0]);
          alpha[4] = c0 * (1.0 + BB[1][1]);
          alpha[5] = c0 * (-P[0] + BB[1][2]);
          alpha[6] = c0 * (-P[1] + BB[2][0]);
          alpha[7] = c0 * (P[0] + BB[2][1]);
          alpha[8] = c0 * (1.0 + BB[2][2]);

          CornerBased_InitStencil(xInit, node, CornerBasedStencil, WeightPG);

          if (use_gc_species) {  // mu b to the corners (:2205-2225)
            const double mu = GetMagneticMoment(ParticleData);
            const double mu_tot = mu * LocalParticleWeight;
            const double absB = sqrt(Braw[0] * Braw[0] + Braw[1] * Braw[1] + Braw[2] * Braw[2]);
            if (absB > 0.0) {
              const double b0 = Braw[0] / absB, b1 = Braw[1] / absB, b2 = Braw[2] / absB;
              for (int iCorner = 0; iCorner < 8; iCorner++) {
                const double w = mu_tot * WeightPG[iCorner];
                MCornerSum[iCorner][0] += w * b0;
                MCornerSum[iCorner][1] += w * b1;
                MCornerSum[iCorner][2] += w * b2;
              }
              cell_has_gc = true;
            }
          }

          double vsqr = vInit[0] * vInit[0] + vInit[1] * vInit[1] + vInit[2] * vInit[2];
          if (use_gc_species) {  // the unresolved gyration carried by v_normal (:2232-2235)
            const double vperp = GetVNormal(ParticleData);
            vsqr += vperp * vperp;
          }
          vmean_cell[spec] += sqrt(vsqr) * GlobalTimeStep;
          ParticleEnergyCell += 0.5 * mass * vsqr;

          double vRot[3] = {0.0, 0.0, 0.0};
          for (int iDim = 0; iDim < 3; iDim++)
            for (int jj = 0; jj < 3; jj++) vRot[iDim] += alpha[3 * iDim + jj] * vInit[jj];
          if (use_gc_species) vRot[0] = vInit[0], vRot[1] = vInit[1], vRot[2] = vInit[2];  // v_eff is deposited as it is (:2253-2257)

          for (int iCorner = 0; iCorner < 8; iCorner++) {
            double t = chargeQ * WeightPG[iCorner];
            double *Jg_iCorner = Jg[iCorner];
            for (int iDim = 0; iDim < 3; iDim++) Jg_iCorner[iDim] += t * vRot[iDim];
          }

          double matrixConst = chargeQ * QdT_over_2m / CellVolume;
          if (!use_gc_species)  // the implicit response is for full-orbit species only (:2310)
          for (int iCorner = 0; iCorner < 8; iCorner++) {
            double tempWeightConst = matrixConst * WeightPG[iCorner];
            for (int jCorner = 0; jCorner <= iCorner; jCorner++) {
              double tempWeightProduct = WeightPG[jCorner] * tempWeightConst;
              double *tmpPtr = MassMatrix_GGD[iCorner][jCorner];
              tmpPtr[0] += alpha[0] * tempWeightProduct;
              tmpPtr[1] += alpha[1] * tempWeightProduct;
              tmpPtr[2] += alpha[2] * tempWeightProduct;
              tmpPtr[3] += alpha[3] * tempWeightProduct;
              tmpPtr[4] += alpha[4] * tempWeightProduct;
              tmpPtr[5] += alpha[5] * tempWeightProduct;
              tmpPtr[6] += alpha[6] * tempWeightProduct;
              tmpPtr[7] += alpha[7] * tempWeightProduct;
              tmpPtr[8] += alpha[8] * tempWeightProduct;
            }
          }
          particleNumber[spec]++;
        }
        cnt++;

        if (ptrNext == -1) {
          CellData->ParticleEnergy += ParticleEnergyCell;
          for (int iSp = 0; iSp < nSpecies; iSp++) {
            CellData->cflCell[iSp] = vmean_cell[iSp] / (particleNumber[iSp] * sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]));
          }
          for (int iCorner = 0; iCorner < 8; iCorner++) {
            double *CornerJ = CellData->CornerData[iCorner].CornerJ;
            for (int ii = 0; ii < 3; ii++) CornerJ[ii] += (Jg[iCorner][ii]) / CellVolume;
          }
          if (cell_has_gc) AddGuidingCenterMagnetizationCurrentToCorners(CellData, MCornerSum, CellVolume, dx);  // :2376
          for (int iCorner = 0; iCorner < 8; iCorner++) {
            for (int jCorner = 0; jCorner <= iCorner; jCorner++) {
              if (iCorner == jCorner) {
                double *CornerMassMatrix = CellData->CornerData[iCorner].CornerMassMatrix;
                for (int ii = 0; ii < 3; ii++)
                  for (int jj = 0; jj < 3; jj++) CornerMassMatrix[3 * ii + jj] += MassMatrix_GGD[iCorner][iCorner][3 * ii + jj];
              } else {
                double *CornerMassMatrix_iCorner = CellData->CornerData[iCorner].CornerMassMatrix;
                double *CornerMassMatrix_jCorner = CellData->CornerData[jCorner].CornerMassMatrix;
                for (int ii = 0; ii < 3; ii++)
                  for (int jj = 0; jj < 3; jj++) {
                    CornerMassMatrix_iCorner[9 * IndexMatrix[iCorner][jCorner] + 3 * ii + jj] += MassMatrix_GGD[iCorner][jCorner][3 * ii + jj];
                    CornerMassMatrix_jCorner[9 * IndexMatrix[jCorner][iCorner] + 3 * ii + jj] += MassMatrix_GGD[iCorner][jCorner][3 * ii + jj];
                  }
              }
            }
          }
        }
      }
    }
    return res;
  }
  static std::vector<double *> &tlsBCenter() {
    static thread_local std::vector<double *> v;
    return v;
  }
};

// =============================================================================================
// C API
// =============================================================================================
extern "C" {

oracle_ctx *oracle_create(const amps_gpu_config *cfg, const amps_gpu_mesh *m) {
  oracle_ctx *o = new oracle_ctx();
  o->cfg = *cfg;
  o->_BLOCK_CELLS_X_ = cfg->block_cells[0], o->_BLOCK_CELLS_Y_ = cfg->block_cells[1], o->_BLOCK_CELLS_Z_ = cfg->block_cells[2];
  o->_GHOST_CELLS_X_ = cfg->ghost_cells[0], o->_GHOST_CELLS_Y_ = cfg->ghost_cells[1], o->_GHOST_CELLS_Z_ = cfg->ghost_cells[2];
  o->_TOTAL_BLOCK_CELLS_X_ = o->_BLOCK_CELLS_X_ + 2 * o->_GHOST_CELLS_X_;
  o->_TOTAL_BLOCK_CELLS_Y_ = o->_BLOCK_CELLS_Y_ + 2 * o->_GHOST_CELLS_Y_;
  o->_TOTAL_BLOCK_CELLS_Z_ = o->_BLOCK_CELLS_Z_ + 2 * o->_GHOST_CELLS_Z_;
  for (int d = 0; d < 3; d++) {
    o->xGlobalMin[d] = m->x_global_min[d], o->xGlobalMax[d] = m->x_global_max[d];
    o->dx_max_refinment[d] = m->dx_max_refinement[d], o->dxRootBlock[d] = m->dx_root_block[d];
    o->nRoot[d] = m->n_root[d];
  }
  o->EPS = m->eps;
  o->maxRefinementLevel = m->max_refinement_level;
  o->InitFaceTable();

  // unique node pools
  o->n_corners = m->n_corners, o->n_centers = m->n_centers;
  const int cornerLen = CornerDataLength + 6;  // + corner B_cur, B_prev
  o->cornerData.assign((size_t)m->n_corners * cornerLen, 0.0);
  o->centerData.assign((size_t)m->n_centers * CenterDataLength, 0.0);
  o->cornerPool = std::vector<cCornerNode>(m->n_corners);
  o->centerPool = std::vector<cCenterNode>(m->n_centers);
  for (int i = 0; i < m->n_corners; i++) o->cornerPool[i].data = o->cornerData.data() + (size_t)i * cornerLen;
  for (int i = 0; i < m->n_centers; i++) o->centerPool[i].data = o->centerData.data() + (size_t)i * CenterDataLength;

  // tree
  o->nodes.resize(m->n_nodes);
  o->blocks.resize(m->n_leaves);
  o->BlockTable.assign(m->n_leaves, NULL);
  for (int n = 0; n < m->n_nodes; n++) {
    cTreeNode &t = o->nodes[n];
    t.id = n;
    for (int d = 0; d < 3; d++) {
      t.xmin[d] = m->node_xmin[3 * n + d], t.xmax[d] = m->node_xmax[3 * n + d];
      t.xMinGlobalIndex[d] = m->node_imin[3 * n + d];
    }
    t.NodeGeometricSizeIndex = m->node_isize[n];
    for (int c = 0; c < 8; c++) t.downNode[c] = (m->node_child[8 * n + c] >= 0) ? &o->nodes[m->node_child[8 * n + c]] : NULL;
    t.upNode = (m->node_parent[n] >= 0) ? &o->nodes[m->node_parent[n]] : NULL;
    t.RefinmentLevel = m->node_level[n];
    t.Thread = m->node_thread[n];
    t.IsUsedInCalculationFlag = (m->node_flags[n] & AMPS_NODE_USED) != 0;
    t.IsGhostNodeFlag = (m->node_flags[n] & AMPS_NODE_PERIODIC_GHOST) != 0;
    t.leaf = m->node_leaf[n];
    t.block = NULL;
    t.faceBoundary = 0;
  }
  int nr = m->n_root[0] * m->n_root[1] * m->n_root[2];
  o->rootTable.resize(nr);
  for (int i = 0; i < nr; i++) o->rootTable[i] = &o->nodes[m->root_node[i]];

  const int nC = o->nCellsBlock(), nCor = o->nCornerLocal(), nCen = o->nCenterLocal();
  o->listStorage.assign((size_t)m->n_leaves * nC * 2, -1);
  o->leaf_real.assign(m->leaf_real, m->leaf_real + m->n_leaves);
  for (int l = 0; l < m->n_leaves; l++) {
    cTreeNode *t = &o->nodes[m->leaf_node[l]];
    cBlock &b = o->blocks[l];
    t->block = &b;
    t->faceBoundary = m->leaf_face_boundary[l];
    o->BlockTable[l] = t;
    b.FirstCellParticleTable = o->listStorage.data() + (size_t)l * nC * 2;
    b.tempParticleMovingListTable = b.FirstCellParticleTable + nC;
    b.tempThreadTable = NULL;
    b.cornerNodes = new cCornerNode *[nCor];
    b.centerNodes = new cCenterNode *[nCen];
    for (int i = 0; i < nCor; i++) {
      int uid = m->leaf_corner_uid[(size_t)l * nCor + i];
      b.cornerNodes[i] = (uid >= 0) ? &o->cornerPool[uid] : NULL;
    }
    for (int i = 0; i < nCen; i++) {
      int uid = m->leaf_center_uid[(size_t)l * nCen + i];
      b.centerNodes[i] = (uid >= 0) ? &o->centerPool[uid] : NULL;
    }
  }
  o->nThreadListTables = 0;
  for (int n = 0; n < m->n_nodes; n++) o->SetNeibRefinmentLevelLimits(&o->nodes[n]);

  // PIC::ParticleBuffer::Init, src/pic/pic_pbuffer.cpp:41-222
  o->MaxNPart = cfg->capacity;
  o->ParticleDataLength = oracle_ctx::BASIC_LEN;
  o->ParticleDataBuffer = (byte *)malloc((size_t)o->ParticleDataLength * o->MaxNPart);
  memset(o->ParticleDataBuffer, 0, (size_t)o->ParticleDataLength * o->MaxNPart);
  for (long int ptr = 0; ptr < o->MaxNPart - 1; ptr++) {
    o->SetNext(ptr + 1, ptr);
    oracle_ctx::SetParticleDeleted(o->GetParticleDataPointer(ptr));
  }
  o->SetNext(-1, o->MaxNPart - 1);
  o->FirstPBufferParticle = 0;
  o->NAllPart = 0;
  return o;
}

void oracle_destroy(oracle_ctx *o) {
  if (!o) return;
  for (auto &b : o->blocks) {
    delete[] b.cornerNodes;
    delete[] b.centerNodes;
  }
  free(o->ParticleDataBuffer);
  delete o;
}

const char *oracle_last_error(const oracle_ctx *o) { return o->err.c_str(); }
int64_t oracle_particle_data_length(const oracle_ctx *o) { return o->ParticleDataLength; }
int64_t oracle_particle_count(const oracle_ctx *o) { return o->NAllPart; }

void oracle_set_fields(oracle_ctx *o, const double *E_half, const double *B_prev, const double *B_cur) {
  if (E_half)
    for (int i = 0; i < o->n_corners; i++) memcpy(o->cornerPool[i].data + OffsetE_HalfTimeStep_d, E_half + 3 * (size_t)i, 24);
  if (o->cfg.b_mode == AMPS_B_CORNER_BASED) {  // B_prev, B_cur are [n_corners][3]
    if (B_prev)
      for (int i = 0; i < o->n_corners; i++) memcpy(o->cornerPool[i].data + OffsetB_corner_d + PrevBOffset_d, B_prev + 3 * (size_t)i, 24);
    if (B_cur)
      for (int i = 0; i < o->n_corners; i++) memcpy(o->cornerPool[i].data + OffsetB_corner_d + CurrentBOffset_d, B_cur + 3 * (size_t)i, 24);
    return;
  }
  if (B_prev)
    for (int i = 0; i < o->n_centers; i++) memcpy(o->centerPool[i].data + PrevBOffset_d, B_prev + 3 * (size_t)i, 24);
  if (B_cur)
    for (int i = 0; i < o->n_centers; i++) memcpy(o->centerPool[i].data + CurrentBOffset_d, B_cur + 3 * (size_t)i, 24);
}

void oracle_set_background(oracle_ctx *o, const double *E_center, const double *B_center) {
  if (E_center)
    for (int i = 0; i < o->n_centers; i++) memcpy(o->centerPool[i].data + BackgroundE_d, E_center + 3 * (size_t)i, 24);
  if (B_center)
    for (int i = 0; i < o->n_centers; i++) memcpy(o->centerPool[i].data + BackgroundB_d, B_center + 3 * (size_t)i, 24);
}
// probes of the specfunc.h restatements (pinned against the reference's header in tests/test_reference_mesh.py)
double oracle_probe_gyro_frequency(const double *v, double m, double q, const double *B, double c) { return oracle_ctx::RelGyroFrequency(v, m, q, B, c); }
void oracle_probe_normalize(double *x) { oracle_ctx::Vector3D_Normalize(x); }
int oracle_sample_cells(oracle_ctx *o, double *sample, int64_t *n_sampled) {
  int rc = o->SampleCells();
  if (sample) memcpy(sample, o->cellSample.data(), sizeof(double) * o->cellSample.size());
  if (n_sampled)
    for (int s = 0; s < o->cfg.n_species; s++) n_sampled[s] = o->sampledParticles[s];
  return rc;
}
int oracle_species_moments(oracle_ctx *o, double *spec) {
  int rc = o->ComputeSpeciesMoments();
  if (spec) memcpy(spec, o->cornerSpec.data(), sizeof(double) * o->cornerSpec.size());
  return rc;
}
void oracle_set_phi(oracle_ctx *o, const double *phi_center) { o->phiCenter.assign(phi_center, phi_center + o->n_centers); }
int oracle_correct_particle_location(oracle_ctx *o, double charge_conv, double mass_conv, int32_t *final_cell, int64_t *n_displaced,
                                     int64_t *n_deleted) {
  long long nd = 0, nx = 0;
  int rc = o->CorrectParticleLocation(charge_conv, mass_conv, &nd, &nx);
  if (n_displaced) *n_displaced = nd;
  if (n_deleted) *n_deleted = nx;
  if (final_cell) {
    const int nC = o->nCellsBlock();
    for (long int p = 0; p < o->MaxNPart; p++) final_cell[p] = -1;
    for (size_t l = 0; l < o->blocks.size(); l++)
      for (int c = 0; c < nC; c++)
        for (long int ptr = o->blocks[l].FirstCellParticleTable[c]; ptr != -1; ptr = o->GetNext(ptr)) final_cell[ptr] = (int32_t)(l * nC + c);
  }
  return rc;
}
int oracle_net_charge(oracle_ctx *o, double charge_conv, double *rho) {
  int rc = o->ComputeNetCharge(charge_conv);
  for (int i = 0; i < o->n_centers; i++) rho[i] = o->centerPool[i].data[netChargeNew_d];
  return rc;
}
void oracle_set_background_gca(oracle_ctx *o, const double *var15) {
  for (int i = 0; i < o->n_centers; i++) memcpy(o->centerPool[i].data + BackgroundGCA_d, var15 + 15 * (size_t)i, 15 * 8);
}
void oracle_set_background_gradB(oracle_ctx *o, const double *gradB) {
  for (int i = 0; i < o->n_centers; i++) memcpy(o->centerPool[i].data + BackgroundGradB_d, gradB + 9 * (size_t)i, 9 * 8);
}
// the reduced state of the gyrokinetic movers, by ptr (PB::SetMagneticMoment / SetVParallel at injection)
void oracle_set_reduced_state(oracle_ctx *o, const double *mu, const double *vpar, int64_t n) {
  for (int64_t ptr = 0; ptr < n; ptr++) {
    byte *pd = o->GetParticleDataPointer(ptr);
    if (mu) oracle_ctx::SetMagneticMoment(mu[ptr], pd);
    if (vpar) oracle_ctx::SetVParallel(vpar[ptr], pd);
  }
}
// the current E on the unique corners (slot 0 of the corner data: what ECSIM::GetElectricField reads)
void oracle_set_E_current(oracle_ctx *o, const double *E) {
  for (int i = 0; i < o->n_corners; i++) memcpy(o->cornerPool[i].data + ExOffsetIndex, E + 3 * (size_t)i, 24);
}
// the state of the reference's global StencilTable (its Length): 8 after a ComputeNetCharge of a periodic box, 0 = never filled
void oracle_set_global_stencil_length(oracle_ctx *o, int length) { o->globalStencilLength = length; }
void oracle_set_v_normal(oracle_ctx *o, const double *vnormal, int64_t n) {
  for (int64_t ptr = 0; ptr < n; ptr++) oracle_ctx::SetVNormal(vnormal[ptr], o->GetParticleDataPointer(ptr));
}
void oracle_get_v_parallel(const oracle_ctx *o, double *vpar, int64_t n) {
  for (int64_t ptr = 0; ptr < n; ptr++) vpar[ptr] = oracle_ctx::GetVParallel(o->GetParticleDataPointer(ptr));
}
void oracle_get_magnetic_moment(const oracle_ctx *o, double *mu, uint8_t *init_flag, int64_t n) {
  for (int64_t ptr = 0; ptr < n; ptr++) {
    const byte *pd = o->GetParticleDataPointer(ptr);
    if (mu) mu[ptr] = oracle_ctx::GetMagneticMoment(pd);
    if (init_flag) init_flag[ptr] = oracle_ctx::TestInitFlag(pd) ? 1 : 0;
  }
}
// InitiateMagneticMoment for every particle on the cell lists; mu (by ptr) is returned in mu_out when not NULL
int oracle_magnetic_moment_init(oracle_ctx *o, int mover_id, double *mu_out, int64_t n) {
  const int nC = o->nCellsBlock();
  for (size_t l = 0; l < o->blocks.size(); l++)
    for (int c = 0; c < nC; c++)
      for (long int ptr = o->blocks[l].FirstCellParticleTable[c]; ptr != -1; ptr = o->GetNext(ptr)) {
        byte *pd = o->GetParticleDataPointer(ptr);
        double x[3], v[3];
        oracle_ctx::GetX(x, pd), oracle_ctx::GetV(v, pd);
        if (mover_id == AMPS_MOVER_RELATIVISTIC_GCA) {
          if (!o->RelGCA_InitiateMagneticMoment(oracle_ctx::GetI(pd), x, v, pd, o->BlockTable[l])) return AMPS_GPU_ERR_PARTICLE;
        } else {  // GuidingCenter::InitiateMagneticMoment also aligns v with B (the reference passes GetV(ptr))
          if (!o->GC_InitiateMagneticMoment(oracle_ctx::GetI(pd), x, v, pd, o->BlockTable[l])) return AMPS_GPU_ERR_PARTICLE;
          oracle_ctx::SetV(v, pd);
        }
        if (mu_out && ptr < n) mu_out[ptr] = oracle_ctx::GetMagneticMoment(pd);
      }
  return AMPS_GPU_OK;
}
int64_t oracle_exit_records(oracle_ctx *o, amps_gpu_exit_record *buf, int64_t max_records) {
  int64_t n = (int64_t)o->exitRecords.size();
  if (buf)
    for (int64_t i = 0; i < n && i < max_records; i++) buf[i] = o->exitRecords[i];
  o->exitRecords.clear();
  return n;
}

// PIC::ParticleBuffer::InitiateParticle(..., _PIC_INIT_PARTICLE_MODE__ADD2LIST_), src/pic/pic_pbuffer.cpp:939-1027
int oracle_add_particles(oracle_ctx *o, const double *x, const double *v, const double *w, const uint8_t *species, const int32_t *cells, int64_t n) {
  const int nC = o->nCellsBlock();
  for (int64_t p = 0; p < n; p++) {
    long int ptr = o->GetNewParticle();
    if (ptr < 0) {
      o->err = "particle buffer exhausted";
      return AMPS_GPU_ERR_CAPACITY;
    }
    byte *pd = o->GetParticleDataPointer(ptr);
    double xx[3] = {x[p], x[n + p], x[2 * n + p]}, vv[3] = {v[p], v[n + p], v[2 * n + p]};
    oracle_ctx::SetX(xx, pd);
    oracle_ctx::SetV(vv, pd);
    oracle_ctx::SetI(species[p], pd);
    oracle_ctx::SetIndividualStatWeightCorrection(w ? w[p] : 1.0, pd);
    int leaf = cells[p] / nC, cell = cells[p] % nC;
    long int *first = o->blocks[leaf].FirstCellParticleTable + cell;
    o->SetNext(*first, ptr);
    o->SetPrev(-1, ptr);
    if (*first != -1) o->SetPrev(ptr, *first);
    *first = ptr;
  }
  return AMPS_GPU_OK;
}

void oracle_get_particles(const oracle_ctx *o, double *x, double *v, double *w, uint8_t *species, int32_t *cells, uint8_t *alive, int64_t n) {
  const int nC = o->nCellsBlock();
  if (cells)
    for (int64_t i = 0; i < n; i++) cells[i] = -1;
  for (int64_t ptr = 0; ptr < n; ptr++) {
    const byte *pd = o->GetParticleDataPointer(ptr);
    double t[3];
    if (x) {
      oracle_ctx::GetX(t, pd);
      x[ptr] = t[0], x[n + ptr] = t[1], x[2 * n + ptr] = t[2];
    }
    if (v) {
      oracle_ctx::GetV(t, pd);
      v[ptr] = t[0], v[n + ptr] = t[1], v[2 * n + ptr] = t[2];
    }
    if (w) w[ptr] = oracle_ctx::GetIndividualStatWeightCorrection(pd);
    if (species) species[ptr] = (uint8_t)oracle_ctx::GetI(pd);
    if (alive) alive[ptr] = oracle_ctx::IsParticleAllocated(pd) ? 1 : 0;
  }
  if (cells) {
    for (size_t l = 0; l < o->blocks.size(); l++)
      for (int c = 0; c < nC; c++) {
        long int ptr = o->blocks[l].FirstCellParticleTable[c];
        while (ptr != -1) {
          if (ptr < n) cells[ptr] = (int32_t)(l * nC + c);
          ptr = o->GetNext(ptr);
        }
      }
  }
}

// PIC::Mover::MoveParticles(), src/pic/pic_mover.cpp:580-1088 followed (periodic mode) by
// PIC::BC::ExternalBoundary::Periodic::ExchangeParticles(), src/pic/pic_time_step.cpp:454-506
int oracle_move(oracle_ctx *o, int mover_id, int n_threads, amps_gpu_move_stats *stats, int32_t *ret_code, int32_t *final_cell) {
  if (mover_id != AMPS_MOVER_LAPENTA2017 && mover_id != AMPS_MOVER_RELATIVISTIC_BORIS && mover_id != AMPS_MOVER_BORIS &&
      mover_id != AMPS_MOVER_RELATIVISTIC_GCA && mover_id != AMPS_MOVER_GC_FIRST_ORDER && mover_id != AMPS_MOVER_GC_SECOND_ORDER &&
      mover_id != AMPS_MOVER_MARKIDIS2010 && mover_id != AMPS_MOVER_GYROKINETIC_FIRST_ORDER && mover_id != AMPS_MOVER_GYROKINETIC_SECOND_ORDER) {
    o->err = "oracle_move: mover not restated yet";
    return AMPS_GPU_ERR_ARG;
  }
#ifndef _OPENMP
  n_threads = 1;
#endif
  if (n_threads < 1) n_threads = 1;
  const int nC = o->nCellsBlock();
  const int nLocalBlocks = (int)o->BlockTable.size();

  // per-thread staging arrays, pic_mover.cpp:660-711
  o->E_Corner.resize(n_threads);
  o->B_Center.resize(n_threads);
  for (int t = 0; t < n_threads; t++) {
    o->E_Corner[t].assign((size_t)o->nCornerLocal() * 3, 0.0);
    o->B_Center[t].assign((size_t)((o->cfg.b_mode == AMPS_B_CENTER_BASED) ? o->nCenterLocal() : o->nCornerLocal()) * 3, 0.0);
  }
  if (n_threads > 1 && o->nThreadListTables != n_threads) {
    cTempList init = {-1, -1};
    o->threadListStorage.assign((size_t)nLocalBlocks * n_threads * nC, init);
    for (int l = 0; l < nLocalBlocks; l++) o->blocks[l].tempThreadTable = o->threadListStorage.data() + (size_t)l * n_threads * nC;
    o->nThreadListTables = n_threads;
  }

  long long n_moved = 0, n_cross_cell = 0, n_cross_block = 0, n_left = 0, n_notused = 0, n_error = 0;
  o->nSubSteps = 0;

  auto process_block = [&](int nLocalNode, int thread, long long *cnt) {
    cTreeNode *node = o->BlockTable[nLocalNode];
    cBlock *block = node->block;
    if (!block) return;
    double *E_Corner = o->E_Corner[thread].data(), *B_C = o->B_Center[thread].data();
    if (mover_id == AMPS_MOVER_LAPENTA2017) {
      o->SetBlock_E(E_Corner, node);
      o->SetBlock_B(B_C, node);
    }
    auto mover = [&](long int ptr, cTreeNode **newNode) -> int {
      byte *pd = o->GetParticleDataPointer(ptr);
      if (mover_id == AMPS_MOVER_LAPENTA2017) return o->Lapenta2017(pd, ptr, node, E_Corner, B_C, n_threads, thread, newNode);
      const int spec = oracle_ctx::GetI(pd);
      const double dtLocal = (o->cfg.time_step_mode == AMPS_DT_SPECIES_GLOBAL) ? o->cfg.time_step[spec] : o->cfg.time_step[0];
      if (mover_id == AMPS_MOVER_BORIS) return o->Boris(pd, ptr, dtLocal, node, n_threads, thread, newNode);
      if (mover_id == AMPS_MOVER_RELATIVISTIC_GCA) return o->RelGCA_Mover_FirstOrder(pd, ptr, dtLocal, node, n_threads, thread, newNode);
      if (mover_id == AMPS_MOVER_GC_FIRST_ORDER) return o->GC_Mover_FirstOrder(pd, ptr, dtLocal, node, n_threads, thread, newNode);
      if (mover_id == AMPS_MOVER_GC_SECOND_ORDER) return o->GC_Mover_SecondOrder(pd, ptr, dtLocal, node, n_threads, thread, newNode);
      if (mover_id == AMPS_MOVER_MARKIDIS2010) return o->Markidis2010(pd, ptr, dtLocal, node, n_threads, thread, newNode);
      if (mover_id == AMPS_MOVER_GYROKINETIC_FIRST_ORDER) return o->Gyro_Mover_FirstOrder(pd, ptr, dtLocal, node, n_threads, thread, newNode);
      if (mover_id == AMPS_MOVER_GYROKINETIC_SECOND_ORDER) return o->Gyro_Mover_SecondOrder(pd, ptr, dtLocal, node, n_threads, thread, newNode);
      return o->Relativistic_Boris(pd, ptr, dtLocal, node, n_threads, thread, newNode);
    };
    long int *FirstCellParticleTable = block->FirstCellParticleTable;
    for (int i = 0; i < nC; i++) {
      long int ParticleList = FirstCellParticleTable[i];
      if (n_threads <= 1) {
        // mpi_only, pic_mover.cpp:955-973
        while (FirstCellParticleTable[i] != -1) {
          long int ptr = FirstCellParticleTable[i];
          FirstCellParticleTable[i] = o->GetNext(ptr);
          cTreeNode *newNode = NULL;
          int rc = mover(ptr, &newNode);
          cnt[0]++;
          if (ret_code) ret_code[ptr] = rc;
          if (rc == _PARTICLE_MOTION_FINISHED_) {
            if (newNode != node) cnt[2]++;
          } else if (rc == _PARTICLE_LEFT_THE_DOMAIN_) cnt[3]++;
          else if (rc == _PARTICLE_IN_NOT_IN_USE_NODE_) cnt[4]++;
          else cnt[5]++;
        }
      } else {
        // mpi_openmp__split_blocks, pic_mover.cpp:744-759
        while (ParticleList != -1) {
          long int ptr = ParticleList;
          ParticleList = o->GetNext(ParticleList);
          cTreeNode *newNode = NULL;
          int rc = mover(ptr, &newNode);
          cnt[0]++;
          if (ret_code) ret_code[ptr] = rc;
          if (rc == _PARTICLE_MOTION_FINISHED_) {
            if (newNode != node) cnt[2]++;
          } else if (rc == _PARTICLE_LEFT_THE_DOMAIN_) cnt[3]++;
          else if (rc == _PARTICLE_IN_NOT_IN_USE_NODE_) cnt[4]++;
          else cnt[5]++;
        }
      }
    }
  };

  // remember the starting cell of each particle for the cross-cell statistic
  std::vector<int32_t> startCell;
  if (stats) {
    startCell.assign(o->MaxNPart, -1);
    for (int l = 0; l < nLocalBlocks; l++)
      for (int c = 0; c < nC; c++)
        for (long int ptr = o->blocks[l].FirstCellParticleTable[c]; ptr != -1; ptr = o->GetNext(ptr)) startCell[ptr] = l * nC + c;
  }

  if (n_threads <= 1) {
    long long cnt[6] = {0, 0, 0, 0, 0, 0};
    for (int nLocalNode = 0; nLocalNode < nLocalBlocks; nLocalNode++) process_block(nLocalNode, 0, cnt);
    n_moved = cnt[0], n_cross_block = cnt[2], n_left = cnt[3], n_notused = cnt[4], n_error = cnt[5];
  } else {
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads) reduction(+ : n_moved, n_cross_block, n_left, n_notused, n_error)
    {
      long long cnt[6] = {0, 0, 0, 0, 0, 0};
      int thread = omp_get_thread_num();
#pragma omp for schedule(dynamic, 1)
      for (int nLocalNode = 0; nLocalNode < nLocalBlocks; nLocalNode++) process_block(nLocalNode, thread, cnt);
      n_moved += cnt[0], n_cross_block += cnt[2], n_left += cnt[3], n_notused += cnt[4], n_error += cnt[5];
    }
#endif
  }

  // update the particle lists, pic_mover.cpp:1056-1088
  for (int l = 0; l < nLocalBlocks; l++) {
    cBlock *block = &o->blocks[l];
    if (n_threads <= 1) {  // update_mpi
      memcpy(block->FirstCellParticleTable, block->tempParticleMovingListTable, nC * sizeof(long int));
      for (int c = 0; c < nC; c++) block->tempParticleMovingListTable[c] = -1;
    } else {  // update_openmp
      for (int c = 0; c < nC; c++) block->FirstCellParticleTable[c] = -1;
      for (int thread_OpenMP = 0; thread_OpenMP < n_threads; thread_OpenMP++)
        for (int c = 0; c < nC; c++) {
          cTempList *t = block->tempThreadTable + (size_t)thread_OpenMP * nC + c;
          long int LastParticle = t->last;
          if (LastParticle != -1) {
            long int FirstParticle = t->first;
            long int *FirstCellParticlePtr = block->FirstCellParticleTable + c;
            o->SetNext(*FirstCellParticlePtr, LastParticle);
            if (*FirstCellParticlePtr != -1) o->SetPrev(LastParticle, *FirstCellParticlePtr);
            *FirstCellParticlePtr = FirstParticle;
          }
          t->first = -1;
          t->last = -1;
        }
    }
  }

  // periodic wrap, pic_bc_periodic.cpp:55-83
  long long n_wrap = 0;
  if (o->cfg.periodic) {
    for (int l = 0; l < nLocalBlocks; l++)
      if (o->leaf_real[l] >= 0) n_wrap += o->ExchangeParticlesLocal(o->BlockTable[o->leaf_real[l]], o->BlockTable[l]);
  }

  if (final_cell || stats) {
    std::vector<int32_t> fc(o->MaxNPart, -1);
    for (int l = 0; l < nLocalBlocks; l++)
      for (int c = 0; c < nC; c++)
        for (long int ptr = o->blocks[l].FirstCellParticleTable[c]; ptr != -1; ptr = o->GetNext(ptr)) fc[ptr] = l * nC + c;
    if (final_cell) memcpy(final_cell, fc.data(), sizeof(int32_t) * o->MaxNPart);
    if (stats) {
      n_cross_cell = 0;
      long long xb = 0;
      for (long int p = 0; p < o->MaxNPart; p++)
        if (startCell[p] >= 0 && fc[p] >= 0 && fc[p] != startCell[p]) {
          if (fc[p] / nC == startCell[p] / nC) n_cross_cell++;
          else xb++;
        }
      n_cross_block = xb;
    }
  }
  if (stats) {
    stats->n_moved = n_moved;
    stats->n_cross_cell = n_cross_cell;
    stats->n_cross_block = n_cross_block;
    stats->n_left_domain = n_left;
    stats->n_not_in_use = n_notused;
    stats->n_periodic_wrap = n_wrap;
    stats->n_error = n_error;
    stats->n_sub_steps = o->nSubSteps;
  }
  return n_error ? AMPS_GPU_ERR_PARTICLE : AMPS_GPU_OK;
}

// ECSIM::UpdateJMassMatrix(), src/pic/pic_field_solver_ecsim.cpp:3244-3995 (serial / OpenMP loop :3794-3907)
int oracle_deposit_JM(oracle_ctx *o, int n_threads, double *J, double *M, double *energy, double *cfl) {
#ifndef _OPENMP
  n_threads = 1;
#endif
  if (n_threads < 1) n_threads = 1;
  const int nC = o->nCellsBlock();
  const int nLocalBlocks = (int)o->BlockTable.size();
  const int nSpecies = o->cfg.n_species;
  const int NX = o->_BLOCK_CELLS_X_, NY = o->_BLOCK_CELLS_Y_;

  // SetCornerNodeAssociatedDataValue(0.0,...), :3266-3267
  for (int i = 0; i < o->n_corners; i++) {
    double *d = o->cornerPool[i].data;
    for (int k = 0; k < 3; k++) d[JxOffsetIndex + k] = 0.0;
    for (int k = 0; k < 243; k++) d[MassMatrixOffsetIndex + k] = 0.0;
  }

  double MassTable[AMPS_GPU_MAX_SPECIES], ChargeTable[AMPS_GPU_MAX_SPECIES];
  for (int s = 0; s < AMPS_GPU_MAX_SPECIES; s++) MassTable[s] = o->cfg.mass[s], ChargeTable[s] = o->cfg.charge[s];

  std::vector<double> ParticleEnergyTable(n_threads, 0.0);
  std::vector<double> cflTable((size_t)n_threads * AMPS_GPU_MAX_SPECIES, 0.0);
  std::vector<cCellData> CellDataTable(n_threads);

  const long nTotalCells = (long)nLocalBlocks * nC;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
  {
#ifdef _OPENMP
    int this_thread_id = omp_get_thread_num();
#else
    int this_thread_id = 0;
#endif
    cCellData *CellData_TH = &CellDataTable[this_thread_id];
#ifdef _OPENMP
#pragma omp for schedule(guided)
#endif
    for (long CellCounter = 0; CellCounter < nTotalCells; CellCounter++) {
      int nLocalNode, ii = (int)(CellCounter % nC);
      int i, j, k;
      nLocalNode = (int)(CellCounter / nC);
      k = ii / (NY * NX);
      ii -= k * NY * NX;
      j = ii / NX;
      ii -= j * NX;
      i = ii;

      cTreeNode *node = o->BlockTable[nLocalNode];
      if (node->block == NULL) continue;
      if (o->cfg.periodic) {
        // boundary ("ghost") blocks are skipped, :3815-3825
        if (node->faceBoundary != 0) continue;
      }

      CellData_TH->clean();
      bool flag = o->ProcessCell(i, j, k, node, CellData_TH, MassTable, ChargeTable);

      if (flag == true) {
        int CornerUpdateTable[8] = {0, 1, 2, 3, 4, 5, 6, 7};
        int CornerUpdateTableLength = 8;
        while (CornerUpdateTableLength > 0) {
          for (int it = 0; it < CornerUpdateTableLength; it++) {
            int icor = CornerUpdateTable[it];
            if (CellData_TH->CornerData[icor].CornerNode->lock_associated_data.test_and_set(std::memory_order_acquire) == false) {
              // NOTE (reference quirk): the cell energy is added once per corner, :3860
              ParticleEnergyTable[this_thread_id] += CellData_TH->ParticleEnergy;
              for (int iSp = 0; iSp < nSpecies; iSp++) {
                double *c = &cflTable[(size_t)this_thread_id * AMPS_GPU_MAX_SPECIES + iSp];
                if (CellData_TH->cflCell[iSp] > *c) *c = CellData_TH->cflCell[iSp];
              }
              double *target = CellData_TH->CornerData[icor].CornerJ_ptr;
              double *source = CellData_TH->CornerData[icor].CornerJ;
              for (int idim = 0; idim < 3; idim++) target[idim] += source[idim];
              target = CellData_TH->CornerData[icor].CornerMassMatrix_ptr;
              source = CellData_TH->CornerData[icor].CornerMassMatrix;
              for (int q = 0; q < 243; q++) target[q] += source[q];
              CornerUpdateTable[it] = CornerUpdateTable[CornerUpdateTableLength - 1];
              CornerUpdateTableLength--;
              CellData_TH->CornerData[icor].CornerNode->lock_associated_data.clear(std::memory_order_release);
            }
          }
        }
      }
    }
  }

  double ParticleEnergy = 0.0;
  for (int i = 0; i < n_threads; i++) ParticleEnergy += ParticleEnergyTable[i];
  if (energy) *energy = ParticleEnergy;
  if (cfl)
    for (int iSp = 0; iSp < nSpecies; iSp++) {
      cfl[iSp] = cflTable[iSp];
      for (int i = 0; i < n_threads; i++)
        if (cfl[iSp] < cflTable[(size_t)i * AMPS_GPU_MAX_SPECIES + iSp]) cfl[iSp] = cflTable[(size_t)i * AMPS_GPU_MAX_SPECIES + iSp];
    }
  // periodic / shared corners are the same node objects here, so ProcessJMassMatrix (:1383) is implicit
  if (J)
    for (int i = 0; i < o->n_corners; i++) memcpy(J + 3 * (size_t)i, o->cornerPool[i].data + JxOffsetIndex, 24);
  if (M)
    for (int i = 0; i < o->n_corners; i++) memcpy(M + 243 * (size_t)i, o->cornerPool[i].data + MassMatrixOffsetIndex, 243 * 8);
  return AMPS_GPU_OK;
}

int oracle_find_tree_node(const oracle_ctx *o, const double *x, int start_leaf) {
  cTreeNode *start = (start_leaf >= 0) ? o->BlockTable[start_leaf] : NULL;
  cTreeNode *r = o->findTreeNode(x, start);
  return r ? r->leaf : -1;
}
int oracle_find_cell_index(const oracle_ctx *o, const double *x, int leaf, int *ijk) {
  int i = 0, j = 0, k = 0;
  long int r = o->FindCellIndex(x, i, j, k, o->BlockTable[leaf]);
  ijk[0] = i, ijk[1] = j, ijk[2] = k;
  return (int)r;
}
int oracle_corner_stencil(const oracle_ctx *o, double *x_inout, int leaf, double *W, int *ids, double *wnorm) {
  cStencil s;
  if (!o->CornerBased_InitStencil(x_inout, o->BlockTable[leaf], s, W)) return -1;
  for (int i = 0; i < s.Length; i++) ids[i] = s.LocalCellID[i], wnorm[i] = s.Weight[i];
  return s.Length;
}
int oracle_center_stencil(const oracle_ctx *o, const double *x, int leaf, int *ids, double *w) {
  cStencil s;
  o->CellCentered_Linear_InitStencil(x, o->BlockTable[leaf], s, true);
  for (int i = 0; i < s.Length; i++) ids[i] = s.LocalCellID[i], w[i] = s.Weight[i];
  return s.Length;
}

// PIC::CPLR::InitInterpolationStencil at x in `leaf` (AMR capable): unique centre-node ids and weights; -1 = the reference exit()s
int oracle_coupler_stencil(const oracle_ctx *o, const double *x, int leaf, int *uids, double *w) {
  cStencil s;
  if (!o->CplrInitStencil(x, o->BlockTable[leaf], s)) return -1;
  for (int i = 0; i < s.Length; i++) uids[i] = (int)(s.cell[i] - o->centerPool.data()), w[i] = s.Weight[i];
  return s.Length;
}
// ECSIM::GetElectricField / GetMagneticField / GetMagneticFieldGradient (pic_field_solver_ecsim.cpp:7440-7547) at n points, each in
// its `leaf`: E[n][3], B[n][3], gradB[n][9]; returns the number of points where the reference would have exit()ed
int oracle_ecsim_fields(const oracle_ctx *o, int64_t n, const double *x, const int32_t *leaf, double *E, double *B, double *gradB) {
  int bad = 0;
  for (int64_t i = 0; i < n; i++) {
    cTreeNode *node = o->BlockTable[leaf[i]];
    bool ok = o->ECSIM_GetElectricField(E + 3 * i, x + 3 * i, node);
    ok = o->ECSIM_GetMagneticField(B + 3 * i, x + 3 * i, node) && ok;
    ok = o->ECSIM_GetMagneticFieldGradient(gradB + 9 * i, x + 3 * i, node) && ok;
    if (!ok) bad++;
  }
  return bad;
}
// neighbour of a leaf: kind 0 = GetNeibFace(idx,0,0), 1 = GetNeibEdge(idx,0), 2 = GetNeibCorner(idx); returns the node id or -1,
// geometry in lo/hi/level
int oracle_neib(const oracle_ctx *o, int leaf, int kind, int idx, double *lo, double *hi, int *level) {
  cTreeNode *n = o->BlockTable[leaf];
  cTreeNode *nb = (kind == 0) ? o->GetNeibFace(n, idx, 0, 0) : (kind == 1) ? o->GetNeibEdge(n, idx, 0) : o->GetNeibCorner(n, idx);
  if (!nb) return -1;
  for (int d = 0; d < 3; d++) lo[d] = nb->xmin[d], hi[d] = nb->xmax[d];
  *level = nb->RefinmentLevel;
  return nb->id;
}
// (min, max) neighbour refinement levels of a leaf (SetNeibRefinmentLevelLimits)
void oracle_neib_levels(const oracle_ctx *o, int leaf, int *minmax) {
  minmax[0] = o->BlockTable[leaf]->minNeibRefinmentLevel, minmax[1] = o->BlockTable[leaf]->maxNeibRefinmentLevel;
}

// PIC::ParticleBuffer::CheckParticleList, src/pic/pic_pbuffer.cpp:807-: every allocated particle is on
// exactly one cell list and prev/next are consistent
int oracle_check_particle_lists(const oracle_ctx *o) {
  const int nC = o->nCellsBlock();
  std::vector<uint8_t> seen(o->MaxNPart, 0);
  long int nList = 0;
  for (size_t l = 0; l < o->blocks.size(); l++)
    for (int c = 0; c < nC; c++) {
      long int prev = -1;
      for (long int ptr = o->blocks[l].FirstCellParticleTable[c]; ptr != -1; ptr = o->GetNext(ptr)) {
        const byte *pd = o->GetParticleDataPointer(ptr);
        if (!oracle_ctx::IsParticleAllocated(pd)) return 1;
        if (seen[ptr]) return 2;
        if (oracle_ctx::GetPrev(pd) != prev) return 3;
        seen[ptr] = 1;
        prev = ptr;
        nList++;
      }
    }
  if (nList != o->NAllPart) return 4;
  return 0;
}

}  // extern "C"
