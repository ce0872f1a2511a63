// cplr_stencil.cuh -- a4, AMR branch of the coupler's cell-centred linear stencil (device side).
//
//   PIC::InterpolationRoutines::CellCentered::Linear::InitStencil      pic_interpolation_routines.cpp:224-706
//   GetTriliniarInterpolationStencil :820-909, GetTriliniarInterpolationMutiBlockStencil :912-1070,
//   Constant::InitStencil :168-220, cStencilGeneric::AddCell / Add / MultiplyScalar  pic.h:7228-7292,
//   neighbours by lattice probe  meshAMRgeneric.h:505-725
//
// Near a coarse/fine interface the stencil is built on the COARSE lattice: each of the 8 logical coarse centres is
// either a cell of the coarse block (through its ghost layer) or the average of the 2x2x2 fine cells that cover
// it; inside the fine block, between half a cell and one cell from the interface, the coarse stencil is blended
// with the block's own trilinear stencil (alpha = (dmin-0.5)/0.5).  A stencil therefore holds up to 64 cells of
// several blocks; entries are unique centre-node ids and the fields are gathered from the unique-node tables.
// Only test-particle movers come here, and only in leaves whose neighbourhood is not single-level
// (LeafGeo::neib); include after find_tree_node_plain / find_cell_index.
#pragma once

// CPLR_INLINE: the kernels that include this header grew to 17 000 instructions with everything inlined (instruction-cache bound,
// profiles/r2_tp_movers_ncu_summary.txt); the three stencil builders are calls now
#ifndef CPLR_INLINE
#define CPLR_INLINE __noinline__
#endif

namespace amps {

constexpr int CPLR_MAX_STENCIL = 64;  // nMaxStencilLength, pic.h:7185

struct BgStencil {
  double w[CPLR_MAX_STENCIL];
  int nd[CPLR_MAX_STENCIL];  // tile-local centre number, or unique centre id when `uid` is set
  int n;
  int uid;
  int overflow;  // the reference exit()s when Length would exceed nMaxStencilLength
};

__device__ __forceinline__ int leaf_min_neib(const LeafGeo &lg) { return (int)(short)(lg.neib & 0xffff); }
__device__ __forceinline__ int leaf_max_neib(const LeafGeo &lg) { return (int)(short)((lg.neib >> 16) & 0xffff); }
__device__ __forceinline__ bool leaf_single_level(const LeafGeo &lg) { return leaf_min_neib(lg) == lg.level && leaf_max_neib(lg) == lg.level; }

// cStencilGeneric::AddCell: a centre outside the global box is not added (non-periodic), pic.h:7235-7245
__device__ __forceinline__ bool cs_center_outside(const DevMesh &m, int node, const int ijk[3]) {
  if (m.periodic) return false;
  {
    // a block with a neighbour across every face lies inside the box in all three directions with a margin of half its width (the
    // neighbour is at most one level finer), and so do the centres of its ghost layers while g - 1/2 < N/2
    const int leaf = m.nodeLeaf[node];
    if (leaf >= 0 && m.leaf[leaf].face == 0 && 2 * m.g[0] - 1 < m.N[0] && 2 * m.g[1] - 1 < m.N[1] && 2 * m.g[2] - 1 < m.N[2]) return false;
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double lo = m.nxmin[3 * node + d], hi = m.nxmax[3 * node + d];
    const double x = lo + (ijk[d] + 0.5) * ((hi - lo) / m.N[d]);
    if (x < m.xGlobalMin[d] || x > m.xGlobalMax[d]) return true;
  }
  return false;
}
// unique id of centre (i,j,k) of `node`, -1 = no such node (outside the tile, block not allocated here, no node)
__device__ __forceinline__ int cs_center_uid(const DevMesh &m, int node, const int ijk[3]) {
  const int leaf = m.nodeLeaf[node];
  if (leaf < 0) return -1;
#pragma unroll
  for (int d = 0; d < 3; d++)
    if (ijk[d] < -m.g[d] || ijk[d] > m.N[d] + m.g[d] - 1) return -1;
  return m.centerUid[(size_t)leaf * m.nCenterLocal + centerLocalNumber(m, ijk[0], ijk[1], ijk[2])];
}
__device__ __forceinline__ void cs_add_cell(const DevMesh &m, BgStencil &S, double w, int node, const int ijk[3], int uid) {
  if (S.n == CPLR_MAX_STENCIL) {
    S.overflow = 1;
    return;
  }
  if (cs_center_outside(m, node, ijk)) return;
  S.w[S.n] = w, S.nd[S.n] = uid;
  S.n++;
}
// AddPhysicalStencilCell, pic_interpolation_routines.cpp:124-134
__device__ __forceinline__ void cs_add_physical(const DevMesh &m, BgStencil &S, double w, int node, const int ijk[3], int uid) {
  for (int e = 0; e < S.n; e++)
    if (S.nd[e] == uid) {
      S.w[e] += w;
      return;
    }
  cs_add_cell(m, S, w, node, ijk, uid);
}
__device__ __forceinline__ void cs_flush(BgStencil &S) { S.n = 0, S.uid = 1, S.overflow = 0; }

// Constant::InitStencil; false where the reference exit()s
__device__ CPLR_INLINE bool cs_constant(const DevMesh &m, const double x[3], int node, BgStencil &S) {
  cs_flush(S);
  if (node < 0 || m.nodeLeaf[node] < 0) return false;
  int ijk[3];
  if (!find_cell_index(m, x, node, ijk)) return false;
  const int uid = cs_center_uid(m, node, ijk);
  if (uid < 0) return false;
  cs_add_cell(m, S, 1.0, node, ijk, uid);
  return true;
}

// GetTriliniarInterpolationStencil; tableLen points at the Length the reference tests at :903 (StencilTable->Length)
__device__ CPLR_INLINE bool cs_trilinear(const DevMesh &m, const double loc[3], const double x[3], int node, BgStencil &S, const int *tableLen) {
  if (m.nodeLeaf[node] < 0) return false;
  cs_flush(S);
  int o[3];
  double w[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    o[d] = (loc[d] < 0.5) ? -1 : (int)(loc[d] - 0.50);
    if (o[d] < -m.g[d] || o[d] + 1 > m.N[d] + m.g[d] - 1) return false;  // out-of-bounds read in the reference
    w[d] = loc[d] - (o[d] + 0.5);
  }
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++)
      for (int k = 0; k < 2; k++) {
        const double wi = i ? w[0] : 1.0 - w[0], wj = j ? w[1] : 1.0 - w[1], wk = k ? w[2] : 1.0 - w[2];
        const double W = wi * wj * wk;
        const int ijk[3] = {o[0] + i, o[1] + j, o[2] + k};
        const int uid = cs_center_uid(m, node, ijk);
        if (uid >= 0) cs_add_cell(m, S, W, node, ijk, uid);
      }
  if (S.n == 0) return cs_constant(m, x, node, S);
  if (*tableLen != 8) {  // Normalize()
    double norm = 0.0;
    for (int e = 0; e < S.n; e++) norm += S.w[e];
    if (norm > 0.0)
      for (int e = 0; e < S.n; e++) S.w[e] /= norm;
  }
  return true;
}

// GetTriliniarInterpolationMutiBlockStencil
__device__ __forceinline__ bool cs_multiblock_cached(const DevMesh &m, const double x[3], int node, BgStencil &S, bool &ok);
__device__ CPLR_INLINE bool cs_multiblock(const DevMesh &m, const double x[3], int node, BgStencil &S) {
  {
    bool ok;
    if (cs_multiblock_cached(m, x, node, S, ok)) return ok;
  }
  cs_flush(S);
  double dxCell[3], xLoc[3], nlo[3];
  int ijkMin[3];
  const int nodeLevel = m.nodeLevel[node];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    nlo[d] = m.nxmin[3 * node + d];
    dxCell[d] = (m.nxmax[3 * node + d] - nlo[d]) / m.N[d];
    ijkMin[d] = (x[d] - nlo[d] < 0.5 * dxCell[d]) ? -1 : (int)((x[d] - nlo[d] - 0.5 * dxCell[d]) / dxCell[d]);
    const double xLower = nlo[d] + (ijkMin[d] + 0.5) * dxCell[d];
    xLoc[d] = (x[d] - xLower) / dxCell[d];
    if (xLoc[d] < 0.0) xLoc[d] = 0.0;
    if (xLoc[d] > 1.0) xLoc[d] = 1.0;
  }
  bool geometryAvailable = true, physicalAvailable = true;
  for (int di = 0; di < 2; di++)
    for (int dj = 0; dj < 2; dj++)
      for (int dk = 0; dk < 2; dk++) {
        const int ijk[3] = {ijkMin[0] + di, ijkMin[1] + dj, ijkMin[2] + dk};
        double xLogical[3];
#pragma unroll
        for (int d = 0; d < 3; d++) xLogical[d] = nlo[d] + (ijk[d] + 0.5) * dxCell[d];
        const double wi = di ? xLoc[0] : 1.0 - xLoc[0], wj = dj ? xLoc[1] : 1.0 - xLoc[1], wk = dk ? xLoc[2] : 1.0 - xLoc[2];
        const double W = wi * wj * wk;
        const int sn = find_tree_node_plain(m, xLogical, node);
        if (sn < 0 || !(m.nodeFlags[sn] & AMPS_NODE_USED)) {
          geometryAvailable = false;
          continue;
        }
        if (m.nodeLevel[sn] == nodeLevel) {
          const int uid = cs_center_uid(m, node, ijk);
          if (uid >= 0) cs_add_physical(m, S, W, node, ijk, uid);
          else physicalAvailable = false;
        } else {
          int iNeib[3];
#pragma unroll
          for (int d = 0; d < 3; d++) iNeib[d] = 2 * ((int)((xLogical[d] - m.nxmin[3 * sn + d]) / dxCell[d]));
          const double fw = (1.0 / 8.0) * W;
          for (int ii = 0; ii < 2; ii++)
            for (int jj = 0; jj < 2; jj++)
              for (int kk = 0; kk < 2; kk++) {
                const int f[3] = {iNeib[0] + ii, iNeib[1] + jj, iNeib[2] + kk};
                const int uid = cs_center_uid(m, sn, f);
                if (uid >= 0) cs_add_physical(m, S, fw, sn, f, uid);
                else physicalAvailable = false;
              }
        }
      }
  if (!geometryAvailable || !physicalAvailable) {
    const int in = find_tree_node_plain(m, x, node);
    if (in >= 0 && m.nodeLeaf[in] >= 0) return cs_constant(m, x, in, S);
    cs_flush(S);
  }
  return true;
}

// ---- structure cache of cs_multiblock -------------------------------------------------------------------------------------------
// Which cells the 8 logical coarse centres of a multi-block stencil resolve to (one coarse cell, the 8 fine cells that cover it, or
// nothing) depends on the coarse block and on the dual cell ijkMin in [-1, N]^3 only, not on the point: 8 tree searches, up to 64
// centre look-ups and the duplicate search of AddPhysicalStencilCell per particle and sub-step (16 000 instructions of a divergent
// warp, profiles/r2_tp_movers_*) become a table entry that is replayed with the point's trilinear weights in the same order, so
// the weights are bit-identical to the uncached path (tests/test_cplr_cache_gpu.py).
constexpr int MB_ENTRY = 392;  // int hdr (adds | uniques << 8 | flags << 16), int pad, int uid[64], u8 slot[64], u8 code[64]
constexpr int MB_UID = 8, MB_SLOT = 8 + 256, MB_CODE = 8 + 256 + 64;
constexpr int MB_FALLBACK = 1, MB_OVERFLOW = 2;  // hdr flags: a logical centre had no geometry / physical cell; Length hit nMaxStencilLength
constexpr int MB_FINE = 8, MB_FIRST = 16;        // code bits next to the corner number 4 di + 2 dj + dk

__device__ __forceinline__ int mb_entries(const DevMesh &m) { return (m.N[0] + 2) * (m.N[1] + 2) * (m.N[2] + 2); }

// the structure of cs_multiblock(m, *, node, *) for the dual cell ijkMin, written to ent[MB_ENTRY]
__device__ inline void cs_multiblock_structure(const DevMesh &m, int node, const int ijkMin[3], unsigned char *ent) {
  int *uid = reinterpret_cast<int *>(ent + MB_UID);
  int nAdds = 0, nU = 0, flags = 0;
  double dxCell[3], nlo[3];
  const int nodeLevel = m.nodeLevel[node];
  for (int d = 0; d < 3; d++) {
    nlo[d] = m.nxmin[3 * node + d];
    dxCell[d] = (m.nxmax[3 * node + d] - nlo[d]) / m.N[d];
  }
  auto add = [&](int id, int code, int owner, const int ijk[3]) {  // cs_add_physical
    for (int e = 0; e < nU; e++)
      if (uid[e] == id) {
        ent[MB_SLOT + nAdds] = (unsigned char)e, ent[MB_CODE + nAdds] = (unsigned char)code;
        nAdds++;
        return;
      }
    if (nU == CPLR_MAX_STENCIL) {
      flags |= MB_OVERFLOW;
      return;
    }
    if (cs_center_outside(m, owner, ijk)) return;
    uid[nU] = id;
    ent[MB_SLOT + nAdds] = (unsigned char)nU, ent[MB_CODE + nAdds] = (unsigned char)(code | MB_FIRST);
    nAdds++, nU++;
  };
  for (int di = 0; di < 2; di++)
    for (int dj = 0; dj < 2; dj++)
      for (int dk = 0; dk < 2; dk++) {
        const int corner = 4 * di + 2 * dj + dk;
        const int ijk[3] = {ijkMin[0] + di, ijkMin[1] + dj, ijkMin[2] + dk};
        double xLogical[3];
        for (int d = 0; d < 3; d++) xLogical[d] = nlo[d] + (ijk[d] + 0.5) * dxCell[d];
        const int sn = find_tree_node_plain(m, xLogical, node);
        if (sn < 0 || !(m.nodeFlags[sn] & AMPS_NODE_USED)) {
          flags |= MB_FALLBACK;
          continue;
        }
        if (m.nodeLevel[sn] == nodeLevel) {
          const int id = cs_center_uid(m, node, ijk);
          if (id >= 0) add(id, corner, node, ijk);
          else flags |= MB_FALLBACK;
        } else {
          int iNeib[3];
          for (int d = 0; d < 3; d++) iNeib[d] = 2 * ((int)((xLogical[d] - m.nxmin[3 * sn + d]) / dxCell[d]));
          for (int ii = 0; ii < 2; ii++)
            for (int jj = 0; jj < 2; jj++)
              for (int kk = 0; kk < 2; kk++) {
                const int f[3] = {iNeib[0] + ii, iNeib[1] + jj, iNeib[2] + kk};
                const int id = cs_center_uid(m, sn, f);
                if (id >= 0) add(id, corner | MB_FINE, sn, f);
                else flags |= MB_FALLBACK;
              }
        }
      }
  *reinterpret_cast<int *>(ent) = nAdds | (nU << 8) | (flags << 16);
}

// cs_multiblock through the table of the coarse block; false = no table for this block / dual cell (the caller builds the stencil)
__device__ __forceinline__ bool cs_multiblock_cached(const DevMesh &m, const double x[3], int node, BgStencil &S, bool &ok) {
  if (m.mbSlot == nullptr) return false;
  const int leaf = m.nodeLeaf[node];
  if (leaf < 0) return false;
  const int tab = m.mbSlot[leaf];
  if (tab < 0) return false;
  double xLoc[3];
  int ijkMin[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double nlo = m.nxmin[3 * node + d];
    const double dxCell = (m.nxmax[3 * node + d] - nlo) / m.N[d];
    ijkMin[d] = (x[d] - nlo < 0.5 * dxCell) ? -1 : (int)((x[d] - nlo - 0.5 * dxCell) / dxCell);
    if (ijkMin[d] > m.N[d]) return false;
    const double xLower = nlo + (ijkMin[d] + 0.5) * dxCell;
    xLoc[d] = (x[d] - xLower) / dxCell;
    if (xLoc[d] < 0.0) xLoc[d] = 0.0;
    if (xLoc[d] > 1.0) xLoc[d] = 1.0;
  }
  const int e = (ijkMin[0] + 1) + (m.N[0] + 2) * ((ijkMin[1] + 1) + (m.N[1] + 2) * (ijkMin[2] + 1));
  const unsigned char *ent = m.mbTab + ((size_t)tab * mb_entries(m) + e) * MB_ENTRY;
  const int hdr = *reinterpret_cast<const int *>(ent);
  const int nAdds = hdr & 255, nU = (hdr >> 8) & 255, flags = hdr >> 16;
  cs_flush(S);
  ok = true;
  if (flags & MB_FALLBACK) {
    const int in = find_tree_node_plain(m, x, node);
    if (in >= 0 && m.nodeLeaf[in] >= 0) ok = cs_constant(m, x, in, S);
    return true;
  }
  double W[8];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const double wi = (c & 4) ? xLoc[0] : 1.0 - xLoc[0], wj = (c & 2) ? xLoc[1] : 1.0 - xLoc[1], wk = (c & 1) ? xLoc[2] : 1.0 - xLoc[2];
    W[c] = wi * wj * wk;
  }
  const int *uid = reinterpret_cast<const int *>(ent + MB_UID);
  for (int q = 0; q < nU; q++) S.nd[q] = uid[q];
  for (int k = 0; k < nAdds; k++) {
    const int code = ent[MB_CODE + k], slot = ent[MB_SLOT + k];
    double w = W[code & 7];
    if (code & MB_FINE) w = (1.0 / 8.0) * w;
    if (code & MB_FIRST) S.w[slot] = w;
    else S.w[slot] += w;
  }
  S.n = nU;
  S.overflow = (flags & MB_OVERFLOW) ? 1 : 0;
  return true;
}

// neighbours through face / edge / corner, first segment (GetNeibFace(f,0,0), GetNeibEdge(e,0), GetNeibCorner(c))
__device__ CPLR_INLINE int cs_neib(const DevMesh &m, int node, const int side[3]) {
  // side[d]: -1 beyond the low face, +1 beyond the high face, 0 = at the block's low index
  if (m.neib26 != nullptr) {
    const int leaf = m.nodeLeaf[node];
    if (leaf >= 0) return m.neib26[(size_t)leaf * 27 + (side[0] + 1) + 3 * (side[1] + 1) + 9 * (side[2] + 1)];
  }
  int ix[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    ix[d] = m.imin[3 * node + d];
    if (side[d] < 0) ix[d] -= 1;
    else if (side[d] > 0) ix[d] += m.isize[node];
  }
  return find_node_ix(m, ix[0], ix[1], ix[2]);
}

// Linear::InitStencil for a leaf whose neighbourhood is not single-level
__device__ __noinline__ bool cplr_linear_stencil_amr(const DevMesh &m, const double x[3], int leaf, BgStencil &S) {
  const LeafGeo &lg = m.leaf[leaf];
  const int node = lg.node;
  cs_flush(S);
  double loc[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    loc[d] = (x[d] - lg.xmin[d]) / (lg.xmax[d] - lg.xmin[d]) * m.N[d];
    if (!(loc[d] >= -1.0e9 && loc[d] <= 1.0e9)) return false;
  }
  const int minNeib = leaf_min_neib(lg), maxNeib = leaf_max_neib(lg);
  if (lg.level == minNeib && lg.level == maxNeib) return cs_trilinear(m, loc, x, node, S, &S.n);
  if ((1.0 < loc[0]) && (loc[0] < m.N[0] - 1) && (1.0 < loc[1]) && (loc[1] < m.N[1] - 1) && (1.0 < loc[2]) && (loc[2] < m.N[2] - 1))
    return cs_trilinear(m, loc, x, node, S, &S.n);
  if (lg.level == minNeib) {
    if ((0.5 < loc[0]) && (loc[0] < m.N[0] - 0.5) && (0.5 < loc[1]) && (loc[1] < m.N[1] - 0.5) && (0.5 < loc[2]) && (loc[2] < m.N[2] - 0.5))
      return cs_trilinear(m, loc, x, node, S, &S.n);
    return cs_multiblock(m, x, node, S);
  }

  // look for a coarser block next to the point (:332-655)
  int coarser = -1;
  double dxCell[3];
#pragma unroll
  for (int d = 0; d < 3; d++) dxCell[d] = (lg.xmax[d] - lg.xmin[d]) / m.N[d];
  double dmin = 10.0 * m.N[0] * m.N[1] * m.N[2];
  unsigned cornerTested = 0, edgeTested = 0;
  auto usable = [&](int nb) -> bool {
    if (nb < 0) return false;
    if (!((m.nodeLevel[nb] < lg.level) && (m.nodeFlags[nb] & AMPS_NODE_USED))) return false;
    int cnt = 0;
    for (int d = 0; d < 3; d++)
      if ((m.nxmin[3 * nb + d] - dxCell[d] <= x[d]) && (m.nxmax[3 * nb + d] + dxCell[d] >= x[d])) cnt++;
    return cnt == 3;
  };
  // distance to the block boundary across direction d on side s (0 low, 1 high) and the "is next to it" test
  auto near = [&](int d, int s, double &dist) -> bool {
    if (s == 0) {
      dist = loc[d];
      return loc[d] < 1.0;
    }
    dist = m.N[d] - loc[d];
    return loc[d] > m.N[d] - 1;
  };
  for (int idim = 0; idim < 3; idim++) {
    int iFace;
    if (loc[idim] <= 1.0) iFace = 2 * idim;
    else if (loc[idim] >= m.N[idim] - 1.0) iFace = 2 * idim + 1;
    else continue;
    {
      int side[3] = {0, 0, 0};
      side[idim] = (iFace & 1) ? 1 : -1;
      const int nb = cs_neib(m, node, side);
      double dist;
      if (usable(nb) && near(idim, iFace & 1, dist) && dist < dmin) dmin = dist, coarser = nb;
    }
    // edges of the face, in the reference's order faceEdges[iFace][0..3]
    const int faceEdges[6][4] = {{4, 11, 7, 8}, {5, 10, 6, 9}, {0, 9, 3, 8}, {1, 10, 2, 11}, {0, 5, 1, 4}, {3, 6, 2, 7}};
    const int edgeDir[12][2] = {{1, 2}, {1, 2}, {1, 2}, {1, 2}, {0, 2}, {0, 2}, {0, 2}, {0, 2}, {0, 1}, {0, 1}, {0, 1}, {0, 1}};
    const int edgeSide[12][2] = {{0, 0}, {0, 1}, {1, 1}, {1, 0}, {0, 0}, {1, 0}, {1, 1}, {0, 1}, {0, 0}, {1, 0}, {1, 1}, {0, 1}};
    for (int q = 0; q < 4; q++) {
      const int e = faceEdges[iFace][q];
      if (edgeTested & (1u << e)) continue;
      edgeTested |= 1u << e;
      int side[3] = {0, 0, 0};
      side[edgeDir[e][0]] = edgeSide[e][0] ? 1 : -1;
      side[edgeDir[e][1]] = edgeSide[e][1] ? 1 : -1;
      const int nb = cs_neib(m, node, side);
      if (!usable(nb)) continue;
      double d0, d1;
      const bool in0 = near(edgeDir[e][0], edgeSide[e][0], d0), in1 = near(edgeDir[e][1], edgeSide[e][1], d1);
      if (in0 && in1) {
        if (d0 < dmin) dmin = d0, coarser = nb;
        if (d1 < dmin) dmin = d1, coarser = nb;
      }
    }
    const int faceNodeMap[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
    for (int q = 0; q < 4; q++) {
      const int c = faceNodeMap[iFace][q];
      if (cornerTested & (1u << c)) continue;
      cornerTested |= 1u << c;
      const int side[3] = {(c & 1) ? 1 : -1, (c & 2) ? 1 : -1, (c & 4) ? 1 : -1};
      const int nb = cs_neib(m, node, side);
      if (!usable(nb)) continue;
      double dd[3];
      bool in = true;
      for (int d = 0; d < 3; d++) in = near(d, (c >> d) & 1, dd[d]) && in;
      if (in)
        for (int d = 0; d < 3; d++)
          if (dd[d] < dmin) dmin = dd[d], coarser = nb;
    }
  }

  if (coarser >= 0) {
    if (!cs_multiblock(m, x, coarser, S)) return false;
    if ((0.5 < dmin) && (dmin <= 1.0)) {
      BgStencil F;
      // the fine stencil is normalised iff the OUTER stencil's Length != 8 (:903 tests the global StencilTable)
      if (!cs_trilinear(m, loc, x, node, F, &S.n)) return false;
      const double a = 1.0 - (dmin - 0.5) / 0.5, b = (dmin - 0.5) / 0.5;
      for (int e = 0; e < S.n; e++) S.w[e] *= a;
      for (int e = 0; e < F.n; e++) F.w[e] *= b;
      for (int i = 0; i < F.n; i++) {  // cStencilGeneric::Add
        bool found = false;
        for (int j = 0; j < S.n; j++)
          if (S.nd[j] == F.nd[i]) {
            S.w[j] += F.w[i];
            found = true;
            break;
          }
        if (!found) {
          if (S.n == CPLR_MAX_STENCIL) {
            S.overflow = 1;
            break;
          }
          S.w[S.n] = F.w[i], S.nd[S.n] = F.nd[i];
          S.n++;
        }
      }
    }
    return !S.overflow;
  }
  return cs_trilinear(m, loc, x, node, S, &S.n);
}

}  // namespace amps
