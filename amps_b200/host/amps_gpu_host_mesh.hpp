// amps_gpu_host_mesh.hpp -- flattens the reference's AMR mesh (cMeshAMRgeneric / cTreeNodeAMR, src/meshAMR/meshAMRgeneric.h) into the
// amps_gpu_mesh description of the C ABI (include/amps_gpu.h), and walks the unique-node tables back for the scatter of results.
// This is the `UploadMesh()` of INTEGRATION.md as real code: AMPS calls it from DomainBlockDecomposition::UpdateBlockTable
// (pic_mesh.cpp:1644) whenever nMeshModificationCounter changes.
//
//   tree          cTreeNodeAMR: downNode[8], upNode, xmin, xmax, RefinmentLevel, xMinGlobalIndex[3], NodeGeometricSizeIndex, Thread,
//                 IsUsedInCalculationFlag, block                                                     (meshAMRgeneric.h:825-838)
//   lattice       xMinGlobalIndex / NodeGeometricSizeIndex count steps of dx_max_refinment = (xGlobalMax - xGlobalMin) >> _MAX_REFINMENT_LEVEL_
//   unique nodes  one id per physical corner / centre node.  The reference shares cDataCornerNode objects between the blocks that
//                 touch them; here the id comes from integer keys on the finest lattice (corner: imin N + i isize, centre:
//                 2 imin N + (2 i + 1) isize per dimension), so periodic images fold onto one node by a modulo, which is what the
//                 device arrays need (the reference keeps separate copies in its periodic ghost blocks and syncs them)
//   periodic      the reference's periodic mode wraps the user's domain in a shell of ghost blocks
//                 (PIC::BC::ExternalBoundary::Periodic::Init, pic_bc_periodic.cpp:705-747): a leaf that touches the boundary of the
//                 extended domain is a ghost, its real image is BlockPairTable's (:555-571) -- given here by the period on the lattice
//
// Header only, C++17, needs nothing but amps_gpu.h.  The node numbering is pre-order (a node, then its children 0..7), the leaf
// numbering is the order leaves are met -- the same as amps_b200/mesh.py, so both builders give identical descriptions of the same
// tree (tests/test_reference_mesh.py::test_cpp_flattener_matches_the_python_builder runs this header on the reference's own mesh class).
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "amps_gpu.h"

namespace amps_b200 {

struct FlattenOptions {
  int block_cells[3] = {8, 8, 8}, ghost_cells[3] = {1, 1, 1};
  int max_refinement_level = 12;  // _MAX_REFINMENT_LEVEL_
  bool periodic = false;          // _PIC_BC__PERIODIC_MODE_ON_: boundary leaves are ghosts of period user_span on the lattice
  bool all_leaves = false;        // true: describe every leaf (one rank); false: the leaves with an allocated block (node->block)
  int this_rank = 0, n_ranks = 1;
};

struct FlatMesh {
  std::vector<int32_t> node_parent, node_child, node_level, node_imin, node_isize, node_leaf, node_flags, node_thread, root_node;
  std::vector<double> node_xmin, node_xmax;
  std::vector<int32_t> leaf_node, leaf_real, leaf_face_boundary, leaf_corner_uid, leaf_center_uid, leaf_owner, leaf_global_id, global_leaf_to_local;
  std::vector<double> corner_x, center_x;  // position of every unique node [n][3]
  amps_gpu_mesh c{};                       // the view handed to amps_gpu_mesh_upload (points into the vectors above)
  int n_corner_local = 0, n_center_local = 0;

  // every (leaf, block-local corner number incl. ghost layers) of the leaves this rank deposits into, with its unique id: the
  // scatter of J[uid][3], M[uid][243] into the reference's corner buffers is  f(leaf, local, uid)  ->
  //   BlockTable[leaf]->block->GetCornerNode(local)->GetAssociatedDataBufferPointer() + ... JxOffsetIndex / MassMatrixOffsetIndex
  template <class F>
  void ForEachOwnCorner(F f) const {
    const int nl = (int)leaf_node.size();
    for (int l = 0; l < nl; l++) {
      if (c.n_ranks > 1 && leaf_owner[l] != c.this_rank) continue;
      for (int q = 0; q < n_corner_local; q++) {
        const int u = leaf_corner_uid[(size_t)l * n_corner_local + q];
        if (u >= 0) f(l, q, u);
      }
    }
  }
};

namespace detail {
struct Key3 {
  long long k[3];
  bool operator<(const Key3 &o) const { return k[2] != o.k[2] ? k[2] < o.k[2] : (k[1] != o.k[1] ? k[1] < o.k[1] : k[0] < o.k[0]); }
  bool operator==(const Key3 &o) const { return k[0] == o.k[0] && k[1] == o.k[1] && k[2] == o.k[2]; }
};
}  // namespace detail

// Mesh: cMeshAMRgeneric<...> (rootTree, xGlobalMin, xGlobalMax, dx_max_refinment, EPS); Node: cTreeNodeAMR<...>
template <class Mesh, class Node>
FlatMesh FlattenMesh(Mesh *mesh, const FlattenOptions &opt) {
  FlatMesh m;
  const int L = opt.max_refinement_level;
  const long long S = 1LL << L;
  const int *N = opt.block_cells, *g = opt.ghost_cells;

  // ---- tree, pre-order ----
  std::vector<Node *> nodes;
  struct Walk {
    FlatMesh &m;
    std::vector<Node *> &nodes;
    int add(Node *n, int parent) {
      const int id = (int)nodes.size();
      nodes.push_back(n);
      m.node_parent.push_back(parent);
      for (int q = 0; q < 8; q++) m.node_child.push_back(-1);
      m.node_level.push_back((int)n->RefinmentLevel);
      for (int d = 0; d < 3; d++) m.node_imin.push_back(n->xMinGlobalIndex[d]), m.node_xmin.push_back(n->xmin[d]), m.node_xmax.push_back(n->xmax[d]);
      m.node_isize.push_back(n->NodeGeometricSizeIndex);
      m.node_flags.push_back(n->IsUsedInCalculationFlag ? AMPS_NODE_USED : 0);
      m.node_thread.push_back(n->Thread);
      for (int q = 0; q < 8; q++)
        if (n->downNode[q] != NULL) {
          const int c = add(n->downNode[q], id);
          m.node_child[(size_t)8 * id + q] = c;
        }
      return id;
    }
  } walk{m, nodes};
  m.root_node.push_back(walk.add(mesh->rootTree, -1));
  const int nNodes = (int)nodes.size();

  // ---- leaves ----
  std::vector<int> gleaf;  // every leaf of the tree, in node order
  for (int n = 0; n < nNodes; n++) {
    bool leaf = true;
    for (int q = 0; q < 8; q++) leaf = leaf && m.node_child[(size_t)8 * n + q] < 0;
    if (leaf) gleaf.push_back(n);
  }
  m.node_leaf.assign((size_t)nNodes, -1);
  m.global_leaf_to_local.assign(gleaf.size(), -1);
  for (size_t gi = 0; gi < gleaf.size(); gi++) {
    const int n = gleaf[gi];
    if (!opt.all_leaves && nodes[n]->block == NULL) continue;
    m.node_leaf[n] = (int)m.leaf_node.size();
    m.global_leaf_to_local[gi] = (int)m.leaf_node.size();
    m.leaf_node.push_back(n);
    m.leaf_global_id.push_back((int)gi);
    m.leaf_owner.push_back(nodes[n]->Thread);
  }
  const int nLeaves = (int)m.leaf_node.size();

  // ---- boundary faces, periodic ghosts and their real images ----
  m.leaf_face_boundary.assign((size_t)nLeaves, 0);
  m.leaf_real.assign((size_t)nLeaves, -1);
  auto find_leaf_ix = [&](long long ix0, long long ix1, long long ix2) -> int {  // host mirror of findTreeNode on the lattice
    const long long ix[3] = {ix0, ix1, ix2};
    for (int d = 0; d < 3; d++)
      if (ix[d] < 0 || ix[d] >= S) return -1;
    int n = m.root_node[0];
    while (m.node_child[(size_t)8 * n] >= 0) {
      const long long h = m.node_isize[n] / 2;
      int o[3];
      for (int d = 0; d < 3; d++) o[d] = (ix[d] - m.node_imin[(size_t)3 * n + d] < h) ? 0 : 1;
      n = m.node_child[(size_t)8 * n + o[0] + 2 * (o[1] + 2 * o[2])];
    }
    return m.node_leaf[n];
  };
  // periodic mode: the user's domain is the extended one minus a shell of the coarsest boundary blocks; its period on the lattice
  long long shell[3] = {0, 0, 0}, period[3] = {S, S, S};
  if (opt.periodic) {
    for (int d = 0; d < 3; d++) {
      long long sh = 0;
      for (int l = 0; l < nLeaves; l++) {
        const int n = m.leaf_node[l];
        if (m.node_imin[(size_t)3 * n + d] == 0) sh = std::max<long long>(sh, m.node_isize[n]);
      }
      shell[d] = sh, period[d] = S - 2 * sh;
    }
  }
  for (int l = 0; l < nLeaves; l++) {
    const int n = m.leaf_node[l];
    int face = 0;
    for (int d = 0; d < 3; d++) {
      if (m.node_imin[(size_t)3 * n + d] == 0) face |= 1 << (2 * d);
      if ((long long)m.node_imin[(size_t)3 * n + d] + m.node_isize[n] == S) face |= 1 << (2 * d + 1);
    }
    m.leaf_face_boundary[l] = face;
    if (opt.periodic && face) {
      m.node_flags[n] |= AMPS_NODE_PERIODIC_GHOST;
      long long c[3];
      for (int d = 0; d < 3; d++) {
        c[d] = m.node_imin[(size_t)3 * n + d] + m.node_isize[n] / 2;
        c[d] = ((c[d] - shell[d]) % period[d] + period[d]) % period[d] + shell[d];  // findCorrespondingRealBlock, pic_bc_periodic.cpp:502-519
      }
      m.leaf_real[l] = find_leaf_ix(c[0], c[1], c[2]);
    }
  }

  // ---- unique corner / centre nodes by lattice keys ----
  bool amr = false;
  for (int l = 0; l < nLeaves; l++) amr = amr || m.node_level[m.leaf_node[l]] != m.node_level[m.leaf_node[0]];
  auto uid_table = [&](bool corner, std::vector<int32_t> &uid, std::vector<double> &xs) -> int {
    int ext[3];
    for (int d = 0; d < 3; d++) ext[d] = N[d] + 2 * g[d] + (corner ? 1 : 0);
    const int nloc = ext[0] * ext[1] * ext[2];
    (corner ? m.n_corner_local : m.n_center_local) = nloc;
    std::vector<detail::Key3> key((size_t)nLeaves * nloc);
    std::vector<char> valid((size_t)nLeaves * nloc, 1), inside((size_t)nloc, 1);
    long long span[3], org[3];
    for (int d = 0; d < 3; d++) {
      span[d] = (corner ? 1 : 2) * (opt.periodic ? period[d] : S) * N[d];
      org[d] = (corner ? 1 : 2) * shell[d] * N[d];
    }
    for (int l = 0; l < nLeaves; l++) {
      const int n = m.leaf_node[l];
      const bool has = opt.n_ranks <= 1 || m.leaf_owner[l] == opt.this_rank;  // node tables only for the leaves that hold particles here
      int q = 0;
      for (int k = -g[2]; k < ext[2] - g[2]; k++)
        for (int j = -g[1]; j < ext[1] - g[1]; j++)
          for (int i = -g[0]; i < ext[0] - g[0]; i++, q++) {
            const int loc[3] = {i, j, k};
            detail::Key3 kk;
            bool ok = has, in = true;
            for (int d = 0; d < 3; d++) {
              const long long im = m.node_imin[(size_t)3 * n + d], sz = m.node_isize[n];
              long long v = corner ? im * N[d] + (long long)loc[d] * sz : 2 * im * N[d] + (2LL * loc[d] + 1) * sz;
              if (opt.periodic) v = ((v - org[d]) % span[d] + span[d]) % span[d];
              else if (v < 0 || v > span[d]) ok = false;
              kk.k[d] = v;
              in = in && loc[d] >= 0 && loc[d] < N[d] + (corner ? 1 : 0);
            }
            key[(size_t)l * nloc + q] = kk;
            valid[(size_t)l * nloc + q] = ok ? 1 : 0;
            if (l == 0) inside[q] = in ? 1 : 0;
          }
    }
    // the pool of nodes: single level on one rank -- the nodes blocks really own (ghost-layer positions only resolve to such nodes);
    // otherwise every node an own block's tile can touch (next to a coarser / finer block the ghost cells are nodes of their own)
    std::vector<detail::Key3> pool;
    const bool ownOnly = opt.n_ranks <= 1 && !amr;
    for (int l = 0; l < nLeaves; l++)
      for (int q = 0; q < nloc; q++)
        if (valid[(size_t)l * nloc + q] && (!ownOnly || inside[q])) pool.push_back(key[(size_t)l * nloc + q]);
    std::sort(pool.begin(), pool.end());
    pool.erase(std::unique(pool.begin(), pool.end()), pool.end());
    uid.assign((size_t)nLeaves * nloc, -1);
    for (size_t e = 0; e < key.size(); e++) {
      if (!valid[e]) continue;
      auto it = std::lower_bound(pool.begin(), pool.end(), key[e]);
      if (it != pool.end() && *it == key[e]) uid[e] = (int32_t)(it - pool.begin());
    }
    xs.assign(pool.size() * 3, 0.0);
    for (size_t u = 0; u < pool.size(); u++)
      for (int d = 0; d < 3; d++) {
        const double den = (corner ? 1.0 : 2.0) * (double)S * N[d];
        const double x0 = mesh->xGlobalMin[d] + (opt.periodic ? (double)shell[d] / (double)S * (mesh->xGlobalMax[d] - mesh->xGlobalMin[d]) : 0.0);
        xs[3 * u + d] = x0 + (double)pool[u].k[d] / den * (mesh->xGlobalMax[d] - mesh->xGlobalMin[d]);
      }
    return (int)pool.size();
  };
  const int nCorners = uid_table(true, m.leaf_corner_uid, m.corner_x);
  const int nCenters = uid_table(false, m.leaf_center_uid, m.center_x);

  // ---- the C view ----
  amps_gpu_mesh &c = m.c;
  for (int d = 0; d < 3; d++) {
    c.n_root[d] = 1;
    c.x_global_min[d] = mesh->xGlobalMin[d], c.x_global_max[d] = mesh->xGlobalMax[d];
    c.dx_max_refinement[d] = mesh->dx_max_refinment[d];
    c.dx_root_block[d] = mesh->xGlobalMax[d] - mesh->xGlobalMin[d];
  }
  c.max_refinement_level = L;
  c.eps = mesh->EPS;
  c.n_nodes = nNodes, c.n_leaves = nLeaves, c.n_corners = nCorners, c.n_centers = nCenters;
  c.node_parent = m.node_parent.data(), c.node_child = m.node_child.data(), c.node_level = m.node_level.data();
  c.node_imin = m.node_imin.data(), c.node_isize = m.node_isize.data(), c.node_xmin = m.node_xmin.data(), c.node_xmax = m.node_xmax.data();
  c.node_leaf = m.node_leaf.data(), c.node_flags = m.node_flags.data(), c.node_thread = m.node_thread.data(), c.root_node = m.root_node.data();
  c.leaf_node = m.leaf_node.data(), c.leaf_real = m.leaf_real.data(), c.leaf_face_boundary = m.leaf_face_boundary.data();
  c.leaf_corner_uid = m.leaf_corner_uid.data(), c.leaf_center_uid = m.leaf_center_uid.data();
  c.this_rank = opt.this_rank, c.n_ranks = opt.n_ranks, c.n_global_leaves = (int)gleaf.size();
  c.leaf_owner = m.leaf_owner.data(), c.leaf_global_id = m.leaf_global_id.data(), c.global_leaf_to_local = m.global_leaf_to_local.data();
  return m;
}

}  // namespace amps_b200
