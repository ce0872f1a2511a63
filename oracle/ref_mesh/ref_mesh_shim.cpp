// ref_mesh_shim.cpp -- TEST INFRASTRUCTURE.  Compiles the reference's OWN AMR mesh class (src/meshAMR/meshAMR3d.h ->
// meshAMRgeneric.h, header-only templates) from the sources where they lie under /root/reference, with three stand-ins that are
// ours: a single-process <mpi.h> (mpi.h here), an empty generated-configuration include (.general.conf, UserDefinition.meshAMR.h)
// and empty user-data classes (ref_stubs.h).  Built into oracle/_ref/libref_mesh.so by oracle/Makefile; used by
// tests/test_reference_mesh.py to pin the restated tree search (a14), the neighbour probes and the neighbour level limits (a4)
// against the reference implementation itself.  No reference source is copied.
//
// Compile-time settings are the reference defaults of meshAMRdef.h: 5x5x5 cells per block, 2 ghost layers,
// _MAX_REFINMENT_LEVEL_ 15 -- the configuration of input/gca_mover.input.
#include "meshAMR3d.h"

#include <cstring>

#include "../../amps_b200/host/amps_gpu_host_mesh.hpp"  // the PRODUCT's mesh flattener, run here on the reference's own mesh class

// ---- globals / functions the headers expect from other translation units of AMPS ----
MPI_Comm MPI_GLOBAL_COMMUNICATOR = 0;
int ThisThread = 0, TotalThreadsNumber = 1;
double _MESH_AMR_XMAX_[3], _MESH_AMR_XMIN_[3];
int cInternalRotationBodyData::nAxisSurfaceElements = 0;
namespace CutCell {
cTriangleFace *BoundaryTriangleFaces = NULL;
int nBoundaryTriangleFaces = 0;
cAMRstack<cTriangleFaceDescriptor> BoundaryTriangleFaceDescriptor;
}  // namespace CutCell
void exit(long int nline, const char *fname, const char *msg) {
  fprintf(stderr, "reference exit(): %s, line %ld: %s\n", fname, nline, msg ? msg : "");
  abort();
}
extern "C" {
int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
int MPI_Initialized(int *f) { *f = 1; return 0; }
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static size_t tsize(MPI_Datatype t) { return (t == MPI_BYTE || t == MPI_CHAR || t == MPI_UNSIGNED_CHAR) ? 1 : (t == MPI_INT || t == MPI_UNSIGNED || t == MPI_FLOAT) ? 4 : 8; }
int MPI_Gather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, int, MPI_Comm) { if (s != MPI_IN_PLACE) memcpy(r, s, n * tsize(t)); return 0; }
int MPI_Allgather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, MPI_Comm) { if (s != MPI_IN_PLACE) memcpy(r, s, n * tsize(t)); return 0; }
int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { abort(); }
int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { abort(); }
}

typedef cBasicBlockAMR<cBasicCornerNode, cBasicCenterNode> Block;
typedef cMeshAMR3d<cBasicCornerNode, cBasicCenterNode, Block> Mesh;
typedef cTreeNodeAMR<Block> Node;
static Mesh *mesh = nullptr;
static double g_center[3], g_radii[8], g_dx0;
static int g_nlev;
// requested cell size: dx0 outside every sphere, halved inside each nested sphere radii[l] about the domain centre
static double local_resolution(double *x) {
  double r = 0;
  for (int d = 0; d < 3; d++) r += (x[d] - g_center[d]) * (x[d] - g_center[d]);
  r = sqrt(r);
  double dx = g_dx0;
  for (int l = 0; l < g_nlev; l++)
    if (r < g_radii[l]) dx = g_dx0 / (2 << l);
  return dx;
}
static void fill(Node *n, double *lo, double *hi, int *level) {
  for (int d = 0; d < 3; d++) lo[d] = n->xmin[d], hi[d] = n->xmax[d];
  *level = n->RefinmentLevel;
}

extern "C" {
int ref_mesh_block_cells() { return _BLOCK_CELLS_X_; }
int ref_mesh_ghost_cells() { return _GHOST_CELLS_X_; }
int ref_mesh_max_refinement_level() { return _MAX_REFINMENT_LEVEL_; }

// cMeshAMRgeneric::init + buildMesh (meshAMRgeneric.h:2328-2400)
int ref_mesh_build(const double *xmin, const double *xmax, double dx0, int nlev, const double *radii) {
  double a[3], b[3];
  for (int d = 0; d < 3; d++) a[d] = xmin[d], b[d] = xmax[d], g_center[d] = 0.5 * (xmin[d] + xmax[d]);
  g_dx0 = dx0, g_nlev = nlev;
  for (int l = 0; l < nlev; l++) g_radii[l] = radii[l];
  mesh = new Mesh();
  mesh->AllowBlockAllocation = false;
  mesh->init(a, b, local_resolution);
  mesh->buildMesh();
  return 0;
}
double ref_mesh_eps() { return mesh->EPS; }
void ref_mesh_dx_max_refinement(double *dx) { for (int d = 0; d < 3; d++) dx[d] = mesh->dx_max_refinment[d]; }

// findTreeNode(double*), meshAMRgeneric.h:2851-2882
int ref_find_tree_node(const double *x, double *lo, double *hi, int *level) {
  double xx[3] = {x[0], x[1], x[2]};
  Node *n = mesh->findTreeNode(xx, NULL);
  if (!n) return -1;
  fill(n, lo, hi, level);
  return 0;
}
// FindCellIndex, meshAMRgeneric.h:2256-2323 in the leaf that contains x; returns the local cell number or -1
long int ref_find_cell_index(const double *x, int *ijk) {
  double xx[3] = {x[0], x[1], x[2]};
  Node *n = mesh->findTreeNode(xx, NULL);
  if (!n) return -2;
  int i, j, k;
  long int nd = mesh->FindCellIndex(xx, i, j, k, n, false);
  ijk[0] = i, ijk[1] = j, ijk[2] = k;
  return nd;
}
// neighbours of the leaf that contains x: kind 0 = GetNeibFace(idx,0,0), 1 = GetNeibEdge(idx,0), 2 = GetNeibCorner(idx)
int ref_neib(const double *x, int kind, int idx, double *lo, double *hi, int *level) {
  double xx[3] = {x[0], x[1], x[2]};
  Node *n = mesh->findTreeNode(xx, NULL);
  if (!n) return -2;
  Node *nb = (kind == 0) ? n->GetNeibFace(idx, 0, 0, mesh) : (kind == 1) ? n->GetNeibEdge(idx, 0, mesh) : n->GetNeibCorner(idx, mesh);
  if (!nb) return -1;
  fill(nb, lo, hi, level);
  return 0;
}
// SetNeibRefinmentLevelLimits (meshAMRgeneric.h:1018-1048) of the leaf that contains x
int ref_neib_levels(const double *x, int *minmax) {
  double xx[3] = {x[0], x[1], x[2]};
  Node *n = mesh->findTreeNode(xx, NULL);
  if (!n) return -2;
  n->SetNeibRefinmentLevelLimits(mesh);
  minmax[0] = n->minNeibRefinmentLevel, minmax[1] = n->maxNeibRefinmentLevel;
  return 0;
}
// amps_b200::FlattenMesh (amps_b200/host/amps_gpu_host_mesh.hpp) on the reference's mesh: sizes first (out == NULL), then the arrays
static amps_b200::FlatMesh g_flat;
int ref_flatten(int block_cells, int ghost_cells, int *sizes) {
  amps_b200::FlattenOptions o;
  for (int d = 0; d < 3; d++) o.block_cells[d] = block_cells, o.ghost_cells[d] = ghost_cells;
  o.max_refinement_level = _MAX_REFINMENT_LEVEL_;
  o.all_leaves = true;
  g_flat = amps_b200::FlattenMesh<Mesh, Node>(mesh, o);
  sizes[0] = g_flat.c.n_nodes, sizes[1] = g_flat.c.n_leaves, sizes[2] = g_flat.c.n_corners, sizes[3] = g_flat.c.n_centers;
  sizes[4] = g_flat.n_corner_local, sizes[5] = g_flat.n_center_local;
  return 0;
}
void ref_flatten_arrays(int *node_child, int *node_level, int *node_imin, int *node_isize, double *node_xmin, double *node_xmax, int *leaf_node,
                        int *leaf_face, int *corner_uid, int *center_uid, double *corner_x, double *center_x) {
  auto cp = [](auto *dst, const auto &v) { memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
  cp(node_child, g_flat.node_child), cp(node_level, g_flat.node_level), cp(node_imin, g_flat.node_imin), cp(node_isize, g_flat.node_isize);
  cp(node_xmin, g_flat.node_xmin), cp(node_xmax, g_flat.node_xmax), cp(leaf_node, g_flat.leaf_node), cp(leaf_face, g_flat.leaf_face_boundary);
  cp(corner_uid, g_flat.leaf_corner_uid), cp(center_uid, g_flat.leaf_center_uid), cp(corner_x, g_flat.corner_x), cp(center_x, g_flat.center_x);
}
// helpers of src/general/specfunc.h that the restated movers use
double ref_gyro_frequency(const double *v, double m, double q, const double *B) {
  double vv[3] = {v[0], v[1], v[2]}, bb[3] = {B[0], B[1], B[2]};
  return ::Relativistic::GetGyroFrequency(vv, m, q, bb);
}
void ref_normalize(double *x) { Vector3D::Normalize(x); }
double ref_speed_of_light() { return SpeedOfLight; }
}
