// C++ driver of amps_b200/host/amps_gpu_host.hpp: reads a case written by tests/test_cpp_host.py (configuration, flattened mesh,
// an AMPS-layout AoS particle buffer with its cell lists, fields), runs MoveParticles + UpdateJMassMatrix + DownloadParticles
// through the C++ host layer and writes the results back for comparison with the CPU oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "amps_gpu_host.hpp"

static std::vector<unsigned char> read_blob(FILE *f) {
  int64_t n = 0;
  if (fread(&n, 8, 1, f) != 1) {
    fprintf(stderr, "short read\n");
    exit(2);
  }
  std::vector<unsigned char> b((size_t)n);
  if (n && fread(b.data(), 1, (size_t)n, f) != (size_t)n) {
    fprintf(stderr, "short read\n");
    exit(2);
  }
  return b;
}
static void write_blob(FILE *f, const void *p, int64_t n) {
  fwrite(&n, 8, 1, f);
  if (n) fwrite(p, 1, (size_t)n, f);
}

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  auto cfgb = read_blob(f);
  if (cfgb.size() != sizeof(amps_gpu_config)) {
    fprintf(stderr, "config size %zu != %zu\n", cfgb.size(), sizeof(amps_gpu_config));
    return 2;
  }
  amps_gpu_config cfg;
  memcpy(&cfg, cfgb.data(), sizeof(cfg));
  auto scal = read_blob(f);  // n_root[3], L, n_nodes, n_leaves, n_corners, n_centers (int32) then 13 doubles
  amps_gpu_mesh mesh;
  memset(&mesh, 0, sizeof(mesh));
  const int32_t *si = (const int32_t *)scal.data();
  const double *sd = (const double *)(scal.data() + 8 * 4);
  for (int d = 0; d < 3; d++) mesh.n_root[d] = si[d];
  mesh.max_refinement_level = si[3], mesh.n_nodes = si[4], mesh.n_leaves = si[5], mesh.n_corners = si[6], mesh.n_centers = si[7];
  for (int d = 0; d < 3; d++)
    mesh.x_global_min[d] = sd[d], mesh.x_global_max[d] = sd[3 + d], mesh.dx_max_refinement[d] = sd[6 + d], mesh.dx_root_block[d] = sd[9 + d];
  mesh.eps = sd[12];
  std::vector<std::vector<unsigned char>> arr;
  for (int i = 0; i < 16; i++) arr.push_back(read_blob(f));
  mesh.node_parent = (const int32_t *)arr[0].data(), mesh.node_child = (const int32_t *)arr[1].data();
  mesh.node_level = (const int32_t *)arr[2].data(), mesh.node_imin = (const int32_t *)arr[3].data();
  mesh.node_isize = (const int32_t *)arr[4].data(), mesh.node_xmin = (const double *)arr[5].data();
  mesh.node_xmax = (const double *)arr[6].data(), mesh.node_leaf = (const int32_t *)arr[7].data();
  mesh.node_flags = (const int32_t *)arr[8].data(), mesh.node_thread = (const int32_t *)arr[9].data();
  mesh.root_node = (const int32_t *)arr[10].data(), mesh.leaf_node = (const int32_t *)arr[11].data();
  mesh.leaf_real = (const int32_t *)arr[12].data(), mesh.leaf_face_boundary = (const int32_t *)arr[13].data();
  mesh.leaf_corner_uid = (const int32_t *)arr[14].data(), mesh.leaf_center_uid = (const int32_t *)arr[15].data();
  mesh.this_rank = 0, mesh.n_ranks = 1;

  auto layb = read_blob(f);
  amps_b200::ParticleBufferView pb;
  memcpy(&pb.layout, layb.data(), sizeof(pb.layout));
  auto buffer = read_blob(f);
  pb.ParticleDataBuffer = buffer.data();
  pb.MaxNPart = (long int)(buffer.size() / pb.layout.stride);
  auto firstb = read_blob(f);
  std::vector<long int> first(firstb.size() / 8);
  memcpy(first.data(), firstb.data(), firstb.size());
  auto E = read_blob(f), Bp = read_blob(f), Bc = read_blob(f);
  fclose(f);

  try {
    amps_b200::EcsimHost host(cfg, mesh);
    if ((int64_t)first.size() != host.n_cells()) throw std::runtime_error("cell table size");
    host.SetFields((const double *)E.data(), (const double *)Bp.data(), (const double *)Bc.data());
    host.UploadParticles(pb, first.data());
    amps_gpu_move_stats st = host.MoveParticles();
    std::vector<double> J((size_t)mesh.n_corners * 3), M((size_t)mesh.n_corners * 243), cfl(AMPS_GPU_MAX_SPECIES, 0.0);
    double energy = 0.0;
    host.UpdateJMassMatrix(J.data(), M.data(), &energy, cfl.data());
    // the other particle passes of the host layer on the moved, re-filed store
    std::vector<double> rho((size_t)mesh.n_centers);
    host.ComputeNetCharge(0.7, rho.data());
    std::vector<double> sample((size_t)host.n_cells() * cfg.n_species * 13);
    std::vector<int64_t> nSampled(AMPS_GPU_MAX_SPECIES, 0);
    host.Sampling();
    host.SampledData(sample.data(), nSampled.data(), true);
    // wipe the lists so that the download provably rebuilds them
    std::fill(first.begin(), first.end(), -7L);
    const int64_t n = host.DownloadParticles(pb, first.data());

    FILE *o = fopen(argv[2], "wb");
    write_blob(o, &st, sizeof(st));
    write_blob(o, &n, 8);
    write_blob(o, buffer.data(), (int64_t)buffer.size());
    write_blob(o, first.data(), (int64_t)first.size() * 8);
    write_blob(o, J.data(), (int64_t)J.size() * 8);
    write_blob(o, M.data(), (int64_t)M.size() * 8);
    write_blob(o, &energy, 8);
    write_blob(o, cfl.data(), (int64_t)cfl.size() * 8);
    write_blob(o, rho.data(), (int64_t)rho.size() * 8);
    write_blob(o, sample.data(), (int64_t)sample.size() * 8);
    write_blob(o, nSampled.data(), (int64_t)nSampled.size() * 8);
    fclose(o);
  } catch (const std::exception &e) {
    fprintf(stderr, "host_roundtrip: %s\n", e.what());
    return 1;
  }
  printf("HOST_ROUNDTRIP_OK\n");
  return 0;
}
