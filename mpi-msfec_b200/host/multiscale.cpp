#include "multiscale.h"

#include <sys/stat.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

namespace msfec {

namespace {
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int env_int(const char *a, const char *b, const char *c, int dflt) {
  for (const char *n : {a, b, c})
    if (n) if (const char *v = std::getenv(n)) return std::atoi(v);
  return dflt;
}
const char *block_names(int pairing, int blk) {
  if (pairing == MSFEC_Q) return "u";
  return blk == 0 ? "sigma" : "u";
}
}  // namespace

template <int PAIRING>
Multiscale<PAIRING>::Multiscale(const ParametersMs &prm, const std::string &prm_file, int rank, int world, int device, const char *name,
                                msfec_comm *shared_comm)
    : parameters(prm), parameter_filename(prm_file), rank_(rank), world_(world), device_(device), name_(name), comm_(shared_comm) {
  // NCCL communicator for the two cross-rank steps; a single rank needs none (MSFEC_NCCL=1 creates it anyway)
  if (!comm_ && (world_ > 1 || (std::getenv("MSFEC_NCCL") && std::atoi(std::getenv("MSFEC_NCCL")) != 0))) {
    if (msfec_comm_create(rank_, world_, device_, &comm_)) throw std::runtime_error(std::string("NCCL communicator: ") + msfec_comm_last_error());
    owns_comm_ = true;
    if (rank_ == 0) std::cout << "NCCL communicator over " << world_ << " rank(s) created (one rank per GPU)." << std::endl;
  }
}

template <int PAIRING>
Multiscale<PAIRING>::~Multiscale() { if (owns_comm_) msfec_comm_destroy(comm_); }

// hyper_cube + refine_global; ownership = contiguous chunks of the z-order (ned_rt_global.cc:36-46, 12-15)
template <int PAIRING>
void Multiscale<PAIRING>::make_grid() {
  n_global_cells_ = 1LL << (3 * parameters.n_refine_global);
  owned_range(n_global_cells_, rank_, world_, lo_, hi_);
  first_cell_ = CellId(0, parameters.n_refine_global);     // "first cell" of the whole mesh: rank 0's first (ned_rt_global.cc:66-70)
}

// ned_rt_global.cc:49-99
template <int PAIRING>
void Multiscale<PAIRING>::initialize_and_compute_basis() {
  const double t0 = now();
  batch_ = std::make_shared<BasisBatch>(parameters, device_);
  for (long long id = lo_; id < hi_; ++id) {
    const CoarseCell cell(parameters.n_refine_global, id);
    Basis current_cell_problem(parameters, parameter_filename, cell, first_cell_, (unsigned)rank_, batch_);
    cell_basis_map.emplace(cell.id, current_cell_problem);   // stored by value: copied before run(), as in the reference
  }
  for (auto &kv : cell_basis_map) kv.second.run();
  t_basis_ = now() - t0;
  const msfec_stats &st = batch_->stats();
  const char *solver = st.solver == 3 ? "none (0 local refinements)" : st.solver == 2 ? "multifrontal LDL^T" : st.solver == 1 ? "banded block LDL^T" : "MINRES";
  std::cout << "[rank " << rank_ << "] " << name_ << " basis initialization and computation: " << (hi_ - lo_) << " cells in " << t_basis_
            << " s (device " << st.ms_total << " ms; assemble " << st.ms_assemble << ", lift " << st.ms_lift << ", solve "
            << st.ms_solve << ", coarse matrices " << st.ms_gram << "; solver " << solver << ", max its " << st.iterations_max
            << ", residual " << st.residual_max << ", " << st.kernel_launches << " kernel launches)" << std::endl;
}

// ned_rt_global.cc:100-160
template <int PAIRING>
void Multiscale<PAIRING>::setup_system_matrix() {
  coarse_.reset(new CoarseProblem(PAIRING, parameters.n_refine_global));
  // the nested coarse iteration runs on this rank's GPU (MSFEC_COARSE_SOLVER=host keeps it on the host cores;
  // MSFEC_COARSE_DENSE_LIMIT: unknowns up to which the exact dense LU is used instead, default 6000)
  {
    const char *where = std::getenv("MSFEC_COARSE_SOLVER");
    if (where && std::string(where) != "host" && std::string(where) != "device")
      throw std::invalid_argument("MSFEC_COARSE_SOLVER must be host or device");
    if (!where || std::string(where) == "device") coarse_->set_device(device_);
    if (const char *lim = std::getenv("MSFEC_COARSE_DENSE_LIMIT")) coarse_->set_dense_limit(std::atoi(lim));
  }
  if (rank_ == 0) {
    std::cout << "Number of active cells: " << n_global_cells_ << std::endl
              << "Total number of cells: " << ((8 * n_global_cells_ - 1) / 7) << " (on " << parameters.n_refine_global + 1 << " levels)" << std::endl
              << "Number of degrees of freedom: " << coarse_->n_dofs();
    if (coarse_->n_block1()) std::cout << " (" << coarse_->n_block0() << '+' << coarse_->n_block1() << ')';
    std::cout << std::endl;
  }
  setup_constraints();
}

// Homogeneous essential data on the boundary for Q (q_global.cc:127-139) and Q_Ned (q_ned_global.cc:175-207); none for
// Ned_RT / RT_DQ (ned_rt_global.cc:162-190).  The coarse mesh has no hanging nodes.  CoarseProblem holds the mask.
template <int PAIRING>
void Multiscale<PAIRING>::setup_constraints() {}

// ned_rt_global.cc:192-317.  The reference adds each rank's element matrices to a distributed Trilinos matrix; here the
// ranks exchange the element matrices of their cells (ONE ncclAllGather) and every rank assembles the whole coarse system.
template <int PAIRING>
void Multiscale<PAIRING>::assemble_system() {
  const double t0 = now();
  const int k = batch_->k();
  const size_t per = (size_t)k * k + k;
  long long max_chunk = 0;
  for (int r = 0; r < world_; ++r) { long long a, b; owned_range(n_global_cells_, r, world_, a, b); max_chunk = std::max(max_chunk, b - a); }
  std::vector<double> mine((size_t)max_chunk * per, 0.0);
  {
    size_t o = 0;
    for (auto &kv : cell_basis_map) {
      const FullMatrix &M = kv.second.get_global_element_matrix();
      const Vector &r = kv.second.get_global_element_rhs();
      std::copy(M.data, M.data + (size_t)k * k, mine.begin() + o);
      std::copy(r.data, r.data + k, mine.begin() + o + (size_t)k * k);
      o += per;
    }
  }
  std::vector<double> gathered;
  if (comm_) {
    gathered.resize((size_t)world_ * mine.size());
    if (msfec_comm_allgather(comm_, mine.data(), mine.size(), gathered.data())) throw std::runtime_error(std::string("ncclAllGather: ") + msfec_comm_last_error());
  } else {
    gathered = mine;
  }
  for (int r = 0; r < world_; ++r) {
    long long a, b;
    owned_range(n_global_cells_, r, world_, a, b);
    const double *base = gathered.data() + (size_t)r * mine.size();
    for (long long id = a; id < b; ++id) {
      const double *e = base + (size_t)(id - a) * per;
      coarse_->add_cell(id, e, e + (size_t)k * k);
    }
  }
  t_assemble_ = now() - t0;
}

// ned_rt_global.cc:330-461 (Schur-complement CG); Q: q_global.cc:327-363 (CG)
template <int PAIRING>
void Multiscale<PAIRING>::solve_iterative() {
  const double t0 = now();
  const std::string info = coarse_->solve();
  t_solve_ = now() - t0;
  if (rank_ == 0) std::cout << "   Coarse solver: " << info << "." << std::endl << "   Outer solver completed." << std::endl;
}

// ned_rt_global.cc:464-488
template <int PAIRING>
void Multiscale<PAIRING>::send_global_weights_to_cell() {
  const int dofs_per_cell = batch_->k();
  std::vector<double> extracted_weights(dofs_per_cell, 0);
  for (auto &kv : cell_basis_map) {
    coarse_->cell_weights(kv.first.index, extracted_weights.data());
    kv.second.set_global_weights(extracted_weights);
  }
}

// harness-defined norms of the multiscale solution (the reference computes none): per-cell squared norms on the device,
// summed over the rank's cells, then ONE ncclAllReduce(sum, FP64) over the ranks
template <int PAIRING>
void Multiscale<PAIRING>::compute_norms() {
  std::array<double, 4> s = batch_->solution_norms_squared();
  if (comm_ && msfec_comm_allreduce_sum(comm_, s.data(), 4)) throw std::runtime_error(std::string("ncclAllReduce: ") + msfec_comm_last_error());
  for (int i = 0; i < 4; ++i) norms_[i] = std::sqrt(s[i]);
  if (rank_ == 0) {
    const char *semi0 = PAIRING == MSFEC_NED_RT ? "H(curl)" : PAIRING == MSFEC_RT_DQ ? "H(div)" : "H1";
    const char *semi1 = PAIRING == MSFEC_NED_RT ? "H(div)" : "H(curl)";
    std::printf("   %s solution norms: ||%s||_L2 = %.15e  |%s|_%s = %.15e", parameters.standard ? "Standard" : "Multiscale", block_names(PAIRING, 0), norms_[0], block_names(PAIRING, 0), semi0, norms_[1]);
    if (PAIRING != MSFEC_Q) {
      std::printf("  ||u||_L2 = %.15e", norms_[2]);
      if (PAIRING != MSFEC_RT_DQ) std::printf("  |u|_%s = %.15e", semi1, norms_[3]);
    }
    std::printf("\n");
    std::fflush(stdout);
  }
}

// ned_rt_global.cc:557-567
template <int PAIRING>
std::vector<std::string> Multiscale<PAIRING>::collect_filenames_on_mpi_process() {
  std::vector<std::string> filename_list;
  for (auto &kv : cell_basis_map) filename_list.push_back(kv.second.get_filename_global());
  return filename_list;
}

// ned_rt_global.cc:571-700: per-cell fine solution files, one coarse-level file per rank, two .pvtu records on rank 0
template <int PAIRING>
void Multiscale<PAIRING>::output_results() {
  const double t0 = now();
  ::mkdir(parameters.dirname_output.c_str(), 0755);
  // MSFEC_MAX_OUTPUT_CELLS limits the number of per-cell files a rank writes (the reference writes all of them)
  long long limit = std::getenv("MSFEC_MAX_OUTPUT_CELLS") ? std::atoll(std::getenv("MSFEC_MAX_OUTPUT_CELLS")) : -1;
  long long written = 0;
  // the fine-grid comparator has no fine grid below its cells: one file per rank and the .pvtu record (ned_rt_ref.cc:699-730)
  if (parameters.standard) limit = 0;
  for (auto &kv : cell_basis_map) {
    if (limit >= 0 && written >= limit) break;
    kv.second.output_global_solution_in_cell();
    ++written;
  }
  const int g = parameters.n_refine_global, k = batch_->k();
  const std::string stem = parameters.filename_output + "_n_refine-" + int_to_string(g, 2);
  const bool two = PAIRING != MSFEC_Q;
  {
    // coarse-level file of this rank: owned coarse cells, the coarse solution evaluated at the cell centres with the
    // standard lowest-order shape functions, and the subdomain id
    std::ofstream f(parameters.dirname_output + "/" + stem + "." + int_to_string(rank_, 4) + ".vtu");
    f.precision(12);
    const long long nc = hi_ - lo_;
    f << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n"
      << "<Piece NumberOfPoints=\"" << 8 * nc << "\" NumberOfCells=\"" << nc << "\">\n<Points>\n<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n";
    static const int vtk_order[8] = {0, 1, 3, 2, 4, 5, 7, 6};
    for (long long id = lo_; id < hi_; ++id) {
      const CoarseCell c(g, id);
      for (int v = 0; v < 8; ++v) f << c.vertices[vtk_order[v]][0] << ' ' << c.vertices[vtk_order[v]][1] << ' ' << c.vertices[vtk_order[v]][2] << '\n';
    }
    f << "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">\n";
    for (long long e = 0; e < nc; ++e) { for (int v = 0; v < 8; ++v) f << 8 * e + v << ' '; f << '\n'; }
    f << "</DataArray>\n<DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n";
    for (long long e = 1; e <= nc; ++e) f << 8 * e << '\n';
    f << "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n";
    for (long long e = 0; e < nc; ++e) f << "12\n";
    f << "</DataArray>\n</Cells>\n<CellData>\n";
    const double H = 1.0 / (double)(1LL << g);
    std::vector<double> w(k);
    auto centre_value = [&](int kind, const double *v, double out[3]) {   // kind: 0 vertex, 1 edge, 2 face, 3 cell
      out[0] = out[1] = out[2] = 0;
      if (kind == 0) { for (int i = 0; i < 8; ++i) out[0] += v[i] / 8.0; }
      else if (kind == 1) {
        out[0] = (v[2] + v[3] + v[6] + v[7]) / (4.0 * H); out[1] = (v[0] + v[1] + v[4] + v[5]) / (4.0 * H); out[2] = (v[8] + v[9] + v[10] + v[11]) / (4.0 * H);
      } else if (kind == 2) { for (int d = 0; d < 3; ++d) out[d] = (v[2 * d] + v[2 * d + 1]) / (2.0 * H * H); }
      else out[0] = v[0];
    };
    const int kind0 = PAIRING == MSFEC_Q || PAIRING == MSFEC_Q_NED ? 0 : PAIRING == MSFEC_NED_RT ? 1 : 2;
    const int kind1 = PAIRING == MSFEC_Q_NED ? 1 : PAIRING == MSFEC_NED_RT ? 2 : 3;
    const int l0 = kind0 == 0 ? 8 : kind0 == 1 ? 12 : 6;
    for (int blk = 0; blk < (two ? 2 : 1); ++blk) {
      const int kind = blk == 0 ? kind0 : kind1;
      const bool vec = kind == 1 || kind == 2;
      f << "<DataArray type=\"Float64\" Name=\"" << block_names(PAIRING, blk) << "\"" << (vec ? " NumberOfComponents=\"3\"" : "") << " format=\"ascii\">\n";
      for (long long id = lo_; id < hi_; ++id) {
        coarse_->cell_weights(id, w.data());
        double o[3];
        centre_value(kind, w.data() + (blk ? l0 : 0), o);
        if (vec) f << o[0] << ' ' << o[1] << ' ' << o[2] << '\n'; else f << o[0] << '\n';
      }
      f << "</DataArray>\n";
    }
    f << "<DataArray type=\"Float32\" Name=\"subdomain_id\" format=\"ascii\">\n";
    for (long long e = 0; e < nc; ++e) f << rank_ << '\n';
    f << "</DataArray>\n</CellData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n";
  }
  if (rank_ == 0) {
    auto pvtu = [&](const std::string &path, const std::vector<std::string> &pieces, bool coarse) {
      std::ofstream f(path);
      f << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<PUnstructuredGrid GhostLevel=\"0\">\n";
      const bool nodal0 = PAIRING == MSFEC_Q || PAIRING == MSFEC_Q_NED;
      if (!coarse && nodal0) f << "<PPointData Scalars=\"" << block_names(PAIRING, 0) << "\">\n<PDataArray type=\"Float64\" Name=\"" << block_names(PAIRING, 0) << "\" format=\"ascii\"/>\n</PPointData>\n";
      f << "<PCellData>\n";
      if (coarse) {
        f << "<PDataArray type=\"Float64\" Name=\"" << block_names(PAIRING, 0) << "\"" << (nodal0 ? "" : " NumberOfComponents=\"3\"") << " format=\"ascii\"/>\n";
        if (two) f << "<PDataArray type=\"Float64\" Name=\"u\"" << (PAIRING == MSFEC_RT_DQ ? "" : " NumberOfComponents=\"3\"") << " format=\"ascii\"/>\n";
        f << "<PDataArray type=\"Float32\" Name=\"subdomain_id\" format=\"ascii\"/>\n";
      } else {
        if (PAIRING == MSFEC_Q_NED) f << "<PDataArray type=\"Float64\" Name=\"u\" NumberOfComponents=\"3\" format=\"ascii\"/>\n";
        if (PAIRING == MSFEC_NED_RT) f << "<PDataArray type=\"Float64\" Name=\"sigma\" NumberOfComponents=\"3\" format=\"ascii\"/>\n<PDataArray type=\"Float64\" Name=\"u\" NumberOfComponents=\"3\" format=\"ascii\"/>\n<PDataArray type=\"Float64\" Name=\"div_u\" format=\"ascii\"/>\n";
        if (PAIRING == MSFEC_RT_DQ) f << "<PDataArray type=\"Float64\" Name=\"sigma\" NumberOfComponents=\"3\" format=\"ascii\"/>\n<PDataArray type=\"Float64\" Name=\"u\" format=\"ascii\"/>\n";
      }
      f << "</PCellData>\n<PPoints>\n<PDataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\"/>\n</PPoints>\n";
      for (auto &p : pieces) f << "<Piece Source=\"" << p << "\"/>\n";
      f << "</PUnstructuredGrid>\n</VTKFile>\n";
    };
    // pvtu-record for all local coarse outputs
    std::vector<std::string> local_filenames;
    for (int i = 0; i < world_; ++i) local_filenames.push_back(stem + "." + int_to_string(i, 4) + ".vtu");
    pvtu(parameters.dirname_output + "/" + stem + ".pvtu", local_filenames, true);
    // pvtu-record for all local fine outputs: the names of the other ranks follow from their z-order chunks
    // (the reference gathers them with Utilities::MPI::gather, ned_rt_global.cc:585-593)
    std::vector<std::string> filenames_on_cell;
    for (int r = 0; r < world_; ++r) {
      long long a, b;
      owned_range(n_global_cells_, r, world_, a, b);
      long long cnt = 0;
      for (long long id = a; id < b; ++id) {
        if (limit >= 0 && cnt++ >= limit) break;
        filenames_on_cell.push_back(parameters.filename_output + "." + int_to_string(r, 5) + ".cell-" + CellId(id, g).to_string() + ".vtu");
      }
    }
    if (!parameters.standard) {
    const std::string filename_master = parameters.filename_output + "_fine_refine-" + int_to_string(g, 2) + "-" + int_to_string(parameters.n_refine_local, 2) + ".pvtu";
    pvtu(parameters.dirname_output + "/" + filename_master, filenames_on_cell, false);
    }
  }
  t_output_ = now() - t0;
}

// ned_rt_global.cc:704-771
template <int PAIRING>
void Multiscale<PAIRING>::run() {
  if (!parameters.compute_solution) {
    // ned_rt_ref.cc:736-742 / ned_rt_global.cc:707-713
    if (rank_ == 0) std::cout << "Run of " << (parameters.standard ? "standard" : "multiscale") << " problem is explicitly disabled in parameter file." << std::endl;
    return;
  }
  if (rank_ == 0)
    std::cout << "MsFEC_" << name_ << ": running on " << world_ << " rank(s), one GPU each" << std::endl
              << "===========================================" << std::endl
              << "Solving >> " << (parameters.standard ? "STANDARD" : "MULTISCALE") << " << problem in 3D." << std::endl;
  make_grid();
  initialize_and_compute_basis();
  setup_system_matrix();
  assemble_system();
  solve_iterative();
  send_global_weights_to_cell();
  compute_norms();
  output_results();
  // element matrices of the rank's cells (consumed by the tests and by external coarse solvers)
  {
    const int k = batch_->k();
    const std::string out = parameters.dirname_output + "/" + name_ + "_element_matrices.rank" + std::to_string(rank_) + ".bin";
    std::ofstream f(out, std::ios::binary);
    const long long hdr[4] = {hi_ - lo_, k, lo_, PAIRING};
    f.write((const char *)hdr, sizeof(hdr));
    double checksum = 0;
    for (auto &kv : cell_basis_map) {
      const FullMatrix &M = kv.second.get_global_element_matrix();
      const Vector &r = kv.second.get_global_element_rhs();
      f.write((const char *)M.data, sizeof(double) * k * k);
      f.write((const char *)r.data, sizeof(double) * k);
      for (int q = 0; q < k * k; ++q) checksum += M.data[q];
    }
    std::printf("[rank %d] wrote %s ; sum of all matrix entries = %.15e\n", rank_, out.c_str(), checksum);
  }
  if (rank_ == 0) {
    // coarse weights of all cells (every rank holds the same replicated coarse solution)
    const int k = batch_->k();
    std::ofstream f(parameters.dirname_output + "/" + name_ + "_coarse_weights.bin", std::ios::binary);
    const long long hdr[2] = {n_global_cells_, k};
    f.write((const char *)hdr, sizeof(hdr));
    std::vector<double> w(k);
    for (long long id = 0; id < n_global_cells_; ++id) { coarse_->cell_weights(id, w.data()); f.write((const char *)w.data(), sizeof(double) * k); }
  }
  double t[4] = {t_basis_, t_assemble_, t_solve_, t_output_};
  if (comm_ && msfec_comm_allreduce_max(comm_, t, 4)) throw std::runtime_error(std::string("ncclAllReduce: ") + msfec_comm_last_error());
  if (rank_ == 0) {
    std::printf("+---------------------------------------------+------------+\n| Wall clock (max over ranks)                 |  seconds   |\n");
    std::printf("| %-43s | %10.4f |\n| %-43s | %10.4f |\n| %-43s | %10.4f |\n| %-43s | %10.4f |\n+---------------------------------------------+------------+\n",
                (name_ + " basis initialization and computation").c_str(), t[0], "multiscale assembly", t[1], "coarse solve", t[2], "vtu output", t[3]);
    std::fflush(stdout);
  }
}

template class Multiscale<MSFEC_Q>;
template class Multiscale<MSFEC_Q_NED>;
template class Multiscale<MSFEC_NED_RT>;
template class Multiscale<MSFEC_RT_DQ>;

// Mirrors source/main_ned_rt.cxx:15-117: parse "-p <prm>", construct XMultiscale, run(), catch-all.
int driver_main(int argc, char **argv, int pairing, const char *name) {
  try {
    std::string prm_file;
    for (int i = 1; i < argc; ++i) {
      const std::string a = argv[i];
      if (a == "-p" && i + 1 < argc) prm_file = argv[++i];
      else if (a == "-h" || a == "--help") { std::cout << "usage: MsFEC_" << name << " -p parameter_file.prm\n"; return 0; }
      else { std::cerr << "Unknown command line option: " << a << "\nusage: MsFEC_" << name << " -p parameter_file.prm\n"; return 1; }
    }
    if (prm_file.empty()) { std::cerr << "usage: MsFEC_" << name << " -p parameter_file.prm\n"; return 1; }
    // mpirun-style or torchrun-style environment: one rank per GPU
    const int rank = env_int("OMPI_COMM_WORLD_RANK", "PMI_RANK", "RANK", 0);
    const int world = env_int("OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "WORLD_SIZE", 1);
    const int device = env_int("OMPI_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "LOCAL_RANK", 0);
    ParametersMs prm(prm_file, pairing);
    // one NCCL communicator for both runs (a single rank needs none; MSFEC_NCCL=1 creates it anyway)
    msfec_comm *comm = nullptr;
    if (world > 1 || (std::getenv("MSFEC_NCCL") && std::atoi(std::getenv("MSFEC_NCCL")) != 0)) {
      if (msfec_comm_create(rank, world, device, &comm)) throw std::runtime_error(std::string("NCCL communicator: ") + msfec_comm_last_error());
      if (rank == 0) std::cout << "NCCL communicator over " << world << " rank(s) created (one rank per GPU)." << std::endl;
    }
    struct CommGuard { msfec_comm *c; ~CommGuard() { msfec_comm_destroy(c); } } guard{comm};
    auto run_one = [&](const ParametersMs &P, const std::string &label) {
      switch (pairing) {
        case MSFEC_Q: { Multiscale<MSFEC_Q> ms(P, prm_file, rank, world, device, label.c_str(), comm); ms.run(); break; }
        case MSFEC_Q_NED: { Multiscale<MSFEC_Q_NED> ms(P, prm_file, rank, world, device, label.c_str(), comm); ms.run(); break; }
        case MSFEC_NED_RT: { Multiscale<MSFEC_NED_RT> ms(P, prm_file, rank, world, device, label.c_str(), comm); ms.run(); break; }
        default: { Multiscale<MSFEC_RT_DQ> ms(P, prm_file, rank, world, device, label.c_str(), comm); ms.run(); break; }
      }
    };
    // The reference's main_*.cxx (main_ned_rt.cxx:66-82) first runs the fine-grid comparator XStd (`Standard method
    // parameters`), then the multiscale method.  MSFEC_STD_MAX_REFINEMENTS (default 6: 64^3 cells) bounds the comparator's mesh.
    {
      ParametersMs std_prm(prm_file, pairing, /*standard=*/true);
      const int max_ref = std::getenv("MSFEC_STD_MAX_REFINEMENTS") ? std::atoi(std::getenv("MSFEC_STD_MAX_REFINEMENTS")) : 6;
      if (std_prm.compute_solution && std_prm.n_refine_global > max_ref) {
        if (rank == 0) std::cout << "Note: the fine-grid comparator (" << name << "Std) asks for " << std_prm.n_refine_global
                                 << " refinements, MSFEC_STD_MAX_REFINEMENTS = " << max_ref << ": skipped." << std::endl;
      } else {
        run_one(std_prm, std::string(name) + "Std");
      }
    }
    run_one(prm, name);
    return 0;
  } catch (std::exception &exc) {
    std::cerr << "\n----------------------------------------------------\nException on processing:\n" << exc.what()
              << "\nAborting!\n----------------------------------------------------\n";
    return 1;
  } catch (...) {
    std::cerr << "\nUnknown exception!\nAborting!\n";
    return 1;
  }
}

}  // namespace msfec
