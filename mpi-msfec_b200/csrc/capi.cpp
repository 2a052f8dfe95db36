// C ABI glue (include/msfec.h).  No exception crosses this file's extern "C" functions.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>

#include "../../include/msfec.h"
#include "engine.h"
#include "mfplan.h"
#include "prm.h"
#include "topology.h"

using namespace msfec;

struct msfec_ctx {
  ProblemSpec spec;
  Topology topo;
  DirectPlan plan;
  MfPlan mf;
  Engine *engine = nullptr;
  std::string last_error;
};

namespace {

std::mutex g_err_mutex;
std::string g_last_error;

int set_error(msfec_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->last_error = msg;
  std::lock_guard<std::mutex> lk(g_err_mutex);
  g_last_error = msg;
  return code;
}

char *dup_string(const std::string &s) {
  char *p = (char *)std::malloc(s.size() + 1);
  if (p) std::memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

// Euler-angle rotation of the reference (source/equation_data/eqn_coeff_A.cc:8-10,25-56).
void rotation(bool rotate, double R[9]) {
  if (!rotate) {
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    std::memcpy(R, I, sizeof(I));
    return;
  }
  const double a = M_PI / 3, b = M_PI / 6, g = M_PI / 4;
  R[0] = std::cos(a) * std::cos(g) - std::sin(a) * std::cos(b) * std::sin(g);
  R[1] = -std::cos(a) * std::sin(g) - std::sin(a) * std::cos(b) * std::cos(g);
  R[2] = std::sin(a) * std::sin(b);
  R[3] = std::sin(a) * std::cos(g) + std::cos(a) * std::cos(b) * std::sin(g);
  R[4] = -std::sin(a) * std::sin(g) + std::cos(a) * std::cos(b) * std::cos(g);
  R[5] = -std::cos(a) * std::sin(b);
  R[6] = std::sin(b) * std::sin(g);
  R[7] = std::sin(b) * std::cos(g);
  R[8] = std::cos(b);
}

int pairing_k(int pairing) {
  switch (pairing) {
    case MSFEC_Q: return 8;
    case MSFEC_Q_NED: return 20;
    case MSFEC_NED_RT: return 18;
    case MSFEC_RT_DQ: return 7;
    default: return -1;
  }
}

// Resolve strings, compile expressions, fill the kernel parameter block.
void make_spec(const msfec_problem &p, const Topology &topo, ProblemSpec &s) {
  s.p = p;
  s.b_expr = p.b_expression ? p.b_expression : "0";
  s.rhs_expr = p.rhs_expression ? p.rhs_expression : "0";
  s.rhs_consts = p.rhs_constants ? p.rhs_constants : "";
  s.p.b_expression = s.p.rhs_expression = s.p.rhs_constants = nullptr;
  CoefParams &c = s.coef;
  std::memset(&c, 0, sizeof(c));
  rotation(p.a_rotate != 0, c.rot);
  for (int d = 0; d < 3; ++d) { c.a_scale[d] = p.a_scale[d]; c.a_alpha[d] = p.a_alpha[d]; c.a_freq[d] = p.a_freq[d]; }
  c.tensor_inverse = topo.coef_tensor_is_inverse;
  c.scalar_inverse = topo.coef_scalar_is_inverse;
  c.use_random = p.random_field_seed != 0;
  c.seed = p.random_field_seed;
  c.sigma = p.random_field_sigma;
  c.n = topo.n; c.nC = topo.nC;
  c.rhs_ncomp = topo.rhs_ncomp;
  // Diffusion B: constants pi, frequency, scale, alpha (eqn_coeff_B.cc:78-90).  Pairings without a
  // scalar coefficient (Q, RT_DQ) never read the channel; "1" keeps 1/B finite.
  std::map<std::string, double> bc = {{"frequency", (double)p.b_freq}, {"scale", p.b_scale}, {"alpha", p.b_alpha}};
  const bool needs_b = topo.pairing == MSFEC_Q_NED || topo.pairing == MSFEC_NED_RT;
  auto bprog = expr_compile(needs_b ? s.b_expr : std::string("1"), bc);
  c.b_prog_off = 0; c.b_prog_len = (int)bprog.size();
  s.programs = bprog;
  auto comps = expr_split_components(s.rhs_expr);
  if (!p.rhs_expression) comps.assign(topo.rhs_ncomp, "0");   // ParsedFunction default: zero in every component
  if ((int)comps.size() != topo.rhs_ncomp)
    throw std::invalid_argument("right-hand side expression has " + std::to_string(comps.size()) +
                                " components, pairing needs " + std::to_string(topo.rhs_ncomp));
  auto rc = expr_parse_constants(s.rhs_consts);
  for (int i = 0; i < topo.rhs_ncomp; ++i) {
    auto pr = expr_compile(comps[i], rc);
    c.rhs_prog_off[i] = (int)s.programs.size();
    c.rhs_prog_len[i] = (int)pr.size();
    s.programs.insert(s.programs.end(), pr.begin(), pr.end());
  }
}

}  // namespace

extern "C" {

int msfec_abi_version(void) { return MSFEC_ABI_VERSION; }

const char *msfec_last_error(const msfec_ctx *ctx) {
  if (ctx) return ctx->last_error.c_str();
  std::lock_guard<std::mutex> lk(g_err_mutex);
  static thread_local std::string copy;
  copy = g_last_error;
  return copy.c_str();
}

void msfec_problem_defaults(msfec_problem *p, int pairing) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->pairing = pairing;
  p->n_refine_local = 1;        // "local refinements" default (ned_rt_parameters.cc:172-176)
  p->n_refine_global = 2;
  p->a_rotate = 1;              // eqn_coeff_A.cc:128-131
  for (int d = 0; d < 3; ++d) { p->a_scale[d] = 1.0; p->a_alpha[d] = 1.0; }
  p->b_scale = 1.0; p->b_alpha = 1.0;
  p->b_expression = nullptr;    // "0" in the reference (eqn_coeff_B.cc:52-55)
  p->random_field_sigma = std::log(10.0) / 2.0;
}

int msfec_problem_from_prm(const char *prm_path, int pairing, msfec_problem *p) {
  if (!prm_path || !p || pairing_k(pairing) < 0) return set_error(nullptr, MSFEC_EINVAL, "bad argument");
  try {
    const PrmFile f = PrmFile::parse_file(prm_path);
    msfec_problem_defaults(p, pairing);
    const std::string ms = "Multiscale method parameters/", eq = "Equation parameters/";
    p->n_refine_global = (int)f.get_integer(ms + "Mesh/global refinements", 2, 1, 10);
    p->n_refine_local = (int)f.get_integer(ms + "Mesh/local refinements", 1, 1, 10);
    p->use_direct_solver_basis = f.get_bool(ms + "Control flow/use direct solver basis", false);
    p->verbose_basis = f.get_bool(ms + "Control flow/verbose basis", false);
    const char *xyz[3] = {"x", "y", "z"};
    for (int d = 0; d < 3; ++d) {
      p->a_freq[d] = (int)f.get_integer(eq + "Diffusion A/frequency " + xyz[d], 0, 0, 100);
      p->a_scale[d] = f.get_double(eq + "Diffusion A/scale " + xyz[d], 1.0, 0.0001, 10000);
      p->a_alpha[d] = f.get_double(eq + "Diffusion A/alpha " + xyz[d], 1.0, 0.0001, 10000);
    }
    p->a_rotate = f.get_bool(eq + "Diffusion A/rotate", true);
    p->b_freq = (int)f.get_integer(eq + "Diffusion B/frequency", 0, 0, 100);
    p->b_scale = f.get_double(eq + "Diffusion B/scale", 1.0, 0.0001, 10000);
    p->b_alpha = f.get_double(eq + "Diffusion B/alpha", 1.0, 0.0, 10000);
    std::string bexpr = f.get(eq + "Diffusion B/Function expression", "0");
    // "use exact solution = true" swaps the right-hand side for RightHandSideExactLin (ned_rt_basis.cc:396-401,
    // q_ned_basis.cc:394-399) and forces the canonical B (eqn_coeff_B.cc:87-88): the manufactured-solution runs are
    // outside the hot path and not built, so the file is rejected instead of silently computing another problem
    if (f.get_bool(ms + "use exact solution", false))
      return set_error(nullptr, MSFEC_EINVAL, "'use exact solution = true' (manufactured right-hand side) is not supported");
    p->b_expression = dup_string(bexpr);
    p->rhs_expression = dup_string(f.get(eq + "Right-hand side/Function expression", "0"));
    p->rhs_constants = dup_string(f.get(eq + "Right-hand side/Function constants", ""));
    return MSFEC_OK;
  } catch (const std::exception &e) {
    return set_error(nullptr, MSFEC_EPARSE, e.what());
  }
}

void msfec_problem_free(msfec_problem *p) {
  if (!p) return;
  std::free((void *)p->b_expression);
  std::free((void *)p->rhs_expression);
  std::free((void *)p->rhs_constants);
  p->b_expression = p->rhs_expression = p->rhs_constants = nullptr;
}

int msfec_k(int pairing) { return pairing_k(pairing); }

int msfec_n_fine_dofs(int pairing, int L, int *n0, int *n1) {
  if (pairing_k(pairing) < 0 || L < 0 || L > 10) return -1;
  const long n = 1L << L, n1p = n + 1;
  const long V = n1p * n1p * n1p, E = 3 * n * n1p * n1p, F = 3 * n1p * n * n, C = n * n * n;
  long a = 0, b = 0;
  switch (pairing) {
    case MSFEC_Q: a = V; break;
    case MSFEC_Q_NED: a = V; b = E; break;
    case MSFEC_NED_RT: a = E; b = F; break;
    case MSFEC_RT_DQ: a = F; b = C; break;
  }
  if (n0) *n0 = (int)a;
  if (n1) *n1 = (int)b;
  return (int)(a + b);
}

int msfec_create(int device, const msfec_problem *p, msfec_ctx **out) {
  if (!p || !out) return set_error(nullptr, MSFEC_EINVAL, "null argument");
  *out = nullptr;
  if (pairing_k(p->pairing) < 0) return set_error(nullptr, MSFEC_EINVAL, "unknown pairing");
  // 0 local refinements: the coarse cell is its own fine grid, there is nothing to solve and the "multiscale" basis is the
  // standard lowest-order basis -- what the fine-grid comparator of the host driver builds its element matrices with
  if (p->n_refine_local < 0 || p->n_refine_local > 6)
    return set_error(nullptr, MSFEC_EINVAL, "local refinements must be in [0, 6]");
  msfec_ctx *ctx = nullptr;
  try {
    ctx = new msfec_ctx();
    ctx->topo = build_topology(p->pairing, 1 << p->n_refine_local);
    make_spec(*p, ctx->topo, ctx->spec);
    // no interior unknowns (RT_DQ at n = 1: only the pinned cell unknown): no local solve, no factorisation plans
    const bool no_interior = ctx->topo.NI == 0 || (p->pairing == MSFEC_RT_DQ && p->n_refine_local == 0);
    if (no_interior) {
      ctx->plan = DirectPlan();
      ctx->mf = MfPlan();
      ctx->mf.why = "no interior unknowns";
    } else {
      // RT_DQ: a layer block alone is a pure-Neumann sub-problem (singular pivot); keep layer + plane together.
      // Otherwise: layers/planes, or nested dissection when its symbolic flop count is lower (from n = 16 on).
      int ordering = p->pairing == MSFEC_RT_DQ ? 1 : 0;
      bool forced = false;
      if (const char *e = std::getenv("MSFEC_DIRECT_ORDERING")) {
        const std::string v = e;
        if (v == "slab") { ordering = 1; forced = true; }
        else if (v == "split" && p->pairing != MSFEC_RT_DQ) { ordering = 0; forced = true; }
        else if (v == "nd") { ordering = 2; forced = true; }
      }
      std::string why;
      auto try_plan = [&](int ord, DirectPlan &out) {
        try { out = build_direct_plan(ctx->topo, ord); return true; }
        catch (const std::exception &ex) { why = ex.what(); return false; }
      };
      DirectPlan plan;
      bool have = try_plan(ordering, plan);
      // nested dissection only where the factorisation is flop-bound (>= 5 GFLOP per cell) and it saves >= 15 %: its many
      // small blocks mean more (and smaller) launches (C1, Q at n = 16: 0.6 GFLOP, 6.8 ms with layers/planes, 9.9 ms with ND)
      if (!forced && ((have && plan.update_flops >= 5e9) || !have)) {
        DirectPlan nd;
        if (try_plan(2, nd) && (!have || nd.update_flops < 0.85 * plan.update_flops)) { plan = std::move(nd); have = true; }
      }
      // a missing plan is only fatal if the direct path is requested (band > 2^31 entries per cell at 5+ local refinements)
      if (!have && p->use_direct_solver_basis) throw std::runtime_error(why);
      ctx->plan = have ? std::move(plan) : DirectPlan();
    }
    // multifrontal plan (fronts resident in shared memory): feasible up to 3 local refinements; the engine prefers it
    if (!no_interior) {
      try { ctx->mf = build_mf_plan(ctx->topo); }
      catch (const std::exception &ex) { ctx->mf = MfPlan(); ctx->mf.why = ex.what(); }
    }
    if (device >= 0) ctx->engine = engine_create(device, ctx->spec, ctx->topo, ctx->plan, ctx->mf);
    *out = ctx;
    return MSFEC_OK;
  } catch (const std::invalid_argument &e) {
    delete ctx;
    return set_error(nullptr, MSFEC_EINVAL, e.what());
  } catch (const std::bad_alloc &) {
    delete ctx;
    return set_error(nullptr, MSFEC_ENOMEM, "out of memory");
  } catch (const NoDeviceError &e) {
    delete ctx;
    return set_error(nullptr, MSFEC_ENODEVICE, e.what());
  } catch (const ParseError &e) {
    delete ctx;
    return set_error(nullptr, MSFEC_EPARSE, e.what());
  } catch (const std::exception &e) {
    delete ctx;
    return set_error(nullptr, MSFEC_ECUDA, e.what());
  }
}

void msfec_destroy(msfec_ctx *ctx) {
  if (!ctx) return;
  if (ctx->engine) engine_destroy(ctx->engine);
  delete ctx;
}

static int need_engine(msfec_ctx *ctx) {
  if (!ctx) return set_error(nullptr, MSFEC_EINVAL, "null context");
  if (!ctx->engine)
    return set_error(ctx, MSFEC_ENODEVICE, "context was created without a device; there is no CPU fallback");
  return MSFEC_OK;
}

int msfec_build_basis(msfec_ctx *ctx, int n_cells, const double *corners, const int64_t *cell_ids,
                      double *elem_matrix, double *elem_rhs, msfec_stats *stats) {
  if (int rc = need_engine(ctx)) return rc;
  if (!corners || !elem_matrix || !elem_rhs || n_cells <= 0) return set_error(ctx, MSFEC_EINVAL, "bad argument");
  std::string err;
  const int rc = engine_build(ctx->engine, n_cells, corners, cell_ids, elem_matrix, elem_rhs, false, stats, err);
  return rc ? set_error(ctx, rc, err) : rc;
}

int msfec_build_basis_device(msfec_ctx *ctx, int n_cells, const double *d_corners, const int64_t *d_cell_ids,
                             double *d_elem_matrix, double *d_elem_rhs, msfec_stats *stats) {
  if (int rc = need_engine(ctx)) return rc;
  if (!d_corners || !d_elem_matrix || !d_elem_rhs || n_cells <= 0) return set_error(ctx, MSFEC_EINVAL, "bad argument");
  std::string err;
  const int rc = engine_build(ctx->engine, n_cells, d_corners, d_cell_ids, d_elem_matrix, d_elem_rhs, true, stats, err);
  return rc ? set_error(ctx, rc, err) : rc;
}

int msfec_set_weights(msfec_ctx *ctx, int n_cells, const double *weights) {
  if (int rc = need_engine(ctx)) return rc;
  if (!weights) return set_error(ctx, MSFEC_EINVAL, "null weights");
  std::string err;
  const int rc = engine_set_weights(ctx->engine, n_cells, weights, err);
  return rc ? set_error(ctx, rc, err) : rc;
}

int msfec_get_fine_solution(msfec_ctx *ctx, int cell, double *block0, double *block1) {
  if (int rc = need_engine(ctx)) return rc;
  std::string err;
  const int rc = engine_get_fine_solution(ctx->engine, cell, block0, block1, err);
  return rc ? set_error(ctx, rc, err) : rc;
}

int msfec_solution_norms(msfec_ctx *ctx, int n_cells, double *norms) {
  if (int rc = need_engine(ctx)) return rc;
  if (!norms || n_cells <= 0) return set_error(ctx, MSFEC_EINVAL, "bad argument");
  std::string err;
  const int rc = engine_solution_norms(ctx->engine, n_cells, norms, err);
  return rc ? set_error(ctx, rc, err) : rc;
}

int msfec_get_basis(msfec_ctx *ctx, int cell, int basis, double *block0, double *block1) {
  if (int rc = need_engine(ctx)) return rc;
  std::string err;
  const int rc = engine_get_basis(ctx->engine, cell, basis, block0, block1, err);
  return rc ? set_error(ctx, rc, err) : rc;
}

int msfec_fine_dof_layout(const msfec_ctx *ctx, int block, double *pos, int32_t *axis, uint8_t *on_boundary) {
  if (!ctx || block < 0 || block > 1 || (block == 1 && !ctx->topo.two_blocks))
    return set_error(nullptr, MSFEC_EINVAL, "bad block");
  const BlockTopo &b = ctx->topo.blk[block];
  if (pos) std::memcpy(pos, b.pos.data(), b.pos.size() * sizeof(double));
  if (axis) std::memcpy(axis, b.axis.data(), b.axis.size() * sizeof(int32_t));
  if (on_boundary) std::memcpy(on_boundary, b.bnd.data(), b.bnd.size());
  return MSFEC_OK;
}

int msfec_debug_table(const msfec_ctx *ctx, const char *name, void *out, size_t *count, int *dtype) {
  if (!ctx || !name || !count) return set_error(nullptr, MSFEC_EINVAL, "bad argument");
  const Topology &t = ctx->topo;
  const std::string n = name;
  const std::vector<int32_t> *iv = nullptr;
  const std::vector<double> *dv = nullptr;
  std::vector<int32_t> dims;
  auto op_field = [&](const std::string &prefix, const RefOperator &o) {
    if (n == prefix + ".cptr") iv = &o.cptr; else if (n == prefix + ".ccol") iv = &o.ccol;
    else if (n == prefix + ".cref") iv = &o.cref; else if (n == prefix + ".sptr") iv = &o.sptr;
    else if (n == prefix + ".scol") iv = &o.scol; else if (n == prefix + ".sval") dv = &o.sval;
  };
  auto asm_field = [&](const std::string &prefix, const AsmTable &a) {
    if (n == prefix + ".contrib_ptr") iv = &a.contrib_ptr; else if (n == prefix + ".contrib_cell") iv = &a.contrib_cell;
    else if (n == prefix + ".contrib_pair") iv = &a.contrib_pair; else if (n == prefix + ".pair_ptr") iv = &a.pair_ptr;
    else if (n == prefix + ".pair_idx") iv = &a.pair_idx; else if (n == prefix + ".pair_w") dv = &a.pair_w;
  };
  op_field("sys", t.sys); op_field("lift", t.lift); op_field("full", t.full); op_field("kint", t.kint);
  asm_field("asm00", t.asm00); asm_field("asm11", t.asm11); asm_field("asm_rhs", t.asm_rhs);
  const DirectPlan &dp = ctx->plan;
  std::vector<int32_t> col_off32;
  if (n == "direct.perm") iv = &dp.perm; else if (n == "direct.inv_perm") iv = &dp.inv_perm;
  else if (n == "direct.bs") iv = &dp.bs; else if (n == "direct.slab_off") iv = &dp.slab_off;
  else if (n == "direct.ld") iv = &dp.ld; else if (n == "direct.cell_dest") iv = &dp.cell_dest;
  else if (n == "direct.cell_ref") iv = &dp.cell_ref; else if (n == "direct.shared_dest") iv = &dp.shared_dest;
  else if (n == "direct.shared_val") dv = &dp.shared_val; else if (n == "direct.const_dest") iv = &dp.const_dest;
  else if (n == "direct.const_val") dv = &dp.const_val; else if (n == "direct.rhs_dest") iv = &dp.rhs_dest;
  else if (n == "direct.front_rows") iv = &dp.front_rows; else if (n == "direct.chunk_off") iv = &dp.chunk_off;
  else if (n == "direct.chunk_blk") iv = &dp.chunk_blk; else if (n == "direct.chunk_local") iv = &dp.chunk_local;
  else if (n == "direct.front_pos") iv = &dp.front_pos;
  else if (n == "direct.col_off") { for (auto v : dp.col_off) col_off32.push_back((int32_t)v); iv = &col_off32; }
  std::vector<NormOperator> nops;
  std::vector<int32_t> nop_hexp;
  if (n.rfind("norm", 0) == 0) {          // norm<i>.ptr / .col / .val (i = 0..3), norm.h_exponents
    nops = build_norm_operators(t);
    if (n == "norm.h_exponents") { for (auto &o : nops) nop_hexp.push_back(o.h_exponent); iv = &nop_hexp; }
    for (int i = 0; i < 4; ++i) {
      const std::string pre = "norm" + std::to_string(i);
      if (n == pre + ".ptr") iv = &nops[i].ptr; else if (n == pre + ".col") iv = &nops[i].col;
      else if (n == pre + ".val") dv = &nops[i].val;
    }
  }
  const MfPlan &mf = ctx->mf;
  std::vector<int32_t> mf_flat;
  if (n == "mf.fronts") {
    for (auto &F : mf.fronts) { const int32_t *q = reinterpret_cast<const int32_t *>(&F); mf_flat.insert(mf_flat.end(), q, q + kMfFrontFields); }
    iv = &mf_flat;
  } else if (n == "mf.children") {
    for (auto &c : mf.children) { mf_flat.push_back(c.front); mf_flat.push_back(c.n_own); mf_flat.push_back(c.cmap_off); mf_flat.push_back(c.pinv_off); }
    iv = &mf_flat;
  }
  else if (n == "mf.front_idx") iv = &mf.front_idx; else if (n == "mf.own_rows") iv = &mf.own_rows;
  else if (n == "mf.cmap") iv = &mf.cmap; else if (n == "mf.pinv") iv = &mf.pinv;
  else if (n == "mf.pe_dest") iv = &mf.pe_dest; else if (n == "mf.pe_ref") iv = &mf.pe_ref;
  else if (n == "mf.ps_dest") iv = &mf.ps_dest; else if (n == "mf.ps_val") dv = &mf.ps_val;
  else if (n == "mf.pc_dest") iv = &mf.pc_dest; else if (n == "mf.pc_val") dv = &mf.pc_val;
  else if (n == "mf.perm") iv = &mf.perm; else if (n == "mf.inv_perm") iv = &mf.inv_perm;
  else if (n == "mf.level_off") iv = &mf.level_off; else if (n == "mf.level_fronts") iv = &mf.level_fronts;
  else if (n == "mf.smem_fwd") iv = &mf.smem_fwd; else if (n == "mf.smem_bwd") iv = &mf.smem_bwd;
  else if (n == "mf.smem_fwd_st") iv = &mf.smem_fwd_st; else if (n == "mf.rt_max") iv = &mf.rt_max;
  std::vector<double> plan_info;
  if (n == "mf.info") {
    plan_info = {mf.feasible ? 1.0 : 0.0, (double)mf.kr, (double)mf.NP, (double)mf.n_levels, (double)mf.l_doubles,
                 (double)mf.c_doubles, mf.flops, mf.bytes_fwd + mf.bytes_bwd, (double)mf.pinned_row, (double)kMfFrontFields};
    dv = &plan_info;
  }
  if (n == "direct.info") { plan_info = {dp.update_flops, (double)dp.band_doubles, (double)dp.n_slabs, (double)dp.NP}; dv = &plan_info; }
  if (n == "G") dv = &t.G; else if (n == "F1") dv = &t.F1;
  else if (n == "diag_slot0") iv = &t.diag_slot0; else if (n == "diag_slot1") iv = &t.diag_slot1;
  else if (n == "blk0.cell_dofs") iv = &t.blk[0].cell_dofs; else if (n == "blk1.cell_dofs") iv = &t.blk[1].cell_dofs;
  else if (n == "dims") {
    dims = {t.pairing, t.n, t.nC, t.k_solve, t.k_gram, t.k0, t.two_blocks ? 1 : 0, t.blk[0].n_total, t.blk[0].n_int,
            t.two_blocks ? t.blk[1].n_total : 0, t.two_blocks ? t.blk[1].n_int : 0, t.NI, t.NB, t.NF, t.n_slots0,
            t.n_slots1, t.asm_rhs.n_slots, t.rhs_ncomp, t.k_h_exponent, t.f1_H_exponent, t.asm00.h_exponent,
            t.asm11.h_exponent, t.asm_rhs.h_exponent, t.rhs_block, t.coef_tensor_is_inverse, t.coef_scalar_is_inverse};
    iv = &dims;
  }
  if (!iv && !dv) return set_error(nullptr, MSFEC_EINVAL, "unknown table '" + n + "'");
  const size_t cnt = iv ? iv->size() : dv->size();
  if (dtype) *dtype = iv ? 0 : 1;
  if (out) {
    if (*count < cnt) return set_error(nullptr, MSFEC_EINVAL, "buffer too small");
    if (iv) std::memcpy(out, iv->data(), cnt * sizeof(int32_t));
    else std::memcpy(out, dv->data(), cnt * sizeof(double));
  }
  *count = cnt;
  return MSFEC_OK;
}

int msfec_debug_cell_values(msfec_ctx *ctx, int cell, double *values, size_t *count) {
  if (int rc = need_engine(ctx)) return rc;
  if (!count) return set_error(ctx, MSFEC_EINVAL, "null count");
  std::string err;
  const int rc = engine_cell_values(ctx->engine, cell, values, count, err);
  return rc ? set_error(ctx, rc, err) : rc;
}

}  // extern "C"
