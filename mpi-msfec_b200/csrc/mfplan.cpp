// Symbolic phase of the batched multifrontal solver (see mfplan.h).  Host only, runs once per (pairing, n).
#include "mfplan.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <map>
#include <numeric>
#include <stdexcept>

#include "../../include/msfec.h"

namespace msfec {

// s8 is a multiple of 8, so s8 + 4 = 4 or 12 (mod 16): the 8 rows x 4 columns a DMMA fragment load touches fall into
// distinct 8-byte banks
int mf_ldx(int s8) { return s8 + 4; }

// factor record of a front = shared-memory image of k_mf_forward: panel m x ldx, 1/d, d, pivot-tile factors (mf.cuh)
static size_t mf_record(const MfFront &f) { return (size_t)f.m * f.ldx + 10 * (size_t)f.s8; }

size_t mf_smem_forward(const MfFront &f, int n_children, int kr) {
  (void)kr;   // must equal mf_fwd_smem_bytes (mf.cuh): the record + the children's inverse maps
  return mf_record(f) * sizeof(double) + (size_t)n_children * f.m * sizeof(int32_t);
}

size_t mf_smem_forward_staged(const MfFront &f, int n_children) {
  // must equal mf_fwd_st_smem_bytes (mf.cuh): the record, kMfStageBufs stage buffers of 8 child columns, the children's inverse
  // maps, one presence flag per (tile column, child)
  const size_t flags = ((size_t)((f.s8 + f.u8) / 8) * n_children + 7) / 8 * 8;
  return (mf_record(f) + (size_t)kMfStageBufs * 8 * f.ldc_max) * sizeof(double) + (size_t)n_children * f.m * sizeof(int32_t) + flags;
}

size_t mf_smem_backward(const MfFront &f, int kr) {
  // the record + a chunk (<= 64 rows, kMfBwdChunk) of x of the reached unknowns + t / x of the own unknowns ; equals
  // mf_bwd_smem_bytes (mf.cuh)
  return (mf_record(f) + (size_t)(std::min(f.u8, 64) + f.s8) * (kr + 4)) * sizeof(double);
}

namespace {

struct Node {
  std::vector<int> rows;   // stacked interior rows, ascending (sigma-type first)
};

}  // namespace

MfPlan build_mf_plan(const Topology &t, int smem_budget, int min_cells) {
  MfPlan P;
  const int n = t.n, NI0 = t.blk[0].n_int, NI = t.NI;
  P.kr = (t.k_solve + 7) / 8 * 8;
  const int kr = P.kr;
  if (const char *e = std::getenv("MSFEC_MF_MIN_CELLS")) min_cells = std::max(1, std::atoi(e));
  int leaf_x_factor = 1;
  if (const char *e = std::getenv("MSFEC_MF_LEAF_X")) leaf_x_factor = std::max(1, std::atoi(e));
  // MSFEC_MF_SMEM_KB: tighter shared-memory budget per front (more CTAs per SM at the top of the tree, longer chains)
  if (const char *e = std::getenv("MSFEC_MF_SMEM_KB")) smem_budget = std::min(smem_budget, std::max(16, std::atoi(e)) * 1024);

  // ---- nested dissection in doubled integer coordinates 0 .. 2n ----------------------------------------
  struct Dof { int row; int p[3]; };
  std::vector<Dof> all;
  for (int b = 0; b < (t.two_blocks ? 2 : 1); ++b) {
    const int ni = t.blk[b].n_int;
    for (int d = 0; d < ni; ++d) {
      Dof q; q.row = (b == 0 ? 0 : NI0) + d;
      for (int c = 0; c < 3; ++c) q.p[c] = (int)std::lround(2.0 * t.blk[b].pos[3 * d + c]);
      all.push_back(q);
    }
  }
  std::vector<Node> nodes;
  auto emit = [&](const std::vector<int> &idx) {
    if (idx.empty()) return;
    Node nd;
    for (int i : idx) nd.rows.push_back(all[i].row);
    std::sort(nd.rows.begin(), nd.rows.end());
    nodes.push_back(std::move(nd));
  };
  // RT_DQ: a box whose boundary faces are all still uneliminated is a pure-Neumann problem; every box hands ONE cell
  // DoF up to the separator that joins it with its sibling (same rule as the dissected band plan, topology.cpp).
  const bool defer_u = t.pairing == MSFEC_RT_DQ;
  // MSFEC_MF_MERGE=1 (default): relaxed supernodes at the bottom of the tree.  The separator between two leaf boxes is only 8
  // unknowns wide (Ned_RT) but its front would read and rewrite contribution blocks as large as its parent's; such a box hands
  // its separator up to the separator of the enclosing box, which then eliminates [sep A | sep B | own] as one supernode (sigma-
  // type unknowns of all three first, so the no-pivoting argument of DESIGN.md s.3.2 still holds) and has four leaf children.
  bool merge_small = false;
  if (const char *e = std::getenv("MSFEC_MF_MERGE")) merge_small = std::atoi(e) != 0;
  auto cut_axis = [&](const int *lo, const int *hi) {
    int d = -1, best = min_cells;
    for (int a : {2, 1, 0}) if ((hi[a] - lo[a]) / 2 > best) { best = (hi[a] - lo[a]) / 2; d = a; }
    // leaf_x_factor = 2: the last cut (along x) is not made, leaves are boxes of 2 min_cells x min_cells x min_cells fine cells
    // (half as many fronts and one tree level less for ~1.8 x the flops)
    if (d == 0 && (hi[0] - lo[0]) / 2 <= leaf_x_factor * min_cells) d = -1;
    return d;
  };
  std::function<int(const int *, const int *, std::vector<int> &, std::vector<int> *)> rec =
      [&](const int *lo, const int *hi, std::vector<int> &idx, std::vector<int> *hand_up) -> int {
    const int d = cut_axis(lo, hi);
    if (d < 0 || idx.empty()) {
      int deferred = -1;
      if (defer_u) {
        size_t at = 0;
        for (size_t i = 0; i < idx.size(); ++i)
          if (all[idx[i]].row >= NI0 && (deferred < 0 || all[idx[i]].row > all[deferred].row)) { deferred = idx[i]; at = i; }
        if (deferred >= 0) idx.erase(idx.begin() + at);
      }
      emit(idx);
      return deferred;
    }
    const int mid = lo[d] + ((hi[d] - lo[d]) / 4) * 2;
    std::vector<int> L, R, Sp;
    for (int i : idx) (all[i].p[d] < mid ? L : all[i].p[d] > mid ? R : Sp).push_back(i);
    int l2[3] = {lo[0], lo[1], lo[2]}, h2[3] = {hi[0], hi[1], hi[2]};
    h2[d] = mid;
    const bool left_is_leaf = cut_axis(l2, h2) < 0;
    const int dl = rec(l2, h2, L, &Sp);
    h2[d] = hi[d]; l2[d] = mid;
    const bool right_is_leaf = cut_axis(l2, h2) < 0;
    const int dr = rec(l2, h2, R, &Sp);
    if (dl >= 0) Sp.push_back(dl);
    if (merge_small && hand_up && left_is_leaf && right_is_leaf) hand_up->insert(hand_up->end(), Sp.begin(), Sp.end());
    else emit(Sp);
    return dr;
  };
  {
    std::vector<int> idx(all.size());
    std::iota(idx.begin(), idx.end(), 0);
    const int lo[3] = {0, 0, 0}, hi[3] = {2 * n, 2 * n, 2 * n};
    const int left_over = rec(lo, hi, idx, nullptr);
    if (left_over >= 0) {
      nodes.back().rows.push_back(all[left_over].row);
      std::sort(nodes.back().rows.begin(), nodes.back().rows.end());
      P.pinned_row = all[left_over].row;
    }
  }

  // ---- adjacency of the interior rows ---------------------------------------------------------------------
  const RefOperator &S = t.sys;
  std::vector<std::vector<int>> adj(NI);
  for (int r = 0; r < NI; ++r) {
    for (int e = S.cptr[r]; e < S.cptr[r + 1]; ++e) adj[r].push_back(S.ccol[e]);
    for (int e = S.sptr[r]; e < S.sptr[r + 1]; ++e) adj[r].push_back(S.scol[e]);
  }

  // ---- symbolic factorisation on supernodes; wide supernodes are split into chains so that panels fit ------
  struct Sym { std::vector<int> reach; int parent = -1; std::vector<int> kids; };
  std::vector<Sym> sym;
  std::vector<int> base, s8v, node_of_pos;
  auto symbolic = [&]() {
    const int nf = (int)nodes.size();
    base.assign(nf, 0); s8v.assign(nf, 0);
    int off = 0;
    for (int f = 0; f < nf; ++f) { base[f] = off; s8v[f] = ((int)nodes[f].rows.size() + 7) / 8 * 8; off += s8v[f]; }
    P.NP = (off + 31) / 32 * 32;
    P.perm.assign(NI, -1); P.inv_perm.assign(P.NP, -1);
    node_of_pos.assign(P.NP, -1);
    for (int f = 0; f < nf; ++f) {
      for (int i = 0; i < (int)nodes[f].rows.size(); ++i) { P.perm[nodes[f].rows[i]] = base[f] + i; P.inv_perm[base[f] + i] = nodes[f].rows[i]; }
      for (int i = 0; i < s8v[f]; ++i) node_of_pos[base[f] + i] = f;
    }
    for (int r = 0; r < NI; ++r) if (P.perm[r] < 0) throw std::runtime_error("multifrontal plan: interior row not ordered");
    sym.assign(nf, Sym());
    for (int f = 0; f < nf; ++f) {
      std::vector<int> reach;
      const int end = base[f] + s8v[f];
      for (int r : nodes[f].rows) {
        if (r == P.pinned_row) continue;
        for (int c : adj[r]) if (c != P.pinned_row && P.perm[c] >= end) reach.push_back(P.perm[c]);
      }
      for (int c : sym[f].kids) for (int q : sym[c].reach) if (q >= end) reach.push_back(q);
      std::sort(reach.begin(), reach.end());
      reach.erase(std::unique(reach.begin(), reach.end()), reach.end());
      sym[f].reach = std::move(reach);
      if (!sym[f].reach.empty()) {
        sym[f].parent = node_of_pos[sym[f].reach.front()];
        sym[sym[f].parent].kids.push_back(f);
      }
    }
  };
  symbolic();
  {
    // panel budget: m * ldx doubles (+ pivots, child maps) must fit; split too wide supernodes into a chain of pieces
    bool split_any = false;
    std::vector<Node> out;
    for (size_t f = 0; f < nodes.size(); ++f) {
      const int s = (int)nodes[f].rows.size(), s8 = (s + 7) / 8 * 8, u8 = ((int)sym[f].reach.size() + 7) / 8 * 8;
      const int m = s8 + u8 + kr;
      const int nk = std::max<int>(1, (int)sym[f].kids.size());
      int w = s8;
      auto fits = [&](int w8) {
        MfFront ff{}; ff.s8 = w8; ff.u8 = m - kr - w8; ff.m = m; ff.ldx = mf_ldx(w8);
        return (long)mf_smem_forward(ff, nk, kr) <= smem_budget && (long)mf_smem_backward(ff, kr) <= smem_budget;
      };
      while (w > 8 && !fits(w)) w -= 8;
      if (!fits(w)) { P.why = "a front of " + std::to_string(m) + " rows does not fit the shared memory of one SM"; return P; }
      if (w >= s8) { out.push_back(nodes[f]); continue; }
      split_any = true;
      const int pieces = (s + w - 1) / w;
      const int per = ((s + pieces - 1) / pieces + 7) / 8 * 8;
      for (int lo = 0; lo < s; lo += per) {
        Node nd;
        nd.rows.assign(nodes[f].rows.begin() + lo, nodes[f].rows.begin() + std::min(s, lo + per));
        out.push_back(std::move(nd));
      }
    }
    if (split_any) { nodes = std::move(out); symbolic(); }
  }

  // ---- fronts --------------------------------------------------------------------------------------------
  const int nf = (int)nodes.size();
  P.fronts.assign(nf, MfFront());
  std::vector<std::map<int, int>> row_of_pos(nf);   // padded position -> front row (reached part)
  for (int f = 0; f < nf; ++f) {
    MfFront &F = P.fronts[f];
    F.s = (int)nodes[f].rows.size(); F.s8 = s8v[f];
    F.u = (int)sym[f].reach.size(); F.u8 = (F.u + 7) / 8 * 8;
    F.m = F.s8 + F.u8 + kr; F.ldx = mf_ldx(F.s8);
    F.parent = sym[f].parent; F.own_base = base[f];
    F.n_rows_real = F.s + F.u;
    F.level = 0;
    for (int c : sym[f].kids) F.level = std::max(F.level, P.fronts[c].level + 1);
    F.idx_off = (int)P.front_idx.size();
    for (int i = 0; i < F.s8; ++i) P.front_idx.push_back(base[f] + i);
    for (int i = 0; i < F.u8; ++i) {
      P.front_idx.push_back(i < F.u ? sym[f].reach[i] : -1);
      if (i < F.u) row_of_pos[f][sym[f].reach[i]] = F.s8 + i;
    }
    F.row_off = (int)P.own_rows.size();
    for (int i = 0; i < F.s8; ++i) {
      const int r = i < F.s ? nodes[f].rows[i] : -1;
      P.own_rows.push_back(r == P.pinned_row ? -1 : r);
    }
  }
  auto front_row = [&](int f, int pos) -> int {     // row of padded position `pos` inside front f
    if (node_of_pos[pos] == f) return pos - base[f];
    auto it = row_of_pos[f].find(pos);
    if (it == row_of_pos[f].end()) throw std::runtime_error("multifrontal plan: entry outside the symbolic structure");
    return it->second;
  };
  // assembly lists, front by front
  {
    std::vector<std::vector<std::pair<int, int>>> pe(nf), pc(nf);
    std::vector<std::vector<std::pair<int, double>>> ps(nf);
    for (int r = 0; r < NI; ++r) {
      if (r == P.pinned_row) continue;
      const int pr = P.perm[r];
      for (int e = S.cptr[r]; e < S.cptr[r + 1]; ++e) {
        const int c = S.ccol[e], pcn = P.perm[c];
        if (pr < pcn || c == P.pinned_row) continue;
        const int f = node_of_pos[pcn];
        pe[f].push_back({front_row(f, pr) * P.fronts[f].ldx + (pcn - base[f]), S.cref[e]});
      }
      for (int e = S.sptr[r]; e < S.sptr[r + 1]; ++e) {
        const int c = S.scol[e], pcn = P.perm[c];
        if (pr < pcn || c == P.pinned_row) continue;
        const int f = node_of_pos[pcn];
        ps[f].push_back({front_row(f, pr) * P.fronts[f].ldx + (pcn - base[f]), S.sval[e]});
      }
    }
    for (int f = 0; f < nf; ++f) {
      MfFront &F = P.fronts[f];
      for (int i = 0; i < F.s8; ++i) {
        const int r = i < F.s ? nodes[f].rows[i] : -1;
        if (r < 0) pc[f].push_back({i * F.ldx + i, 0});
        else if (r == P.pinned_row) pc[f].push_back({i * F.ldx + i, 1});
      }
      // duplicates (an operator that lists an entry twice) would need "+=": merge them here
      std::sort(pe[f].begin(), pe[f].end());
      std::sort(ps[f].begin(), ps[f].end(), [](const std::pair<int, double> &a, const std::pair<int, double> &b) { return a.first < b.first; });
      for (size_t i = 1; i < pe[f].size(); ++i)
        if (pe[f][i].first == pe[f][i - 1].first) throw std::runtime_error("multifrontal plan: duplicate per-cell entry");
      F.pe_lo = (int)P.pe_dest.size();
      for (auto &x : pe[f]) { P.pe_dest.push_back(x.first); P.pe_ref.push_back(x.second); }
      F.pe_hi = (int)P.pe_dest.size();
      F.ps_lo = (int)P.ps_dest.size();
      for (size_t i = 0; i < ps[f].size(); ++i) {
        if (i > 0 && ps[f][i].first == ps[f][i - 1].first) { P.ps_val.back() += ps[f][i].second; continue; }
        P.ps_dest.push_back(ps[f][i].first); P.ps_val.push_back(ps[f][i].second);
      }
      F.ps_hi = (int)P.ps_dest.size();
      F.pc_lo = (int)P.pc_dest.size();
      for (auto &x : pc[f]) { P.pc_dest.push_back(x.first); P.pc_val.push_back(x.second ? -1.0 : 1.0); }
      F.pc_hi = (int)P.pc_dest.size();
    }
  }
  // per-cell slot entries and shared entries must not collide (the kernel writes both with '=')
  // children and their index maps
  for (int f = 0; f < nf; ++f) {
    MfFront &F = P.fronts[f];
    F.ch_lo = (int)P.children.size();
    for (int c : sym[f].kids) {
      const MfFront &Cf = P.fronts[c];
      MfChild ch{};
      ch.front = c;
      ch.cmap_off = (int)P.cmap.size();
      ch.n_own = 0;
      for (int i = 0; i < Cf.u8 + kr; ++i) {
        int row = -1;
        if (i < Cf.u) {
          row = front_row(f, sym[c].reach[i]);
          if (row < F.s8) {
            if (i != ch.n_own) throw std::runtime_error("multifrontal plan: child columns of the parent are not leading");
            ++ch.n_own;
          }
        } else if (i >= Cf.u8) row = F.s8 + F.u8 + (i - Cf.u8);
        P.cmap.push_back(row);
      }
      F.ldc_max = std::max(F.ldc_max, Cf.u8 + kr);
      ch.pinv_off = (int)P.pinv.size();
      std::vector<int32_t> inv(F.m, -1);
      for (int i = 0; i < Cf.u8 + kr; ++i) { const int row = P.cmap[ch.cmap_off + i]; if (row >= 0) inv[row] = i; }
      P.pinv.insert(P.pinv.end(), inv.begin(), inv.end());
      P.children.push_back(ch);
    }
    F.ch_hi = (int)P.children.size();
    if (F.ch_hi - F.ch_lo > 8) { P.why = "a front has more than 8 children"; return P; }   // kMfMaxChildren (mf.cuh)
  }
  // levels
  for (auto &F : P.fronts) P.n_levels = std::max(P.n_levels, F.level + 1);
  P.level_off.assign(P.n_levels + 1, 0);
  for (auto &F : P.fronts) P.level_off[F.level + 1]++;
  for (int l = 0; l < P.n_levels; ++l) P.level_off[l + 1] += P.level_off[l];
  P.level_fronts.assign(nf, 0);
  {
    std::vector<int> fill(P.level_off.begin(), P.level_off.end() - 1);
    for (int f = 0; f < nf; ++f) P.level_fronts[fill[P.fronts[f].level]++] = f;
  }
  // storage: factor records one after the other.  A contribution block lives from its front's level until its parent's level
  // is done: first-fit arena allocation over the levels (freed blocks are coalesced), so the arena is about two levels deep
  // instead of the sum over all fronts (4 local refinements: 60 MB instead of 1.5 GB per cell)
  {
    int64_t loff = 0;
    for (auto &F : P.fronts) { F.l_off = (int32_t)loff; loff += (int64_t)mf_record(F); }
    P.l_doubles = loff;
    std::map<int64_t, int64_t> free_list;      // offset -> size
    int64_t arena = 0;
    auto release = [&](int64_t off, int64_t sz) {
      if (sz == 0) return;
      auto it = free_list.emplace(off, sz).first;
      auto nx = std::next(it);
      if (nx != free_list.end() && it->first + it->second == nx->first) { it->second += nx->second; free_list.erase(nx); }
      if (it != free_list.begin()) { auto pv = std::prev(it); if (pv->first + pv->second == it->first) { pv->second += it->second; free_list.erase(it); } }
    };
    auto acquire = [&](int64_t sz) -> int64_t {
      if (sz == 0) return 0;
      for (auto it = free_list.begin(); it != free_list.end(); ++it)
        if (it->second >= sz) {
          const int64_t off = it->first, rest = it->second - sz;
          free_list.erase(it);
          if (rest) free_list.emplace(off + sz, rest);
          return off;
        }
      // extend the arena (absorbing a free block at its end)
      int64_t off = arena;
      if (!free_list.empty()) { auto last = std::prev(free_list.end()); if (last->first + last->second == arena) { off = last->first; free_list.erase(last); } }
      arena = off + sz;
      return off;
    };
    for (int l = 0; l < P.n_levels; ++l) {
      for (int i = P.level_off[l]; i < P.level_off[l + 1]; ++i) {
        MfFront &F = P.fronts[P.level_fronts[i]];
        F.c_off = (int32_t)acquire((int64_t)(F.u8 + kr) * F.u8);
      }
      for (int i = P.level_off[l]; i < P.level_off[l + 1]; ++i) {
        const MfFront &F = P.fronts[P.level_fronts[i]];
        for (int c = F.ch_lo; c < F.ch_hi; ++c) { const MfFront &Cf = P.fronts[P.children[c].front]; release(Cf.c_off, (int64_t)(Cf.u8 + kr) * Cf.u8); }
      }
    }
    P.c_doubles = arena;
    if (P.l_doubles >= ((int64_t)1 << 31) || P.c_doubles >= ((int64_t)1 << 31)) { P.why = "multifrontal storage exceeds 2^31 entries per cell"; return P; }
  }
  // shared memory per level, flops, algorithmic bytes
  P.smem_fwd.assign(P.n_levels, 0); P.smem_bwd.assign(P.n_levels, 0);
  P.smem_fwd_st.assign(P.n_levels, 0); P.rt_max.assign(P.n_levels, 0);
  for (auto &F : P.fronts) {
    const int nk = F.ch_hi - F.ch_lo;
    if (nk > 0) P.smem_fwd_st[F.level] = std::max<int32_t>(P.smem_fwd_st[F.level], (int32_t)std::min<size_t>(mf_smem_forward_staged(F, nk), 1u << 30));
    P.rt_max[F.level] = std::max<int32_t>(P.rt_max[F.level], (F.u8 + kr) / 8);
    P.smem_fwd[F.level] = std::max<int32_t>(P.smem_fwd[F.level], (int32_t)mf_smem_forward(F, nk, kr));
    P.smem_bwd[F.level] = std::max<int32_t>(P.smem_bwd[F.level], (int32_t)mf_smem_backward(F, kr));
    const double s = F.s8, u = F.u8, rows_below = F.u8 + kr;
    P.flops += s * s * s / 3.0 + rows_below * s * s + (u * (u + 1) / 2.0 + kr * u) * 2.0 * s;   // LDL^T + X + C -= X L^T
    P.flops += 2.0 * kr * (s * u + s * s / 2.0);                                                  // backward substitution
    // forward: the factor panel is written once, C (lower triangle + rhs rows) is written once and read once by the parent;
    // backward: the panel is read once, the solution of the reached unknowns is gathered, the own solution written
    P.bytes_fwd += 8.0 * (double)F.m * F.s8 + 2.0 * 8.0 * (u * (u + 1) / 2.0 + kr * u);
    P.bytes_bwd += 8.0 * (double)F.m * F.s8 + 8.0 * t.k_solve * (F.u + F.s8);
  }
  if ((long)*std::max_element(P.smem_fwd.begin(), P.smem_fwd.end()) > smem_budget ||
      (long)*std::max_element(P.smem_bwd.begin(), P.smem_bwd.end()) > smem_budget) {
    P.why = "a front does not fit the shared memory of one SM";
    return P;
  }
  if (P.pinned_row >= 0) P.inv_perm[P.perm[P.pinned_row]] = -1;   // solution there stays 0
  P.feasible = true;
  return P;
}

}  // namespace msfec
