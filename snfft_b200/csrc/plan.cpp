// Plan builder (K0): everything the device kernels need, derived from the Young lattice alone --
// no tableau is ever enumerated by filling (the reference scans (n-1)! fillings per shape,
// young_tableau.py:98-125).
//
// The level step of the reference (fft.py:96-112)
//     f_hat(lambda) = sum_{i=1..k} rho_lambda(c_i) . (+)_{mu in lambda-} f_hat_i(mu),  c_i = s_i ... s_{k-1}
// acts on columns independently, and inside the column block of one mu |- k-1 it is a fixed
// linear map L_mu from the k*d_mu input rows (i, U) to the k*d_mu output rows (lambda, T).
// Because rho(s_j) only changes level j of a Gelfand-Tsetlin chain, the matrix entry
// rho_lambda(c_i)[T,U] is a single product of local 2x2 coefficients (yor.py:102-114), and the
// sum over i can be accumulated from the bottom of the chain upwards:
//
//   C_1 = X_1,
//   C_{j+1}[A,T^j..T^1 ; U^{k-1}..U^j] = [U^j = T^j] X_{j+1}[U^{k-1}..U^{j+1},T^j..T^1]
//                                        + sum_V M_j(A,T^j;U^j,V) C_j[T^j..T^1 ; U^{k-1}..U^j,V],
//   f_hat = C_k,
//
// where a row of C_j is indexed by a standard tableau of T^j |- j plus a chain from mu down to
// U^{j-1} with U^{j-1} inside T^j.  For U^j != T^j the sum has one term (a scaling by
// sqrt(1-1/dist^2)); for U^j = T^j it is a small Cauchy-like block [1 | 1/(c(A)-c(V))] of size
// (#corners(T^j)+1).  Every C_j has exactly j*d_mu rows, so the whole level runs in place on
// k*d_mu slots as k-1 stages of d_mu independent blocks; scalings are folded into the blocks that
// consume the row (or into the output scale).  L_mu^{-1} = L_mu^T diag(d_lambda)/(k d_mu), so
// the inverse replays the same tables backwards with transposed blocks.
#include "plan.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <unordered_map>

namespace snfft {

namespace {

void gen_partitions(int n, int start, std::vector<std::vector<int>>& out) {
  // utils.py:52-67
  if (n == 0) {
    out.push_back({});
    return;
  }
  out.push_back({n});
  for (int i = start; i <= n / 2; ++i) {
    std::vector<std::vector<int>> sub;
    gen_partitions(n - i, i, sub);
    for (auto& p : sub) {
      p.push_back(i);
      out.push_back(p);
    }
  }
}

inline int content(const Branch& b) { return b.col - b.row; }

// chains of shape `idx` at level m, in tableau (last-letter) order: ids at levels m, m-1, .., 1
void enum_chains(const HostPlan& plan, int m, int idx, std::vector<uint8_t>& cur,
                 std::vector<std::vector<uint8_t>>& out) {
  cur.push_back((uint8_t)idx);
  if (m == 1) {
    out.push_back(cur);
  } else {
    for (const Branch& b : plan.levels[m].shapes[idx].down) enum_chains(plan, m - 1, b.idx, cur, out);
  }
  cur.pop_back();
}

// index of the tableau given by ids at levels m, m-1, ..., 1
int64_t chain_index(const HostPlan& plan, int m, const uint8_t* ids) {
  int64_t idx = 0;
  for (int lev = m; lev >= 2; --lev) {
    const ShapeInfo& s = plan.levels[lev].shapes[ids[m - lev]];
    const int child = ids[m - lev + 1];
    bool found = false;
    for (const Branch& b : s.down)
      if (b.idx == child) {
        idx += b.off;
        found = true;
        break;
      }
    if (!found) throw std::runtime_error("chain_index: not a chain");
  }
  return idx;
}

const Branch& find_down(const HostPlan& plan, int level, int big, int small) {
  for (const Branch& b : plan.levels[level].shapes[big].down)
    if (b.idx == small) return b;
  throw std::runtime_error("find_down: shapes are not neighbours");
}

struct RegVal {
  int32_t slot;
  double scale;
};

void build_panel(const HostPlan& plan, int k, int mu, LevelProgram& level,
                 std::map<std::vector<double>, uint32_t>& coef_index, PanelProgram& P) {
  const ShapeInfo& smu = plan.levels[k - 1].shapes[mu];
  const int d = (int)smu.dim;
  P.k = k;
  P.mu = mu;
  P.d = d;
  P.in_off = smu.off;
  if ((int64_t)k * d > 65535) throw std::runtime_error("panel has too many rows for 16-bit slots");

  std::vector<std::vector<uint8_t>> chains;
  {
    std::vector<uint8_t> cur;
    enum_chains(plan, k - 1, mu, cur, chains);
  }
  if ((int)chains.size() != d) throw std::runtime_error("chain count != dimension");

  std::unordered_map<std::string, RegVal> reg, next;
  reg.reserve((size_t)k * d * 2);
  // C_1 = X_1 : T = ((1)), U = chain + (empty shape)
  for (int cw = 0; cw < d; ++cw) {
    std::string key;
    key.push_back((char)0);  // T^1 = (1), the only shape of level 1
    key.append((const char*)chains[cw].data(), chains[cw].size());
    key.push_back((char)0);  // U^0 = empty shape
    reg[key] = RegVal{cw, 1.0};
  }

  P.blocks.reserve((size_t)(k - 1) * d);
  std::vector<std::vector<double>> raw_all;  // forward matrices of P.blocks, before folding
  for (int j = 1; j <= k - 1; ++j) {
    next.clear();
    std::vector<Block> stage;
    std::vector<std::vector<double>> raw;  // unfolded forward matrices, same order as `stage`
    std::vector<int> chain_of;             // chain index of each block of `stage`
    stage.reserve(d);
    for (int cw = 0; cw < d; ++cw) {
      const std::vector<uint8_t>& w = chains[cw];  // w[t] = id at level k-1-t
      const int Wj = w[k - 1 - j];
      const std::string lower((const char*)&w[k - 1 - j], (size_t)j);      // levels j..1
      const std::string upper((const char*)w.data(), (size_t)(k - 1 - j));  // levels k-1..j+1
      const ShapeInfo& sj = plan.levels[j].shapes[Wj];
      const int m = 1 + (int)sj.down.size();
      if (m != (int)sj.up.size() || m > kMaxBlock) throw std::runtime_error("bad block size");
      Block blk{};
      blk.m = (uint8_t)m;
      double scale[kMaxBlock];
      blk.slot[0] = (uint16_t)(j * d + cw);  // X_{j+1}[W]
      scale[0] = 1.0;
      for (int b = 1; b < m; ++b) {
        std::string key = lower + upper;
        key.push_back((char)Wj);
        key.push_back((char)sj.down[b - 1].idx);
        auto it = reg.find(key);
        if (it == reg.end()) throw std::runtime_error("missing C_j row");
        blk.slot[b] = (uint16_t)it->second.slot;
        scale[b] = it->second.scale;
        reg.erase(it);
      }
      std::vector<double> cm((size_t)m * m);
      for (int a = 0; a < m; ++a) {
        cm[(size_t)a * m] = 1.0;
        for (int b = 1; b < m; ++b)
          cm[(size_t)a * m + b] =
              scale[b] * (1.0 / (double)(content(sj.up[a]) - content(sj.down[b - 1])));
        std::string key;
        key.push_back((char)sj.up[a].idx);
        key += lower;
        key += upper;
        key.push_back((char)Wj);
        next[key] = RegVal{(int32_t)blk.slot[a], 1.0};
      }
      blk.xoff = 0;
      blk.xoff1 = 0xffffffffu;
      blk.reserved = 0;
      stage.push_back(blk);
      raw.push_back(cm);
      chain_of.push_back(cw);
      level.macs += (double)m * m * d;
    }
    // rows not consumed by a block: U^j != T^j, a pure scaling (folded into a pending scale)
    for (auto& kv : reg) {
      const std::string& key = kv.first;
      const int Tj = (uint8_t)key[0];
      const int Uj = (uint8_t)key[key.size() - 2];
      const int Ujm1 = (uint8_t)key[key.size() - 1];
      if (Tj == Uj) throw std::runtime_error("unconsumed diagonal row");
      const Branch& x = find_down(plan, j, Tj, Ujm1);
      const Branch& y = find_down(plan, j, Uj, Ujm1);
      const double a = 1.0 / (double)(content(x) - content(y));
      const double b = std::sqrt(1 - a * a);
      int Tn = -1;
      for (const Branch& u : plan.levels[j].shapes[Tj].up)
        if (u.row == y.row && u.col == y.col) Tn = u.idx;
      if (Tn < 0) throw std::runtime_error("union shape not found");
      std::string nk;
      nk.push_back((char)Tn);
      nk.append(key, 0, (size_t)j);
      nk.append(key, (size_t)j, key.size() - j - 1);
      next[nk] = RegVal{kv.second.slot, kv.second.scale * b};
    }
    // blocks of equal size together (uniform warps); keep the matrices aligned with them
    std::vector<int> order(stage.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return stage[a].m < stage[b].m; });
    for (int i : order) {
      P.blocks.push_back(stage[i]);
      raw_all.push_back(raw[i]);
      P.block_chain.push_back(chain_of[i]);
    }
    reg.swap(next);
    if (j == 4 && k >= 6) {
      // boundary of the two-phase split: every live row C_5[(T^5, T^4..T^1); ...] belongs to the
      // group of its lower tableau (T^4..T^1); rows X_i[W], i > 5, to the group of W's lower part
      P.low_rho.assign((size_t)k * d, -1);
      P.low_idx.assign((size_t)k * d, -1);
      P.c5_t5.assign((size_t)k * d, -1);
      P.c5_u4.assign((size_t)k * d, -1);
      P.c5_ubase.assign((size_t)k * d, -1);
      P.c5_scale.assign((size_t)k * d, 1.0);
      for (auto& kv : reg) {
        const uint8_t* T = (const uint8_t*)kv.first.data();  // T^5, T^4, .., T^1
        const uint8_t* U = T + 5;                            // U^{k-1}, .., U^5, U^4
        const int slot = kv.second.slot;
        P.low_rho[slot] = T[1];
        P.low_idx[slot] = (int32_t)chain_index(plan, 4, T + 1);
        P.c5_t5[slot] = T[0];
        P.c5_u4[slot] = U[k - 5];
        P.c5_scale[slot] = kv.second.scale;
        // first chain below the upper part (U^{k-1}, .., U^5): offsets of the steps above level 5
        int64_t base = 0;
        for (int lev = k - 1; lev >= 6; --lev) base += find_down(plan, lev, U[k - 1 - lev], U[k - lev]).off;
        P.c5_ubase[slot] = (int32_t)base;
      }
      for (int cw = 0; cw < d; ++cw) {
        const std::vector<uint8_t>& w = chains[cw];
        const uint8_t* low = &w[k - 1 - 4];  // W^4, .., W^1
        for (int i = 5; i < k; ++i) {
          P.low_rho[(size_t)i * d + cw] = low[0];
          P.low_idx[(size_t)i * d + cw] = (int32_t)chain_index(plan, 4, low);
        }
      }
    }
  }

  const int64_t rows = (int64_t)k * d;
  if ((int64_t)reg.size() != rows) throw std::runtime_error("row count mismatch at output");
  P.dest.assign(rows, 0);
  P.rowbase.assign(rows, 0);
  std::vector<double>& fscale = P.fscale;
  std::vector<double>& weight = P.weight;
  fscale.assign(rows, 0.0);
  weight.assign(rows, 0.0);
  P.lam_dim.assign(rows, 0);
  if (k >= 6) {
    P.chain_sigma.resize(d);
    P.chain_base.resize(d);
    for (int cw = 0; cw < d; ++cw) {
      const std::vector<uint8_t>& w = chains[cw];
      P.chain_sigma[cw] = w[k - 1 - 5];
      P.chain_base[cw] = cw - (int32_t)chain_index(plan, 5, &w[k - 1 - 5]);
    }
  }
  std::vector<char> seen(rows, 0);
  for (auto& kv : reg) {
    const std::string& key = kv.first;  // T = (lambda, T^{k-1}, .., T^1), U = (mu)
    if ((int)key.size() != k + 1 || (uint8_t)key[k] != mu) throw std::runtime_error("bad final key");
    const int lam = (uint8_t)key[0];
    const ShapeInfo& sl = plan.levels[k].shapes[lam];
    const int64_t row = chain_index(plan, k, (const uint8_t*)key.data());
    const Branch& bmu = find_down(plan, k, lam, mu);
    const int slot = kv.second.slot;
    if (seen[slot]) throw std::runtime_error("slot used twice");
    seen[slot] = 1;
    P.dest[slot] = (uint32_t)(sl.off + row * sl.dim + bmu.off);
    fscale[slot] = kv.second.scale;
    weight[slot] = (double)sl.dim / ((double)k * (double)d);  // L^-1 = L^T diag(d_lambda/(k d_mu))
    P.lam_dim[slot] = (int32_t)sl.dim;
  }
  for (int i = 0; i < k; ++i)
    for (int u = 0; u < d; ++u)
      P.rowbase[(size_t)i * d + u] = (uint32_t)(i * plan.fact[k - 1] + P.in_off + (int64_t)u * d);

  // Fold the pending output scale of every slot into the block that writes it last, mark those
  // outputs as final, and emit the forward / inverse coefficient matrices.
  std::vector<int> last_writer(rows, -1);
  for (size_t b = 0; b < P.blocks.size(); ++b)
    for (int a = 0; a < P.blocks[b].m; ++a) last_writer[P.blocks[b].slot[a]] = (int)b;
  for (int64_t s = 0; s < rows; ++s)
    if (last_writer[s] < 0) throw std::runtime_error("slot is never written");
  auto intern = [&](const std::vector<double>& cm) -> uint32_t {
    auto it = coef_index.find(cm);
    if (it != coef_index.end()) return it->second;
    uint32_t off = (uint32_t)level.coef.size();
    if (off & 1) {  // keep every matrix 16-byte aligned
      level.coef.push_back(0.0);
      ++off;
    }
    level.coef.insert(level.coef.end(), cm.begin(), cm.end());
    coef_index.emplace(cm, off);
    return off;
  };
  for (size_t b = 0; b < P.blocks.size(); ++b) {
    Block& blk = P.blocks[b];
    const int m = blk.m;
    std::vector<double> F = raw_all[b], I((size_t)m * m);
    blk.final_mask = 0;
    for (int a = 0; a < m; ++a) {
      const int slot = blk.slot[a];
      double wgt = 1.0;
      if (last_writer[slot] == (int)b) {
        blk.final_mask |= (uint8_t)(1u << a);
        for (int c = 0; c < m; ++c) F[(size_t)a * m + c] *= fscale[slot];
        wgt = weight[slot];
      }
      for (int c = 0; c < m; ++c) I[(size_t)c * m + a] = F[(size_t)a * m + c] * wgt;
    }
    blk.coef_f = intern(F);
    blk.coef_i = intern(I);
    P.coef_raw.push_back(intern(raw_all[b]));
    blk.xoff = P.rowbase[blk.slot[0]];
    const int stage_of = (int)(b / d) + 1;
    blk.xoff1 = stage_of == 1 ? P.rowbase[blk.slot[1]] : 0xffffffffu;
  }
}


// Cut the stages of a panel into phases and its rows into groups closed under each phase.
void build_phases(PanelProgram& P) {
  const int k = P.k, d = P.d, rows = k * d;
  const std::vector<std::pair<int, int>> ranges = phase_ranges(k);
  std::vector<int> last_writer(rows, -1);
  for (size_t b = 0; b < P.blocks.size(); ++b)
    for (int a = 0; a < P.blocks[b].m; ++a) last_writer[P.blocks[b].slot[a]] = (int)b;
  for (auto& rg : ranges) {
    PanelPhase ph;
    ph.stage_lo = rg.first;
    ph.stage_hi = rg.second;
    const int b0 = (rg.first - 1) * d, b1 = rg.second * d;  // blocks of these stages
    std::vector<int> parent(rows);
    for (int i = 0; i < rows; ++i) parent[i] = i;
    auto find = [&](int x) {
      while (parent[x] != x) {
        parent[x] = parent[parent[x]];
        x = parent[x];
      }
      return x;
    };
    std::vector<char> touched(rows, 0);
    for (int b = b0; b < b1; ++b) {
      const Block& blk = P.blocks[b];
      const int r0 = find(blk.slot[0]);
      touched[blk.slot[0]] = 1;
      for (int a = 1; a < blk.m; ++a) {
        touched[blk.slot[a]] = 1;
        const int ra = find(blk.slot[a]);
        if (ra != r0) parent[ra] = r0;
      }
    }
    // group id per root, rows in slot order, blocks in stage order
    std::vector<int> gid(rows, -1);
    std::vector<std::vector<int>> grows, gblocks;
    for (int s = 0; s < rows; ++s) {
      if (!touched[s]) continue;
      const int r = find(s);
      if (gid[r] < 0) {
        gid[r] = (int)grows.size();
        grows.emplace_back();
        gblocks.emplace_back();
      }
      grows[gid[r]].push_back(s);
    }
    for (int b = b0; b < b1; ++b) gblocks[gid[find(P.blocks[b].slot[0])]].push_back(b);
    std::vector<int> local(rows, -1);
    for (size_t g = 0; g < grows.size(); ++g) {
      if ((int)grows[g].size() > kPhaseMaxRows) throw std::runtime_error("phase group too large");
      PhaseGroup G;
      G.row_off = (uint32_t)ph.rows.size();
      G.blk_off = (uint32_t)ph.blocks.size();
      G.nrows = (uint32_t)grows[g].size();
      G.nblk = (uint32_t)gblocks[g].size();
      ph.max_rows = std::max<int>(ph.max_rows, G.nrows);
      for (size_t q = 0; q < grows[g].size(); ++q) {
        const int s = grows[g][q];
        local[s] = (int)q;
        const bool fin = last_writer[s] >= b0 && last_writer[s] < b1;
        ph.rows.push_back(PhaseRow{P.rowbase[s], P.dest[s], fin ? 1u : 0u});
        ph.row_slot.push_back((uint32_t)s);
      }
      for (int b : gblocks[g]) {
        const Block& blk = P.blocks[b];
        PhaseBlock pb{};
        pb.m = blk.m;
        pb.stage = (uint8_t)(b / d + 1);
        for (int a = 0; a < blk.m; ++a) pb.loc[a] = (uint16_t)local[blk.slot[a]];
        pb.coef_f = blk.coef_f;
        pb.coef_i = blk.coef_i;
        ph.blocks.push_back(pb);
      }
      ph.groups.push_back(G);
    }
    P.phases.push_back(std::move(ph));
  }
}

}  // namespace

std::vector<std::pair<int, int>> phase_ranges(int k) {
  std::vector<int> ends;
  if (const char* env = std::getenv("SNFFT_WINDOWS")) {
    // "12:8,11/11:10": per level the last stage of every phase after [1..kP1Stages]
    const std::string text(env);
    size_t at = 0;
    while (at < text.size()) {
      size_t stop = text.find('/', at);
      if (stop == std::string::npos) stop = text.size();
      const std::string item = text.substr(at, stop - at);
      const size_t colon = item.find(':');
      if (colon != std::string::npos && std::atoi(item.substr(0, colon).c_str()) == k) {
        size_t q = colon + 1;
        while (q < item.size()) {
          size_t comma = item.find(',', q);
          if (comma == std::string::npos) comma = item.size();
          ends.push_back(std::atoi(item.substr(q, comma - q).c_str()));
          q = comma + 1;
        }
      }
      at = stop + 1;
    }
  }
  if (ends.empty()) {
    // windows of two stages: groups of <= 8 rows, four column chunks per warp in the phase kernel
    // (measured on S_12: 35.5 ms against 37.3 ms with 6..8, 9..k-1)
    for (int e = kP1Stages + 2; e < k - 1; e += 2) ends.push_back(e);
    ends.push_back(k - 1);
  }
  std::vector<std::pair<int, int>> ranges;
  ranges.push_back({1, std::min(kP1Stages, k - 1)});
  int lo = kP1Stages + 1;
  for (int e : ends) {
    if (e < lo || e > k - 1) throw std::runtime_error("SNFFT_WINDOWS: bad stage range");
    ranges.push_back({lo, e});
    lo = e + 1;
  }
  if (lo != k) throw std::runtime_error("SNFFT_WINDOWS: the ranges must end at stage k-1");
  return ranges;
}

void build_host_plan(int n, HostPlan& plan) {
  if (n < 1 || n > kMaxN) throw std::runtime_error("n must be in 1..12");
  plan.n = n;
  plan.fact[0] = 1;
  for (int i = 1; i <= kMaxN + 1; ++i) plan.fact[i] = plan.fact[i - 1] * i;
  plan.levels.assign(n + 2, LevelShapes{});
  std::vector<std::map<std::vector<int>, int>> index(n + 2);
  for (int k = 0; k <= n + 1; ++k) {
    std::vector<std::vector<int>> parts;
    gen_partitions(k, 1, parts);
    if (parts.size() > 255) throw std::runtime_error("too many shapes for 8-bit ids");
    plan.levels[k].shapes.resize(parts.size());
    for (size_t i = 0; i < parts.size(); ++i) {
      plan.levels[k].shapes[i].rows = parts[i];
      index[k][parts[i]] = (int)i;
    }
  }
  for (int k = 0; k <= n; ++k) {
    for (ShapeInfo& s : plan.levels[k].shapes) {
      const std::vector<int>& lam = s.rows;
      const int len = (int)lam.size();
      // down: remove the corner of row r, rows top -> bottom (young_tableau.py:62-81)
      for (int r = 0; r < len; ++r) {
        const int nxt = r + 1 < len ? lam[r + 1] : 0;
        if (lam[r] - 1 >= nxt) {
          std::vector<int> sub;
          for (int q = 0; q < len; ++q) {
            const int v = q == r ? lam[q] - 1 : lam[q];
            if (v > 0) sub.push_back(v);
          }
          s.down.push_back(Branch{index[k - 1].at(sub), r, lam[r] - 1, 0});
        }
      }
      for (int r = 0; r <= len; ++r) {
        const int cur = r < len ? lam[r] : 0;
        if (r == 0 || cur + 1 <= lam[r - 1]) {
          std::vector<int> sup = lam;
          if (r < len)
            sup[r] += 1;
          else
            sup.push_back(1);
          s.up.push_back(Branch{index[k + 1].at(sup), r, cur, 0});
        }
      }
    }
  }
  plan.levels[0].shapes[0].dim = 1;
  for (int k = 1; k <= n; ++k) {
    int64_t off = 0;
    for (ShapeInfo& s : plan.levels[k].shapes) {
      int64_t dsum = 0;
      for (Branch& b : s.down) {
        b.off = dsum;
        dsum += plan.levels[k - 1].shapes[b.idx].dim;
      }
      s.dim = dsum;
      s.off = off;
      off += s.dim * s.dim;
    }
    if (off != plan.fact[k]) throw std::runtime_error("sum d^2 != k!");
  }
  plan.programs.assign(n + 1, LevelProgram{});
  for (int k = 2; k <= n; ++k) {
    LevelProgram& L = plan.programs[k];
    L.k = k;
    const int np = (int)plan.levels[k - 1].shapes.size();
    L.panels.resize(np);
    std::map<std::vector<double>, uint32_t> coef_index;
    for (int mu = 0; mu < np; ++mu) {
      build_panel(plan, k, mu, L, coef_index, L.panels[mu]);
      if (k >= kPhaseMinLevel) build_phases(L.panels[mu]);
    }
    // the phase kernel reads coefficient matrices in 16-byte pieces: keep the last one inside the pool
    L.coef.push_back(0.0);
    L.coef.push_back(0.0);
  }
}

void yor_trans_rows(const HostPlan& plan, int k, int irrep, int j, std::vector<int32_t>& partner,
                    std::vector<double>& diag) {
  if (k < 2 || k > plan.n || j < 1 || j >= k || irrep < 0 ||
      irrep >= (int)plan.levels[k].shapes.size())
    throw std::runtime_error("yor_trans: bad arguments");
  std::vector<std::vector<uint8_t>> chains;
  std::vector<uint8_t> cur;
  enum_chains(plan, k, irrep, cur, chains);
  const size_t d = chains.size();
  partner.assign(d, -1);
  diag.assign(d, 0.0);
  for (size_t t = 0; t < d; ++t) {
    const std::vector<uint8_t>& c = chains[t];  // c[q] = id at level k-q
    const int up_id = c[k - (j + 1)];           // T^{j+1}
    const int mid_id = c[k - j];                // T^j
    const int low_id = j >= 2 ? c[k - (j - 1)] : 0;  // T^{j-1} (empty shape for j = 1)
    const Branch& cell_hi = find_down(plan, j + 1, up_id, mid_id);  // position of j+1
    const Branch& cell_lo = find_down(plan, j, mid_id, low_id);     // position of j
    const double a = 1.0 / (double)(content(cell_hi) - content(cell_lo));
    diag[t] = a;
    // swapping j and j+1: T'^j = T^{j-1} + cell_hi, if that is a shape inside T^{j+1}
    int other = -1;
    for (const Branch& u : plan.levels[j - 1].shapes[low_id].up)
      if (u.row == cell_hi.row && u.col == cell_hi.col) other = u.idx;
    if (other >= 0 && other != mid_id) {
      std::vector<uint8_t> c2 = c;
      c2[k - j] = (uint8_t)other;
      partner[t] = (int32_t)chain_index(plan, k, c2.data());
    }
  }
}

void yor_trans_dense(const HostPlan& plan, int k, int irrep, int j, double* out) {
  std::vector<int32_t> partner;
  std::vector<double> diag;
  yor_trans_rows(plan, k, irrep, j, partner, diag);
  const size_t d = diag.size();
  std::fill(out, out + d * d, 0.0);
  for (size_t t = 0; t < d; ++t) {
    out[t * d + t] = diag[t];
    if (partner[t] >= 0) out[t * d + partner[t]] = std::sqrt(1 - diag[t] * diag[t]);
  }
}

int64_t coset_to_lex(int n, int64_t flat, const int64_t* fact) {
  int p[kMaxN + 1];
  for (int x = 1; x <= n; ++x) p[x] = x;
  for (int k = n; k >= 2; --k) {
    const int i = (int)((flat / fact[k - 1]) % k) + 1;
    // prefix <- prefix o (i i+1 ... k): rotate p[i..k] left by one
    const int first = p[i];
    for (int x = i; x < k; ++x) p[x] = p[x + 1];
    p[k] = first;
  }
  int64_t rank = 0;
  for (int a = 1; a <= n; ++a) {
    int smaller = 0;
    for (int b = a + 1; b <= n; ++b) smaller += p[b] < p[a];
    rank += smaller * fact[n - a];
  }
  return rank;
}

}  // namespace snfft
