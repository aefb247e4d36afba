// C ABI of snfft_b200 (include/snfft_b200.h): plan management, table upload, transform drivers.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/snfft_b200.h"
#include "device_plan.hpp"
#include "kernels.cuh"
#include "plan.hpp"

using namespace snfft;

namespace {

thread_local std::string g_last_error;

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct InvalidArg : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct NoDevice : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define CUDA_OK(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool active = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) throw NoDevice("no CUDA device available");
    if (prev != dev) {
      CUDA_OK(cudaSetDevice(dev));
      active = true;
    }
  }
  ~DeviceGuard() {
    if (active) cudaSetDevice(prev);
  }
};

// device tables of one row-group phase plus the host copies the launcher needs
struct PhaseHost {
  PhaseDev dev{};
  std::vector<int32_t> groups, tiles, pack;  // per panel
};

struct YorTables {
  int32_t* partner = nullptr;
  double* diag = nullptr;
};

}  // namespace

struct snfft_plan {
  HostPlan host;
  int device = -1;
  int num_sms = 148;
  int64_t device_bytes = 0;
  std::vector<void*> allocs;
  LevelDev level_dev[kMaxN + 1];
  Tile* tiles_dev[kMaxN + 1] = {};
  int ntiles[kMaxN + 1] = {};
  size_t level_smem[kMaxN + 1] = {};
  FusedDev fused;
  std::vector<PhaseHost> phase_dev[kMaxN + 1];  // row-group phases of levels >= 9
  P1Dev p1_dev[kMaxN + 1] = {};                // generated in-place stages 1..kP1Stages of levels >= 9
  uint32_t* lex_table = nullptr;
  std::mutex lazy_mu;
  std::map<int, YorTables> yor_tables;
  InvAtItem* inv_at_items = nullptr;    // sampled inverse (snfft_inverse_at): lazily built under lazy_mu
  InvAtIrrep* inv_at_irreps = nullptr;
  int inv_at_nitems = 0;
  size_t inv_at_smem = 0;
  std::mutex dft_mu;             // guards dft_matrix (lazily built; released by snfft_dft_release)
  double* dft_matrix = nullptr;

  template <class T>
  T* upload(const std::vector<T>& v) {
    if (v.empty()) return nullptr;
    void* p = nullptr;
    CUDA_OK(cudaMalloc(&p, v.size() * sizeof(T)));
    allocs.push_back(p);
    device_bytes += (int64_t)(v.size() * sizeof(T));
    CUDA_OK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return (T*)p;
  }
  void* alloc(size_t bytes) {
    void* p = nullptr;
    CUDA_OK(cudaMalloc(&p, bytes));
    allocs.push_back(p);
    device_bytes += (int64_t)bytes;
    return p;
  }
};

namespace {

void plan_tiles(const LevelProgram& prog, const std::vector<int>& width, std::vector<Tile>& tiles,
                size_t& smem_bytes) {
  // width[p] = number of columns of panel p that are present (d_mu for a whole transform)
  // column tiling of every panel: whole panels when they fit in 48 KB, else column strips of up to
  // 96 KB (the widest strip that fits when even one column exceeds that).
  const int64_t kSmall = 6144, kLarge = 12288, kHard = (227 * 1024) / 8;
  int64_t level_elems = kSmall;
  for (size_t p = 0; p < prog.panels.size(); ++p) {
    const PanelProgram& P = prog.panels[p];
    const int64_t rows = (int64_t)P.k * P.d;
    const int64_t cols = width[p];
    if (cols == 0) continue;
    int64_t w;
    if (rows * cols <= kSmall) {
      w = cols;
    } else {
      w = std::max<int64_t>(1, std::min<int64_t>(cols, kLarge / rows));
      const int64_t nt = (cols + w - 1) / w;
      w = (cols + nt - 1) / nt;
    }
    if (rows * w > kHard) throw std::runtime_error("panel column does not fit in shared memory");
    level_elems = std::max(level_elems, rows * w);
    for (int64_t c0 = 0; c0 < cols; c0 += w)
      tiles.push_back(Tile{(int32_t)p, (int32_t)c0, (int32_t)std::min<int64_t>(w, cols - c0), 1});
  }
  for (Tile& t : tiles) {
    const PanelProgram& P = prog.panels[t.panel];
    const int64_t per = (int64_t)P.k * P.d * t.w;
    t.cap = (int32_t)std::max<int64_t>(1, std::min<int64_t>(level_elems / per, 512));
  }
  // heavy tiles first so the tail of the grid is made of cheap ones
  std::stable_sort(tiles.begin(), tiles.end(), [&](const Tile& a, const Tile& b) {
    const PanelProgram& A = prog.panels[a.panel];
    const PanelProgram& B = prog.panels[b.panel];
    return (int64_t)A.k * A.d * a.w > (int64_t)B.k * B.d * b.w;
  });
  smem_bytes = (size_t)level_elems * sizeof(double);
}

void upload_level(snfft_plan* plan, int k) {
  const LevelProgram& prog = plan->host.programs[k];
  std::vector<PanelDev> panels;
  std::vector<Block> blocks;
  std::vector<uint32_t> dest, rowbase;
  int32_t item_off = 0;
  for (const PanelProgram& P : prog.panels) {
    PanelDev pd;
    pd.d = P.d;
    pd.rows = P.k * P.d;
    pd.in_off = (int32_t)P.in_off;
    pd.blk_off = (int32_t)blocks.size();
    pd.slot_off = (int32_t)dest.size();
    pd.item_off = item_off;
    item_off += P.d * P.d;
    panels.push_back(pd);
    blocks.insert(blocks.end(), P.blocks.begin(), P.blocks.end());
    dest.insert(dest.end(), P.dest.begin(), P.dest.end());
    rowbase.insert(rowbase.end(), P.rowbase.begin(), P.rowbase.end());
  }
  LevelDev& L = plan->level_dev[k];
  L.k = k;
  L.npanels = (int32_t)panels.size();
  L.kfact = (int32_t)plan->host.fact[k];
  L.km1fact = (int32_t)plan->host.fact[k - 1];
  L.in_stride = plan->host.fact[k];
  L.out_stride = plan->host.fact[k];
  L.panels = plan->upload(panels);
  L.blocks = plan->upload(blocks);
  L.coef = plan->upload(prog.coef);
  L.dest = plan->upload(dest);
  L.rowbase = plan->upload(rowbase);
  L.item_tab = nullptr;
  if (k == kFusedMax) {
    std::vector<uint32_t> item_tab(plan->host.fact[k - 1]);
    for (size_t p = 0; p < prog.panels.size(); ++p) {
      const PanelProgram& P = prog.panels[p];
      for (int q = 0; q < P.d; ++q)
        for (int c = 0; c < P.d; ++c)
          item_tab[panels[p].item_off + q * P.d + c] = ((uint32_t)p << 24) | ((uint32_t)q << 12) | (uint32_t)c;
    }
    L.item_tab = plan->upload(item_tab);
  }
  std::vector<Tile> tiles;
  std::vector<int> width;
  for (const PanelProgram& P : prog.panels) width.push_back(P.d);
  plan_tiles(prog, width, tiles, plan->level_smem[k]);
  plan->ntiles[k] = (int)tiles.size();
  plan->tiles_dev[k] = plan->upload(tiles);
}

// Layout of one level as seen by the phase kernels: the whole transform (offsets of the plan) or a
// column shard (per-rank compact offsets).
struct LevelLayout {
  std::vector<int> width;                                  // columns of every panel
  std::vector<const uint32_t*> rowbase, dest;              // per panel: slot -> offset
  std::vector<int64_t> panel_off;                          // per panel: offset of the panel in a sub-spectrum
  int64_t in_stride = 0, out_stride = 0, block_stride = 0;  // group strides, distance between X_i blocks
  int64_t lgroups_hint = 1;  // groups of the level in a whole transform of one function (n!/k!)
};

// Device tables of the row-group phases of one level for a given layout (whole transform or column
// shard): per group a program blob, see device_plan.hpp.
template <class Uploader>
void build_phase_devs(const LevelProgram& prog, const LevelLayout& lay, Uploader&& up,
                      std::vector<PhaseHost>& out_devs) {
  if (prog.panels.empty() || prog.panels[0].phases.empty()) return;
  const size_t nphases = prog.panels[0].phases.size();
  out_devs.push_back(PhaseHost{});  // phase 0 (stages 1..kP1Stages) runs as generated programs: placeholder
  for (size_t ph = 1; ph < nphases; ++ph) {
    const int stage_lo = prog.panels[0].phases[ph].stage_lo, stage_hi = prog.panels[0].phases[ph].stage_hi;
    std::vector<PhasePanelDev> panels;
    std::vector<PhaseGroupDev> groups;
    std::vector<U128> blob_f, blob_i;
    int max_rows = 1, max_blob = 1;
    PhaseHost H;
    auto push_words = [](std::vector<U128>& v, const void* src, size_t bytes) {
      const size_t words = (bytes + 15) / 16;
      const size_t at = v.size();
      v.resize(at + words, U128{0, 0, 0, 0});
      std::memcpy(&v[at], src, bytes);
    };
    for (size_t p = 0; p < prog.panels.size(); ++p) {
      const PanelProgram& P = prog.panels[p];
      const PanelPhase& pp = P.phases[ph];
      const int w = lay.width[p];
      PhasePanelDev pd;
      pd.d = w;
      pd.pack = (w > 0 && w <= 16) ? 32 / w : 1;
      pd.tiles = pd.pack > 1 ? 1 : (w + 31) / 32;
      pd.magic = w > 0 ? (65536 + w - 1) / w : 0;
      pd.grp_off = (int32_t)groups.size();
      pd.ngroups = w > 0 ? (int32_t)pp.groups.size() : 0;
      panels.push_back(pd);
      H.groups.push_back(pd.ngroups);
      H.tiles.push_back(pd.tiles);
      H.pack.push_back(pd.pack);
      if (w == 0) continue;
      for (const PhaseGroup& G : pp.groups) {
        const int nr = (int)G.nrows;
        if (nr > 64) throw std::runtime_error("phase group has more than 64 rows (row table of the phase kernel)");
        max_rows = std::max(max_rows, nr);
        // rows that stay on the x side first, then the ones finished in this phase
        std::vector<int> place(nr);
        int nlow = 0;
        for (int r = 0; r < nr; ++r) nlow += pp.rows[G.row_off + r].final ? 0 : 1;
        std::vector<uint32_t> table(2 * (size_t)nr, 0);
        int lo_at = 0, hi_at = nlow;
        for (int r = 0; r < nr; ++r) {
          const uint32_t slot = pp.row_slot[G.row_off + r];
          const int pos = pp.rows[G.row_off + r].final ? hi_at++ : lo_at++;
          place[r] = pos;
          table[pos] = lay.rowbase[p][slot];
          table[nr + pos] = lay.dest[p][slot];
        }
        PhaseGroupDev gd;
        gd.blob_off = (uint32_t)blob_f.size();
        gd.nrows = (uint32_t)nr | ((uint32_t)nlow << 16);
        gd.reserved = 0;
        for (int dir = 0; dir < 2; ++dir) {
          std::vector<U128>& out = dir ? blob_i : blob_f;
          push_words(out, table.data(), table.size() * sizeof(uint32_t));
          bool paired = false;  // the previous block announced this one as its partner
          for (uint32_t bq = 0; bq < G.nblk; ++bq) {
            // forward: stage order; inverse: the same blocks backwards with the inverse matrices
            const PhaseBlock& pb = pp.blocks[G.blk_off + (dir ? G.nblk - 1 - bq : bq)];
            if (pb.stage < stage_lo || pb.stage > stage_hi) throw std::runtime_error("block outside its phase");
            uint16_t off[kMaxBlock] = {0, 0, 0, 0, 0};
            for (int a = 0; a < pb.m; ++a) off[a] = (uint16_t)(place[pb.loc[a]] * 256);
            U128 h;
            h.x = off[0] | ((uint32_t)off[1] << 16);
            h.y = off[2] | ((uint32_t)off[3] << 16);
            h.z = off[4] | ((uint32_t)pb.m << 16);
            // pair flag: the next block belongs to the same stage (blocks of a stage touch disjoint
            // rows) and has the same size <= 3: the kernel runs the two interleaved
            h.w = 0;
            if (!paired && bq + 1 < G.nblk && pb.m <= 3) {
              const PhaseBlock& nx = pp.blocks[G.blk_off + (dir ? G.nblk - 2 - bq : bq + 1)];
              if (nx.stage == pb.stage && nx.m == pb.m) h.w = 1;
            }
            paired = h.w != 0;
            out.push_back(h);
            push_words(out, &prog.coef[dir ? pb.coef_i : pb.coef_f], (size_t)pb.m * pb.m * sizeof(double));
          }
          out.push_back(U128{0, 0, 0, 0});  // terminator (m = 0)
        }
        if (blob_f.size() != blob_i.size()) throw std::runtime_error("blob size mismatch");
        gd.blob_len = (uint32_t)blob_f.size() - gd.blob_off;
        max_blob = std::max(max_blob, (int)gd.blob_len);
        groups.push_back(gd);
      }
    }
    PhaseDev D;
    D.panels = up(panels);
    D.groups = up(groups);
    D.blob_f = up(blob_f);
    D.blob_i = up(blob_i);
    D.in_stride = lay.in_stride;
    D.out_stride = lay.out_stride;
    D.npanels = (int32_t)panels.size();
    D.max_rows = max_rows;
    D.max_blob = max_blob;
    D.nstages = stage_hi - stage_lo + 1;
    H.dev = D;
    out_devs.push_back(std::move(H));
  }
}

template <class Uploader>
void build_p1_dev(const LevelProgram& prog, const LevelLayout& lay, Uploader&& up, P1Dev& D) {
  std::vector<P1Entry> entries;
  std::vector<P0Row> rows;
  for (size_t p = 0; p < prog.panels.size(); ++p) {
    const PanelProgram& P = prog.panels[p];
    const int w = lay.width[p];
    if (w == 0) continue;
    for (int cw = 0; cw < P.d; ++cw)
      if (P.chain_base[cw] == cw)
        entries.push_back(P1Entry{P.chain_sigma[cw], w, lay.panel_off[p] + (int64_t)cw * w, 0});
    // rows whose last writer is a block of stages 1..kP1Stages
    std::vector<int> last(P.k * P.d, -1);
    for (size_t b = 0; b < P.blocks.size(); ++b)
      for (int a = 0; a < P.blocks[b].m; ++a) last[P.blocks[b].slot[a]] = (int)b;
    for (int s = 0; s < P.k * P.d; ++s)
      if (last[s] < kP1Stages * P.d)
        rows.push_back(P0Row{lay.rowbase[p][s], lay.dest[p][s], w, 0, P.fscale[s], P.fscale[s] * P.weight[s], 0});
  }
  std::stable_sort(entries.begin(), entries.end(), [](const P1Entry& a, const P1Entry& b) { return a.sigma < b.sigma; });
  int64_t items = 0;
  std::vector<uint32_t> item_entry, item_row;
  for (size_t i = 0; i < entries.size(); ++i) {
    P1Entry& e = entries[i];
    e.item_off = items;
    const int tiles = (e.d + 31) / 32;
    items += tiles;
    item_entry.insert(item_entry.end(), (size_t)tiles, (uint32_t)i);
  }
  int64_t row_items = 0;
  for (size_t i = 0; i < rows.size(); ++i) {
    P0Row& r = rows[i];
    r.item_off = row_items;
    const int tiles = (r.d + 31) / 32;
    row_items += tiles;
    item_row.insert(item_row.end(), (size_t)tiles, (uint32_t)i);
  }
  D.entries = up(entries);
  D.rows = up(rows);
  D.item_entry = up(item_entry);
  D.item_row = up(item_row);
  D.nitems = items;
  D.nrow_items = row_items;
  D.in_stride = lay.in_stride;
  D.out_stride = lay.out_stride;
  D.rI = lay.block_stride;
  D.nentries = (int32_t)entries.size();
  D.nrows = (int32_t)rows.size();
}

LevelLayout plain_layout(const snfft_plan* plan, int k) {
  const LevelProgram& prog = plan->host.programs[k];
  LevelLayout lay;
  for (const PanelProgram& P : prog.panels) {
    lay.width.push_back(P.d);
    lay.rowbase.push_back(P.rowbase.data());
    lay.dest.push_back(P.dest.data());
    lay.panel_off.push_back(P.in_off);
  }
  lay.in_stride = lay.out_stride = plan->host.fact[k];
  lay.block_stride = plan->host.fact[k - 1];
  lay.lgroups_hint = plan->host.fact[plan->host.n] / plan->host.fact[k];
  return lay;
}

void upload_phases(snfft_plan* plan, int k) {
  const LevelLayout lay = plain_layout(plan, k);
  auto up = [&](const auto& v) { return plan->upload(v); };
  build_phase_devs(plan->host.programs[k], lay, up, plan->phase_dev[k]);
}

void upload_p1(snfft_plan* plan, int k) {
  const LevelLayout lay = plain_layout(plan, k);
  auto up = [&](const auto& v) { return plan->upload(v); };
  build_p1_dev(plan->host.programs[k], lay, up, plan->p1_dev[k]);
}

// cudaFuncSetAttribute is per device: opt the kernels in to large dynamic shared memory once per
// device, also for entry points that take no plan (snfft_dgemm, snfft_c3_dft, snfft_gather)
void ensure_kernels_prepared() {
  static std::mutex mu;
  static bool done[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) throw NoDevice("no CUDA device available");
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= 64 || done[dev]) return;
  CUDA_OK(prepare_kernels());
  done[dev] = true;
}

template <class F>
int guarded(F&& f) {
  try {
    f();
    return SNFFT_OK;
  } catch (const InvalidArg& e) {
    g_last_error = e.what();
    return SNFFT_ERR_INVALID;
  } catch (const NoDevice& e) {
    g_last_error = e.what();
    return SNFFT_ERR_NO_DEVICE;
  } catch (const CudaError& e) {
    g_last_error = e.what();
    return SNFFT_ERR_CUDA;
  } catch (const std::bad_alloc&) {
    g_last_error = "out of host memory";
    return SNFFT_ERR_ALLOC;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return SNFFT_ERR_INVALID;
  }
}

void require_device(const snfft_plan* plan) {
  if (!plan) throw InvalidArg("plan is null");
  if (plan->device < 0) throw NoDevice("plan was created host-only (device = -1); there is no CPU fallback");
}

// the fused kernel stages its input with 16-byte cp.async (kernels.cu: fused7_stage_async)
void require_aligned16(const void* p, const char* name) {
  if (reinterpret_cast<uintptr_t>(p) & 15)
    throw InvalidArg(std::string(name) + " must be 16-byte aligned (a float64 view at an odd element offset is not)");
}

const ShapeInfo& shape_of(const snfft_plan* plan, int k, int irrep) {
  if (!plan) throw InvalidArg("plan is null");
  if (k < 1 || k > plan->host.n) throw InvalidArg("k out of range");
  if (irrep < 0 || irrep >= (int)plan->host.levels[k].shapes.size()) throw InvalidArg("irrep out of range");
  return plan->host.levels[k].shapes[irrep];
}

// Levels >= 9 inside snfft_forward / snfft_inverse: row-group phases.  xside = the buffer holding
// the k sub-spectra per group (forward: the level's input, overwritten with intermediate rows;
// inverse: the level's output), sside = the packed spectra of the level.
void run_phases(const std::vector<PhaseHost>& phases, const P1Dev& p1, int num_sms, double* xside,
                double* sside, int64_t groups, bool inverse, cudaStream_t stream);

void run_level_phases(const snfft_plan* plan, int k, double* xside, double* sside, int64_t groups,
                      bool inverse, cudaStream_t stream) {
  run_phases(plan->phase_dev[k], plan->p1_dev[k], plan->num_sms, xside, sside, groups, inverse, stream);
}

void run_phases(const std::vector<PhaseHost>& phases, const P1Dev& p1, int num_sms, double* xside,
                double* sside, int64_t groups, bool inverse, cudaStream_t stream) {
  const int np = (int)phases.size();
  // phase 0 (stages 1..kP1Stages) runs as generated in-place register programs, the others interpreted
  if (!inverse) CUDA_OK(launch_p1_inplace(p1, xside, sside, groups, false, num_sms, stream));
  for (int i = 1; i < np; ++i)
  {
      const PhaseHost& H = phases[inverse ? np - i : i];
      CUDA_OK(launch_phase(H.dev, H.groups.data(), H.tiles.data(), H.pack.data(), xside, sside, groups, inverse,
                           num_sms, stream));
    }
  if (inverse) CUDA_OK(launch_p1_inplace(p1, xside, sside, groups, true, num_sms, stream));
}

void run_level(const snfft_plan* plan, int k, const double* in, double* out, int64_t groups,
               bool inverse, cudaStream_t stream) {
  CUDA_OK(launch_level(plan->level_dev[k], plan->tiles_dev[k], plan->ntiles[k], plan->level_smem[k],
                       in, out, groups, inverse, g_force_table_driven.load(), plan->num_sms, stream));
}

YorTables& yor_tables_for(snfft_plan* plan, int irrep) {
  std::lock_guard<std::mutex> lock(plan->lazy_mu);
  auto it = plan->yor_tables.find(irrep);
  if (it != plan->yor_tables.end()) return it->second;
  const int n = plan->host.n;
  const int64_t d = plan->host.levels[n].shapes[irrep].dim;
  std::vector<int32_t> partner((size_t)(n - 1) * d);
  std::vector<double> diag((size_t)(n - 1) * d);
  for (int j = 1; j < n; ++j) {
    std::vector<int32_t> p;
    std::vector<double> dg;
    yor_trans_rows(plan->host, n, irrep, j, p, dg);
    std::copy(p.begin(), p.end(), partner.begin() + (size_t)(j - 1) * d);
    std::copy(dg.begin(), dg.end(), diag.begin() + (size_t)(j - 1) * d);
  }
  YorTables t;
  t.partner = plan->upload(partner);
  t.diag = plan->upload(diag);
  return plan->yor_tables.emplace(irrep, t).first->second;
}

}  // namespace


// -----------------------------------------------------------------------------------------------
// column shards (multi-GPU): see include/snfft_b200.h "Sharded transform"
// -----------------------------------------------------------------------------------------------
struct ShardLevel {
  std::vector<int> width;           // per panel mu |- k-1: owned columns w_q(mu)
  std::vector<int64_t> panel_off;   // per panel: offset of its compact block in a level-(k-1) spectrum
  std::vector<uint32_t> dest, rowbase;  // concatenated over panels (like LevelDev)
  std::vector<int> slot_off;        // per panel: first slot
  LevelDev dev;
  Tile* tiles_dev = nullptr;
  int ntiles = 0;
  size_t smem = 0;
  std::vector<PhaseHost> phases;  // k >= 9: row-group phases on the shard's compact layout
  P1Dev p1 = {};
};

struct snfft_shard {
  const snfft_plan* plan = nullptr;
  int rank = 0, world = 1, m = 0;
  // w[q][k][shape]: owned columns of shape |- k (k >= m) on rank q; first[q][shape] for |shape| = m
  std::vector<std::vector<std::vector<int>>> w;
  std::vector<std::vector<int>> first;
  std::vector<std::vector<int64_t>> size;  // size[q][k] = compact size of a level-k spectrum on rank q
  std::vector<ShardLevel> levels;          // levels[k], k = m+1..n, for `rank`
  std::vector<void*> allocs;
  int32_t* pull_idx = nullptr;             // device: gather indices of all ranks, concatenated (lazy)
  std::mutex pull_mutex;
  unsigned long long epoch[2] = {0, 0};    // barrier epochs per channel (every rank counts the same calls)
};

namespace {

int64_t shard_block_begin(int64_t nblocks, int world, int r) {
  return (nblocks / world) * r + std::min<int64_t>(r, nblocks % world);
}

template <class T>
T* shard_upload(snfft_shard* sh, const std::vector<T>& v) {
  if (v.empty()) return nullptr;
  void* p = nullptr;
  CUDA_OK(cudaMalloc(&p, v.size() * sizeof(T)));
  sh->allocs.push_back(p);
  CUDA_OK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return (T*)p;
}

// compact offset of shape `idx` inside a level-k spectrum of rank q
int64_t shard_shape_off(const snfft_shard* sh, int q, int k, int idx) {
  int64_t off = 0;
  const auto& shapes = sh->plan->host.levels[k].shapes;
  for (int i = 0; i < idx; ++i) off += shapes[i].dim * sh->w[q][k][i];
  return off;
}

void build_shard_level(snfft_shard* sh, int k) {
  const HostPlan& host = sh->plan->host;
  const LevelProgram& prog = host.programs[k];
  const int q = sh->rank;
  ShardLevel& SL = sh->levels[k];
  const int64_t s_in = sh->size[q][k - 1];
  for (size_t p = 0; p < prog.panels.size(); ++p) {
    const PanelProgram& P = prog.panels[p];
    const int wq = sh->w[q][k - 1][p];
    SL.width.push_back(wq);
    SL.panel_off.push_back(shard_shape_off(sh, q, k - 1, (int)p));
    SL.slot_off.push_back((int)SL.dest.size());
    // finished row of slot s: (lambda, T) -> off_q(lambda) + row(T) * w_q(lambda) + columns of the
    // branches of lambda that precede mu.  The unsharded dest is off(lambda) + row*d_lambda + off_lambda(mu).
    for (int s = 0; s < P.k * P.d; ++s) {
      const int i = s / P.d, u = s % P.d;
      SL.rowbase.push_back((uint32_t)(i * s_in + SL.panel_off[p] + (int64_t)u * wq));
      // recover (lambda, row) from the unsharded destination
      const int64_t full = P.dest[s];
      int lam = 0;
      const auto& shapes = host.levels[k].shapes;
      while (lam + 1 < (int)shapes.size() && shapes[lam + 1].off <= full) ++lam;
      const int64_t rel = full - shapes[lam].off;
      const int64_t row = rel / shapes[lam].dim;
      int64_t coff = 0;
      bool found = false;
      for (const Branch& b : shapes[lam].down) {
        if (b.idx == (int)p) {
          found = true;
          break;
        }
        coff += sh->w[q][k - 1][b.idx];
      }
      if (!found) throw std::runtime_error("shard: panel is not a branch of its output shape");
      SL.dest.push_back((uint32_t)(shard_shape_off(sh, q, k, lam) + row * sh->w[q][k][lam] + coff));
    }
  }
  if (sh->plan->device < 0) return;
  // device tables: blocks are shared with the plan except for the input offsets used by the inverse
  std::vector<PanelDev> panels;
  std::vector<Block> blocks;
  for (size_t p = 0; p < prog.panels.size(); ++p) {
    const PanelProgram& P = prog.panels[p];
    PanelDev pd;
    pd.d = P.d;
    pd.rows = P.k * P.d;
    pd.in_off = (int32_t)SL.panel_off[p];
    pd.blk_off = (int32_t)blocks.size();
    pd.slot_off = SL.slot_off[p];
    pd.item_off = 0;
    panels.push_back(pd);
    for (size_t b = 0; b < P.blocks.size(); ++b) {
      Block blk = P.blocks[b];
      blk.xoff = SL.rowbase[SL.slot_off[p] + blk.slot[0]];
      if (blk.xoff1 != 0xffffffffu) blk.xoff1 = SL.rowbase[SL.slot_off[p] + blk.slot[1]];
      blocks.push_back(blk);
    }
  }
  LevelDev& L = SL.dev;
  L.k = k;
  L.npanels = (int32_t)panels.size();
  L.kfact = (int32_t)host.fact[k];
  L.km1fact = (int32_t)host.fact[k - 1];
  L.in_stride = (int64_t)k * s_in;
  L.out_stride = sh->size[q][k];
  L.panels = shard_upload(sh, panels);
  L.blocks = shard_upload(sh, blocks);
  L.coef = sh->plan->level_dev[k].coef;  // coefficient pool of the plan
  L.dest = shard_upload(sh, SL.dest);
  L.rowbase = shard_upload(sh, SL.rowbase);
  L.item_tab = nullptr;
  std::vector<Tile> tiles;
  plan_tiles(prog, SL.width, tiles, SL.smem);
  SL.ntiles = (int)tiles.size();
  SL.tiles_dev = shard_upload(sh, tiles);
  if (k >= kPhaseMinLevel) {
    LevelLayout lay;
    for (size_t p = 0; p < prog.panels.size(); ++p) {
      lay.width.push_back(SL.width[p]);
      lay.rowbase.push_back(SL.rowbase.data() + SL.slot_off[p]);
      lay.dest.push_back(SL.dest.data() + SL.slot_off[p]);
      lay.panel_off.push_back(SL.panel_off[p]);
    }
    lay.in_stride = L.in_stride;
    lay.out_stride = L.out_stride;
    lay.block_stride = s_in;
    lay.lgroups_hint = host.fact[host.n] / host.fact[k];
    auto up = [&](const auto& v) { return shard_upload(sh, v); };
    build_phase_devs(prog, lay, up, SL.phases);
    build_p1_dev(prog, lay, up, SL.p1);
  }
}

}  // namespace

extern "C" {

int snfft_abi_version(void) { return 2; }

const char* snfft_last_error(void) { return g_last_error.c_str(); }

int snfft_plan_create(int n, int device, snfft_plan** out) {
  if (!out) {
    g_last_error = "out is null";
    return SNFFT_ERR_INVALID;
  }
  *out = nullptr;
  snfft_plan* plan = nullptr;
  int rc = guarded([&] {
    if (n < 1 || n > SNFFT_MAX_N) throw InvalidArg("n must be in 1..12");
    plan = new snfft_plan();
    build_host_plan(n, plan->host);
    plan->device = device;
    // levels 2..k0 run in the fused kernel (k0 = min(n, 7)); for n <= 3 nothing is fused (k0 = 1)
    plan->fused.k0 = n >= 4 ? std::min(n, kFusedMax) : 1;
    plan->fused.k0fact = (int32_t)plan->host.fact[plan->fused.k0];
    if (device >= 0) {
      int count = 0;
      if (cudaGetDeviceCount(&count) != cudaSuccess || device >= count)
        throw NoDevice("CUDA device " + std::to_string(device) + " is not available");
      DeviceGuard guard(device);
      cudaDeviceProp prop;
      CUDA_OK(cudaGetDeviceProperties(&prop, device));
      if (prop.major < 10) throw NoDevice("snfft_b200 kernels are built for sm_100a only");
      plan->num_sms = prop.multiProcessorCount;
      ensure_kernels_prepared();
      for (int k = 2; k <= n; ++k) upload_level(plan, k);
      for (int k = kPhaseMinLevel; k <= n; ++k) {
        upload_phases(plan, k);
        upload_p1(plan, k);
      }
      for (int k = 2; k <= std::min(n, kFusedMax); ++k) plan->fused.level[k] = plan->level_dev[k];
      if (n <= 10) {
        plan->lex_table = (uint32_t*)plan->alloc((size_t)plan->host.fact[n] * sizeof(uint32_t));
        CUDA_OK(launch_build_lex_table(n, plan->lex_table, 0));
      }
      CUDA_OK(cudaDeviceSynchronize());
    }
  });
  if (rc != SNFFT_OK) {
    if (plan) snfft_plan_destroy(plan);
    return rc;
  }
  *out = plan;
  return SNFFT_OK;
}

int snfft_plan_destroy(snfft_plan* plan) {
  if (!plan) return SNFFT_OK;
  if (plan->device >= 0) {
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(plan->device);
    for (void* p : plan->allocs) cudaFree(p);
    if (plan->dft_matrix) cudaFree(plan->dft_matrix);
    if (prev >= 0) cudaSetDevice(prev);
  }
  delete plan;
  return SNFFT_OK;
}

int snfft_plan_n(const snfft_plan* plan) { return plan ? plan->host.n : -1; }
int snfft_plan_device(const snfft_plan* plan) { return plan ? plan->device : -1; }

int snfft_num_irreps(const snfft_plan* plan, int k) {
  if (!plan || k < 1 || k > plan->host.n) return -1;
  return (int)plan->host.levels[k].shapes.size();
}

int snfft_layout(const snfft_plan* plan, int k, int32_t* parts, int32_t* dims, int64_t* offsets) {
  return guarded([&] {
    if (!plan) throw InvalidArg("plan is null");
    if (k < 1 || k > plan->host.n) throw InvalidArg("k out of range");
    const auto& shapes = plan->host.levels[k].shapes;
    for (size_t i = 0; i < shapes.size(); ++i) {
      if (parts) {
        for (int q = 0; q < k; ++q)
          parts[i * k + q] = q < (int)shapes[i].rows.size() ? shapes[i].rows[q] : 0;
      }
      if (dims) dims[i] = (int32_t)shapes[i].dim;
      if (offsets) offsets[i] = shapes[i].off;
    }
  });
}

int snfft_branch_down(const snfft_plan* plan, int k, int irrep, int32_t* idx, int32_t* off,
                      int32_t* count) {
  return guarded([&] {
    const ShapeInfo& s = shape_of(plan, k, irrep);
    if (!count) throw InvalidArg("count is null");
    if (k == 1) {  // young_tableau.py:67-68: a single box has no branches
      *count = 0;
      return;
    }
    *count = (int32_t)s.down.size();
    for (size_t i = 0; i < s.down.size(); ++i) {
      if (idx) idx[i] = s.down[i].idx;
      if (off) off[i] = (int32_t)s.down[i].off;
    }
  });
}

int snfft_yor_trans(const snfft_plan* plan, int k, int irrep, int j, double* out_host) {
  return guarded([&] {
    shape_of(plan, k, irrep);
    if (!out_host) throw InvalidArg("out_host is null");
    if (j < 1 || j >= k) throw InvalidArg("transposition (j j+1) needs 1 <= j < k");
    yor_trans_dense(plan->host, k, irrep, j, out_host);
  });
}

namespace {

// Levels 2..top of `batch` functions on S_top (top <= n; the same tables serve every top because in
// coset-major order level k of the S_n transform IS level k of each of its S_top blocks).
// f is preserved, the result lands in fhat, work is scratch.  lex: f is in lexicographic order
// (top == n only).
void forward_impl(const snfft_plan* plan, int top, const double* f_dev, double* fhat_dev, double* work_dev,
                  int64_t batch, bool lex, cudaStream_t stream) {
  const int k0 = top >= 4 ? std::min(top, kFusedMax) : 1;
  const int64_t nf = plan->host.fact[top];
  if (top == 1) {
    CUDA_OK(cudaMemcpyAsync(fhat_dev, f_dev, (size_t)batch * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    return;
  }
  // passes: [fused 2..k0 if top >= 4] then one streaming pass per level k0+1..top.  Outputs alternate
  // between work and fhat so that the last one lands in fhat; f is never written.
  const int npasses = (k0 >= 4 ? 1 : 0) + (top - k0);
  auto out_of = [&](int pass) { return ((npasses - 1 - pass) % 2 == 0) ? fhat_dev : work_dev; };
  const double* cur = f_dev;
  if (lex) {
    // reorder into the buffer the first pass does not write (a level pass cannot run in place)
    double* tmp = out_of(0) == fhat_dev ? work_dev : fhat_dev;
    CUDA_OK(launch_reorder(top, plan->lex_table, f_dev, tmp, batch, true, stream));
    cur = tmp;
  }
  int pass = 0;
  if (k0 >= 4) {
    double* nxt = out_of(pass++);
    FusedDev fused = plan->fused;
    fused.k0 = k0;
    fused.k0fact = (int32_t)plan->host.fact[k0];
    CUDA_OK(launch_fused(fused, cur, nxt, batch * (nf / fused.k0fact), false, plan->num_sms, stream));
    cur = nxt;
  }
  for (int k = k0 + 1; k <= top; ++k) {
    double* nxt = out_of(pass++);
    const int64_t groups = batch * (nf / plan->host.fact[k]);
    if (!plan->phase_dev[k].empty() && !g_force_table_driven && cur != f_dev)
      run_level_phases(plan, k, const_cast<double*>(cur), nxt, groups, false, stream);  // cur is scratch
    else
      run_level(plan, k, cur, nxt, groups, false, stream);
    cur = nxt;
  }
}

// The inverse of forward_impl: fhat is preserved, the result lands in f.
void inverse_impl(const snfft_plan* plan, int top, const double* fhat_dev, double* f_dev, double* work_dev,
                  int64_t batch, bool lex, cudaStream_t stream) {
  const int k0 = top >= 4 ? std::min(top, kFusedMax) : 1;
  const int64_t nf = plan->host.fact[top];
  if (top == 1) {
    CUDA_OK(cudaMemcpyAsync(f_dev, fhat_dev, (size_t)batch * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    return;
  }
  // passes: streaming levels top..k0+1, then the fused kernel (levels k0..2).  The last pass writes
  // `target` (work when a reorder to lexicographic order follows, else f); earlier outputs
  // alternate; fhat is never written.
  const int npasses = (k0 >= 4 ? 1 : 0) + (top - k0);
  double* target = lex ? work_dev : f_dev;
  double* other = lex ? f_dev : work_dev;
  auto out_of = [&](int pass) { return ((npasses - 1 - pass) % 2 == 0) ? target : other; };
  const double* cur = fhat_dev;
  int pass = 0;
  for (int k = top; k > k0; --k) {
    double* nxt = out_of(pass++);
    const int64_t groups = batch * (nf / plan->host.fact[k]);
    if (!plan->phase_dev[k].empty() && !g_force_table_driven)
      run_level_phases(plan, k, nxt, const_cast<double*>(cur), groups, true, stream);  // cur is only read
    else
      run_level(plan, k, cur, nxt, groups, true, stream);
    cur = nxt;
  }
  if (k0 >= 4) {
    double* nxt = out_of(pass++);
    FusedDev fused = plan->fused;
    fused.k0 = k0;
    fused.k0fact = (int32_t)plan->host.fact[k0];
    CUDA_OK(launch_fused(fused, cur, nxt, batch * (nf / fused.k0fact), true, plan->num_sms, stream));
    cur = nxt;
  }
  if (lex) CUDA_OK(launch_reorder(top, plan->lex_table, work_dev, f_dev, batch, false, stream));
}

void check_transform_args(const snfft_plan* plan, const void* a, const void* b, const void* c, int64_t batch,
                          int order) {
  require_device(plan);
  if (batch < 0) throw InvalidArg("batch < 0");
  if (batch == 0) return;
  if (!a || !b || !c) throw InvalidArg("null device buffer");
  if (a == b || a == c || b == c) throw InvalidArg("f, fhat and work must be distinct buffers");
  require_aligned16(a, "f_dev");
  require_aligned16(b, "fhat_dev");
  require_aligned16(c, "work_dev");
  if (order != SNFFT_ORDER_COSET && order != SNFFT_ORDER_LEX) throw InvalidArg("bad order flag");
}

}  // namespace

int snfft_forward(const snfft_plan* plan, const double* f_dev, double* fhat_dev, double* work_dev,
                  int64_t batch, int order, void* stream_) {
  return guarded([&] {
    check_transform_args(plan, f_dev, fhat_dev, work_dev, batch, order);
    if (batch == 0) return;
    DeviceGuard guard(plan->device);
    forward_impl(plan, plan->host.n, f_dev, fhat_dev, work_dev, batch, order == SNFFT_ORDER_LEX,
                 (cudaStream_t)stream_);
  });
}

int snfft_inverse(const snfft_plan* plan, const double* fhat_dev, double* f_dev, double* work_dev,
                  int64_t batch, int order, void* stream_) {
  return guarded([&] {
    check_transform_args(plan, f_dev, fhat_dev, work_dev, batch, order);
    if (batch == 0) return;
    DeviceGuard guard(plan->device);
    inverse_impl(plan, plan->host.n, fhat_dev, f_dev, work_dev, batch, order == SNFFT_ORDER_LEX,
                 (cudaStream_t)stream_);
  });
}

int snfft_level(const snfft_plan* plan, int k, const double* in_dev, double* out_dev,
                int64_t groups, int inverse, void* stream) {
  return guarded([&] {
    require_device(plan);
    if (k < 2 || k > plan->host.n) throw InvalidArg("level k must be in 2..n");
    if (!in_dev || !out_dev || in_dev == out_dev) throw InvalidArg("bad buffers");
    if (groups < 0) throw InvalidArg("groups < 0");
    DeviceGuard guard(plan->device);
    run_level(plan, k, in_dev, out_dev, groups, inverse != 0, (cudaStream_t)stream);
  });
}

int snfft_reorder(const snfft_plan* plan, const double* src_dev, double* dst_dev, int64_t batch,
                  int to_coset, void* stream) {
  return guarded([&] {
    require_device(plan);
    if (!src_dev || !dst_dev || src_dev == dst_dev) throw InvalidArg("bad buffers");
    if (batch < 0) throw InvalidArg("batch < 0");
    DeviceGuard guard(plan->device);
    CUDA_OK(launch_reorder(plan->host.n, plan->lex_table, src_dev, dst_dev, batch, to_coset != 0,
                           (cudaStream_t)stream));
  });
}

int snfft_coset_to_lex_host(const snfft_plan* plan, int64_t* out_host) {
  return guarded([&] {
    if (!plan || !out_host) throw InvalidArg("null argument");
    const int n = plan->host.n;
    for (int64_t t = 0; t < plan->host.fact[n]; ++t) out_host[t] = coset_to_lex(n, t, plan->host.fact);
  });
}

namespace {
void yor_impl(const snfft_plan* plan_, int irrep, const int32_t* perms_dev, int64_t count, double* out_dev,
              bool seminormal, void* stream) {
  snfft_plan* plan = const_cast<snfft_plan*>(plan_);  // lazily cached generator tables
  require_device(plan);
  const ShapeInfo& s = shape_of(plan, plan->host.n, irrep);
  if (count < 0) throw InvalidArg("count < 0");
  if (count == 0) return;
  if (!perms_dev || !out_dev) throw InvalidArg("null device buffer");
  DeviceGuard guard(plan->device);
  if (plan->host.n == 1) {  // rho((1)) = [1]: identity "word" on a 1 x 1 matrix, no tables needed
    CUDA_OK(launch_yor(1, 1, nullptr, nullptr, perms_dev, count, out_dev, seminormal, (cudaStream_t)stream));
    return;
  }
  YorTables& t = yor_tables_for(plan, irrep);
  CUDA_OK(launch_yor(plan->host.n, (int)s.dim, t.partner, t.diag, perms_dev, count, out_dev, seminormal,
                     (cudaStream_t)stream));
}
}  // namespace

int snfft_yor(const snfft_plan* plan, int irrep, const int32_t* perms_dev, int64_t count,
              double* out_dev, void* stream) {
  return guarded([&] { yor_impl(plan, irrep, perms_dev, count, out_dev, false, stream); });
}

int snfft_ysemi(const snfft_plan* plan, int irrep, const int32_t* perms_dev, int64_t count,
                double* out_dev, void* stream) {
  return guarded([&] { yor_impl(plan, irrep, perms_dev, count, out_dev, true, stream); });
}

int snfft_inverse_at(const snfft_plan* plan_, const double* fhat_dev, const int32_t* perms_dev, int64_t count,
                     double* out_dev, void* stream) {
  snfft_plan* plan = const_cast<snfft_plan*>(plan_);  // lazily cached generator / item tables
  return guarded([&] {
    require_device(plan);
    if (count < 0) throw InvalidArg("count < 0");
    if (count == 0) return;
    if (!fhat_dev || !perms_dev || !out_dev) throw InvalidArg("null device buffer");
    DeviceGuard guard(plan->device);
    const int n = plan->host.n;
    const auto& shapes = plan->host.levels[n].shapes;
    // generator tables of every irrep first (yor_tables_for takes lazy_mu itself)
    std::vector<YorTables> tabs;
    if (n >= 2)
      for (size_t i = 0; i < shapes.size(); ++i) tabs.push_back(yor_tables_for(plan, (int)i));
    {
      std::lock_guard<std::mutex> lock(plan->lazy_mu);
      if (!plan->inv_at_items) {
        // column tiles: the d x w tile of an irrep has to fit 96 KB of shared memory
        std::vector<InvAtItem> items;
        std::vector<InvAtIrrep> irreps;
        size_t smem = 0;
        for (size_t i = 0; i < shapes.size(); ++i) {
          const int d = (int)shapes[i].dim;
          int w = (int)std::min<int64_t>(d, std::max<int64_t>(1, 12288 / d));
          const int nt = (d + w - 1) / w;
          w = (d + nt - 1) / nt;
          if ((size_t)d * w * sizeof(double) > 200 * 1024) throw std::runtime_error("irrep too large for the sampled inverse");
          smem = std::max(smem, (size_t)d * w * sizeof(double));
          for (int c0 = 0; c0 < d; c0 += w) items.push_back(InvAtItem{(int32_t)i, c0, std::min(w, d - c0), 0});
          InvAtIrrep R{};
          R.d = d;
          R.off = shapes[i].off;
          R.partner = n >= 2 ? tabs[i].partner : nullptr;
          R.diag = n >= 2 ? tabs[i].diag : nullptr;
          R.weight = (double)d / (double)plan->host.fact[n];
          irreps.push_back(R);
        }
        plan->inv_at_irreps = plan->upload(irreps);
        plan->inv_at_nitems = (int)items.size();
        plan->inv_at_smem = smem;
        plan->inv_at_items = plan->upload(items);
      }
    }
    for (int64_t at = 0; at < count; at += 65535) {
      const int64_t cnt = std::min<int64_t>(65535, count - at);
      CUDA_OK(launch_inverse_at(n, plan->inv_at_items, plan->inv_at_nitems, plan->inv_at_irreps, plan->inv_at_smem,
                                fhat_dev, perms_dev + at * n, cnt, out_dev + at, (cudaStream_t)stream));
    }
  });
}

int snfft_dft_direct(snfft_plan* plan, const double* f_dev, double* fhat_dev, int64_t batch,
                     void* stream_) {
  return guarded([&] {
    require_device(plan);
    const int n = plan->host.n;
    if (n > 8) throw InvalidArg("snfft_dft_direct supports n <= 8 (the [n!, n!] matrix is 13 GB at n = 8)");
    if (batch < 0 || batch > 0x7fffffffLL) throw InvalidArg("batch out of range");
    if (batch == 0) return;
    if (!f_dev || !fhat_dev || f_dev == fhat_dev) throw InvalidArg("bad buffers");
    DeviceGuard guard(plan->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t nf = plan->host.fact[n];
    const double* matrix = nullptr;
    {
      // the matrix is built once per plan, on first use (like the yor tables: guarded by lazy_mu)
      std::lock_guard<std::mutex> lock(plan->dft_mu);
      if (!plan->dft_matrix) {
        // R[sigma][packed (lambda,a,b)] = rho_lambda(sigma)[a][b], sigma in lexicographic order
        std::vector<int32_t> perms((size_t)nf * n);
        std::vector<int> p(n);
        for (int i = 0; i < n; ++i) p[i] = i + 1;
        int64_t r = 0;
        do {
          for (int i = 0; i < n; ++i) perms[r * n + i] = p[i];
          ++r;
        } while (std::next_permutation(p.begin(), p.end()));
        struct Scratch {  // freed on every exit path
          void* p = nullptr;
          ~Scratch() { if (p) cudaFree(p); }
        } perms_dev, tmp, R;
        CUDA_OK(cudaMalloc(&perms_dev.p, perms.size() * sizeof(int32_t)));
        CUDA_OK(cudaMemcpy(perms_dev.p, perms.data(), perms.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&R.p, (size_t)nf * nf * sizeof(double)));
        int64_t dmax = 0;
        for (const ShapeInfo& s : plan->host.levels[n].shapes) dmax = std::max(dmax, s.dim);
        CUDA_OK(cudaMalloc(&tmp.p, (size_t)nf * dmax * dmax * sizeof(double)));
        for (size_t irrep = 0; irrep < plan->host.levels[n].shapes.size(); ++irrep) {
          const ShapeInfo& s = plan->host.levels[n].shapes[irrep];
          const int rc = snfft_yor(plan, (int)irrep, (const int32_t*)perms_dev.p, nf, (double*)tmp.p, stream);
          if (rc != SNFFT_OK) throw CudaError(g_last_error);
          CUDA_OK(cudaMemcpy2DAsync((double*)R.p + s.off, (size_t)nf * sizeof(double), tmp.p,
                                    (size_t)(s.dim * s.dim) * sizeof(double),
                                    (size_t)(s.dim * s.dim) * sizeof(double), (size_t)nf,
                                    cudaMemcpyDeviceToDevice, stream));
        }
        CUDA_OK(cudaStreamSynchronize(stream));
        plan->dft_matrix = (double*)R.p;
        plan->device_bytes += (int64_t)((size_t)nf * nf * sizeof(double));
        R.p = nullptr;  // owned by the plan from here on (snfft_dft_release / snfft_plan_destroy)
      }
      matrix = plan->dft_matrix;
    }
    CUDA_OK(launch_dgemm(f_dev, matrix, fhat_dev, (int)batch, (int)nf, (int)nf, stream));
  });
}

int snfft_dft_release(snfft_plan* plan) {
  return guarded([&] {
    if (!plan) throw InvalidArg("plan is null");
    std::lock_guard<std::mutex> lock(plan->dft_mu);
    if (plan->dft_matrix) {
      DeviceGuard guard(plan->device);
      const int64_t nf = plan->host.fact[plan->host.n];
      CUDA_OK(cudaDeviceSynchronize());
      cudaFree(plan->dft_matrix);
      plan->device_bytes -= (int64_t)((size_t)nf * nf * sizeof(double));
      plan->dft_matrix = nullptr;
    }
  });
}


int snfft_shard_create(const snfft_plan* plan, int rank, int world, int m, snfft_shard** out) {
  if (!out) {
    g_last_error = "out is null";
    return SNFFT_ERR_INVALID;
  }
  *out = nullptr;
  snfft_shard* sh = nullptr;
  int rc = guarded([&] {
    if (!plan) throw InvalidArg("plan is null");
    const int n = plan->host.n;
    if (world < 1 || rank < 0 || rank >= world) throw InvalidArg("bad rank / world");
    if (m < 1 || m >= n) throw InvalidArg("split level m must satisfy 1 <= m < n");
    if (plan->host.fact[n] / plan->host.fact[m] < world) throw InvalidArg("fewer S_m blocks than ranks");
    sh = new snfft_shard();
    sh->plan = plan;
    sh->rank = rank;
    sh->world = world;
    sh->m = m;
    const HostPlan& host = plan->host;
    sh->w.assign(world, std::vector<std::vector<int>>(n + 1));
    sh->first.assign(world, {});
    sh->size.assign(world, std::vector<int64_t>(n + 1, 0));
    for (int q = 0; q < world; ++q) {
      const auto& base = host.levels[m].shapes;
      sh->w[q][m].resize(base.size());
      sh->first[q].resize(base.size());
      for (size_t i = 0; i < base.size(); ++i) {
        const int64_t d = base[i].dim;
        const int64_t a = d * q / world, b = d * (q + 1) / world;  // contiguous, nearly equal ranges
        sh->first[q][i] = (int)a;
        sh->w[q][m][i] = (int)(b - a);
      }
      for (int k = m + 1; k <= n; ++k) {
        const auto& shapes = host.levels[k].shapes;
        sh->w[q][k].assign(shapes.size(), 0);
        for (size_t i = 0; i < shapes.size(); ++i)
          for (const Branch& b : shapes[i].down) sh->w[q][k][i] += sh->w[q][k - 1][b.idx];
      }
      for (int k = m; k <= n; ++k) {
        const auto& shapes = host.levels[k].shapes;
        for (size_t i = 0; i < shapes.size(); ++i) sh->size[q][k] += shapes[i].dim * sh->w[q][k][i];
        if (sh->size[q][k] > 0xffffffffLL) throw InvalidArg("shard too large for 32-bit offsets");
      }
    }
    sh->levels.resize(n + 1);
    std::unique_ptr<DeviceGuard> guard;
    if (plan->device >= 0) guard.reset(new DeviceGuard(plan->device));
    for (int k = m + 1; k <= n; ++k) build_shard_level(sh, k);
  });
  if (rc != SNFFT_OK) {
    if (sh) snfft_shard_destroy(sh);
    return rc;
  }
  *out = sh;
  return SNFFT_OK;
}

int snfft_shard_destroy(snfft_shard* sh) {
  if (!sh) return SNFFT_OK;
  if (sh->plan && sh->plan->device >= 0) {
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(sh->plan->device);
    for (void* p : sh->allocs) cudaFree(p);
    if (prev >= 0) cudaSetDevice(prev);
  }
  delete sh;
  return SNFFT_OK;
}

int snfft_shard_blocks(const snfft_shard* sh, int r, int64_t* begin, int64_t* end) {
  return guarded([&] {
    if (!sh || !begin || !end) throw InvalidArg("null argument");
    if (r < 0 || r >= sh->world) throw InvalidArg("rank out of range");
    const HostPlan& host = sh->plan->host;
    const int64_t nb = host.fact[host.n] / host.fact[sh->m];
    *begin = shard_block_begin(nb, sh->world, r);
    *end = shard_block_begin(nb, sh->world, r + 1);
  });
}

int snfft_shard_size(const snfft_shard* sh, int r, int k, int64_t* size) {
  return guarded([&] {
    if (!sh || !size) throw InvalidArg("null argument");
    if (r < 0 || r >= sh->world || k < sh->m || k > sh->plan->host.n) throw InvalidArg("rank or level out of range");
    *size = sh->size[r][k];
  });
}

int snfft_shard_gather_index(const snfft_shard* sh, int r, int64_t* idx) {
  return guarded([&] {
    if (!sh || !idx) throw InvalidArg("null argument");
    if (r < 0 || r >= sh->world) throw InvalidArg("rank out of range");
    const auto& shapes = sh->plan->host.levels[sh->m].shapes;
    int64_t at = 0;
    for (size_t i = 0; i < shapes.size(); ++i) {
      const int64_t d = shapes[i].dim;
      for (int64_t u = 0; u < d; ++u)
        for (int c = 0; c < sh->w[r][sh->m][i]; ++c) idx[at++] = shapes[i].off + u * d + sh->first[r][i] + c;
    }
  });
}

namespace {
void shard_columns_rec(const snfft_shard* sh, int r, int k, int idx, int base, std::vector<int32_t>& out) {
  const ShapeInfo& s = sh->plan->host.levels[k].shapes[idx];
  if (k == sh->m) {
    for (int c = 0; c < sh->w[r][k][idx]; ++c) out.push_back(base + sh->first[r][idx] + c);
    return;
  }
  for (const Branch& b : s.down) shard_columns_rec(sh, r, k - 1, b.idx, base + (int)b.off, out);
}
}  // namespace

int snfft_shard_columns(const snfft_shard* sh, int r, int k, int irrep, int32_t* cols, int32_t* count) {
  return guarded([&] {
    if (!sh || !count) throw InvalidArg("null argument");
    if (r < 0 || r >= sh->world || k < sh->m || k > sh->plan->host.n) throw InvalidArg("rank or level out of range");
    if (irrep < 0 || irrep >= (int)sh->plan->host.levels[k].shapes.size()) throw InvalidArg("irrep out of range");
    std::vector<int32_t> out;
    shard_columns_rec(sh, r, k, irrep, 0, out);
    if ((int)out.size() != sh->w[r][k][irrep]) throw std::runtime_error("shard: column count mismatch");
    *count = (int32_t)out.size();
    if (cols) std::copy(out.begin(), out.end(), cols);
  });
}

int snfft_shard_level(const snfft_shard* sh, int k, const double* in_dev, double* out_dev, int64_t groups,
                      int inverse, void* stream) {
  return guarded([&] {
    if (!sh) throw InvalidArg("shard is null");
    require_device(sh->plan);
    if (k <= sh->m || k > sh->plan->host.n) throw InvalidArg("level k must be in m+1..n");
    if (!in_dev || !out_dev || in_dev == out_dev) throw InvalidArg("bad buffers");
    if (groups < 0) throw InvalidArg("groups < 0");
    DeviceGuard guard(sh->plan->device);
    const ShardLevel& SL = sh->levels[k];
    // always table-driven: the generated level-8 programs assume the full column pitch
    CUDA_OK(launch_level(SL.dev, SL.tiles_dev, SL.ntiles, SL.smem, in_dev, out_dev, groups, inverse != 0,
                         true, sh->plan->num_sms, (cudaStream_t)stream));
  });
}

int snfft_shard_level_scratch(const snfft_shard* sh, int k, double* x_dev, double* s_dev, int64_t groups,
                              int inverse, void* stream) {
  return guarded([&] {
    if (!sh) throw InvalidArg("shard is null");
    require_device(sh->plan);
    if (k <= sh->m || k > sh->plan->host.n) throw InvalidArg("level k must be in m+1..n");
    if (!x_dev || !s_dev || x_dev == s_dev) throw InvalidArg("bad buffers");
    if (groups < 0) throw InvalidArg("groups < 0");
    DeviceGuard guard(sh->plan->device);
    const ShardLevel& SL = sh->levels[k];
    if (!SL.phases.empty() && !g_force_table_driven) {
      run_phases(SL.phases, SL.p1, sh->plan->num_sms, x_dev, s_dev, groups, inverse != 0, (cudaStream_t)stream);
      return;
    }
    CUDA_OK(inverse ? launch_level(SL.dev, SL.tiles_dev, SL.ntiles, SL.smem, s_dev, x_dev, groups, true, true,
                                   sh->plan->num_sms, (cudaStream_t)stream)
                    : launch_level(SL.dev, SL.tiles_dev, SL.ntiles, SL.smem, x_dev, s_dev, groups, false, true,
                                   sh->plan->num_sms, (cudaStream_t)stream));
  });
}

int snfft_shard_export_panel(const snfft_shard* sh, int k, int panel, int32_t* width, int64_t* dest,
                             int64_t* rowbase, int64_t* in_stride, int64_t* out_stride) {
  return guarded([&] {
    if (!sh) throw InvalidArg("shard is null");
    if (k <= sh->m || k > sh->plan->host.n) throw InvalidArg("level k must be in m+1..n");
    const ShardLevel& SL = sh->levels[k];
    if (panel < 0 || panel >= (int)SL.width.size()) throw InvalidArg("panel out of range");
    const PanelProgram& P = sh->plan->host.programs[k].panels[panel];
    if (width) *width = SL.width[panel];
    if (in_stride) *in_stride = (int64_t)k * sh->size[sh->rank][k - 1];
    if (out_stride) *out_stride = sh->size[sh->rank][k];
    for (int s = 0; s < P.k * P.d; ++s) {
      if (dest) dest[s] = SL.dest[SL.slot_off[panel] + s];
      if (rowbase) rowbase[s] = SL.rowbase[SL.slot_off[panel] + s];
    }
  });
}

namespace {

void shard_exchange_impl(snfft_shard* sh, const void* const* peer_bufs, void* local_dev, bool inverse,
                         cudaStream_t stream) {
  if (sh->world > kMaxWorld) throw InvalidArg("peer-memory exchange supports at most 16 ranks");
  const HostPlan& host = sh->plan->host;
  const auto& shapes = host.levels[sh->m].shapes;
  ShardPull X{};
  X.rank = sh->rank;
  X.world = sh->world;
  X.mfact = host.fact[sh->m];
  X.nblocks = host.fact[host.n] / host.fact[sh->m];
  int64_t at = 0;
  for (int r = 0; r < sh->world; ++r) {
    if (!peer_bufs[r]) throw InvalidArg("null peer buffer");
    X.peer[r] = static_cast<const double*>(peer_bufs[r]);
    X.first[r] = shard_block_begin(X.nblocks, sh->world, r);
    X.off[r] = (int32_t)at;
    at += sh->size[r][sh->m];
  }
  X.first[sh->world] = X.nblocks;
  X.off[sh->world] = (int32_t)at;
  if (at != X.mfact) throw std::runtime_error("column shards do not tile the S_m spectrum");
  {
    std::lock_guard<std::mutex> lock(sh->pull_mutex);
    if (!sh->pull_idx) {
      std::vector<int32_t> idx;
      idx.reserve((size_t)X.mfact);
      for (int r = 0; r < sh->world; ++r)
        for (size_t i = 0; i < shapes.size(); ++i) {
          const int64_t d = shapes[i].dim;
          for (int64_t u = 0; u < d; ++u)
            for (int c = 0; c < sh->w[r][sh->m][i]; ++c)
              idx.push_back((int32_t)(shapes[i].off + u * d + sh->first[r][i] + c));
        }
      sh->pull_idx = shard_upload(sh, idx);
    }
  }
  X.idx = sh->pull_idx;
  CUDA_OK(launch_shard_pull(X, static_cast<double*>(local_dev), inverse, stream));
}

// one level m < k <= n on the calling rank's column shard (x side may be overwritten going forward)
void shard_level_impl(const snfft_shard* sh, int k, double* x_dev, double* s_dev, bool inverse, cudaStream_t stream) {
  const HostPlan& host = sh->plan->host;
  const int64_t groups = host.fact[host.n] / host.fact[k];
  const ShardLevel& SL = sh->levels[k];
  if (!SL.phases.empty() && !g_force_table_driven) {
    run_phases(SL.phases, SL.p1, sh->plan->num_sms, x_dev, s_dev, groups, inverse, stream);
    return;
  }
  CUDA_OK(inverse ? launch_level(SL.dev, SL.tiles_dev, SL.ntiles, SL.smem, s_dev, x_dev, groups, true, true,
                                 sh->plan->num_sms, stream)
                  : launch_level(SL.dev, SL.tiles_dev, SL.ntiles, SL.smem, x_dev, s_dev, groups, false, true,
                                 sh->plan->num_sms, stream));
}

void shard_barrier(snfft_shard* sh, const snfft_shard_ws* ws, int channel, cudaStream_t stream) {
  ShardBarrier Bq{};
  for (int r = 0; r < sh->world; ++r) {
    if (!ws->peer_pads[r]) throw InvalidArg("null signal pad");
    Bq.pads[r] = static_cast<unsigned long long*>(ws->peer_pads[r]);
  }
  Bq.epoch = ++sh->epoch[channel];
  Bq.status = ws->status;
  Bq.rank = sh->rank;
  Bq.world = sh->world;
  Bq.channel = channel;
  CUDA_OK(launch_shard_barrier(Bq, stream));
}

void check_ws(const snfft_shard* sh, const snfft_shard_ws* ws) {
  if (!ws) throw InvalidArg("workspace is null");
  if (!ws->local_a || !ws->local_b || !ws->status) throw InvalidArg("workspace has a null buffer");
  if (sh->plan->host.n > sh->m + 1 && (!ws->shard_a || !ws->shard_b)) throw InvalidArg("workspace has a null buffer");
  require_aligned16(ws->local_a, "ws->local_a");
  require_aligned16(ws->local_b, "ws->local_b");
}

}  // namespace

int snfft_shard_exchange(snfft_shard* sh, const void* const* peer_bufs, void* local_dev, int inverse,
                         void* stream) {
  return guarded([&] {
    if (!sh || !peer_bufs || !local_dev) throw InvalidArg("null argument");
    require_device(sh->plan);
    DeviceGuard guard(sh->plan->device);
    shard_exchange_impl(sh, peer_bufs, local_dev, inverse != 0, (cudaStream_t)stream);
  });
}

int snfft_shard_ws_sizes(const snfft_shard* sh, int64_t* sizes) {
  return guarded([&] {
    if (!sh || !sizes) throw InvalidArg("null argument");
    const HostPlan& host = sh->plan->host;
    const int n = host.n, m = sh->m;
    const int64_t nblocks = host.fact[n] / host.fact[m];
    int64_t cols = 0;
    for (int r = 0; r < sh->world; ++r) cols = std::max(cols, nblocks * sh->size[r][m]);
    int64_t b0 = 0, b1 = 0;
    snfft_shard_blocks(sh, sh->rank, &b0, &b1);
    int64_t work = 0;
    for (int k = m + 1; k < n; ++k) work = std::max(work, (host.fact[n] / host.fact[k]) * sh->size[sh->rank][k]);
    sizes[0] = cols;                      // column-shard buffer (same on every rank: symmetric allocation)
    sizes[1] = (b1 - b0) * host.fact[m];  // local_a, local_b
    sizes[2] = work;                      // shard_a, shard_b
    sizes[3] = 2 * kMaxWorld;             // signal pad, in 8-byte words
  });
}

int snfft_forward_sharded(snfft_shard* sh, const double* f_local_dev, double* fhat_compact_dev,
                          const snfft_shard_ws* ws, void* stream_) {
  return guarded([&] {
    if (!sh || !f_local_dev || !fhat_compact_dev) throw InvalidArg("null argument");
    require_device(sh->plan);
    check_ws(sh, ws);
    require_aligned16(f_local_dev, "f_local_dev");
    DeviceGuard guard(sh->plan->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const HostPlan& host = sh->plan->host;
    const int n = host.n, m = sh->m, q = sh->rank;
    int64_t b0 = 0, b1 = 0;
    snfft_shard_blocks(sh, q, &b0, &b1);
    double* spectra = static_cast<double*>(ws->local_a);
    // levels 2..m on this rank's blocks
    forward_impl(sh->plan, m, f_local_dev, spectra, static_cast<double*>(ws->local_b), b1 - b0, false, stream);
    shard_barrier(sh, ws, 0, stream);  // nobody still reads its column buffer
    shard_exchange_impl(sh, ws->peer_cols, spectra, false, stream);  // push my blocks' columns to their owners
    shard_barrier(sh, ws, 1, stream);  // every rank has pushed: my columns are complete
    // levels m+1..n on my column shard
    double* cur = static_cast<double*>(ws->peer_cols[q]);
    for (int k = m + 1; k <= n; ++k) {
      double* nxt = k == n ? fhat_compact_dev
                           : static_cast<double*>((k - m) % 2 ? ws->shard_a : ws->shard_b);
      shard_level_impl(sh, k, cur, nxt, false, stream);
      cur = nxt;
    }
  });
}

int snfft_inverse_sharded(snfft_shard* sh, const double* fhat_compact_dev, double* f_local_dev,
                          const snfft_shard_ws* ws, void* stream_) {
  return guarded([&] {
    if (!sh || !f_local_dev || !fhat_compact_dev) throw InvalidArg("null argument");
    require_device(sh->plan);
    check_ws(sh, ws);
    require_aligned16(f_local_dev, "f_local_dev");
    DeviceGuard guard(sh->plan->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const HostPlan& host = sh->plan->host;
    const int n = host.n, m = sh->m, q = sh->rank;
    int64_t b0 = 0, b1 = 0;
    snfft_shard_blocks(sh, q, &b0, &b1);
    // levels n..m+1 backwards on my column shard; the last one writes my column buffer
    double* cur = const_cast<double*>(fhat_compact_dev);  // only read
    for (int k = n; k > m; --k) {
      double* nxt;
      if (k == m + 1) {
        shard_barrier(sh, ws, 0, stream);  // nobody still pulls from my column buffer
        nxt = static_cast<double*>(ws->peer_cols[q]);
      } else {
        nxt = static_cast<double*>((k - m) % 2 ? ws->shard_a : ws->shard_b);
      }
      shard_level_impl(sh, k, nxt, cur, true, stream);
      cur = nxt;
    }
    shard_barrier(sh, ws, 1, stream);  // every rank's columns are complete
    double* spectra = static_cast<double*>(ws->local_a);
    shard_exchange_impl(sh, ws->peer_cols, spectra, true, stream);  // pull my blocks out of every rank's columns
    inverse_impl(sh->plan, m, spectra, f_local_dev, static_cast<double*>(ws->local_b), b1 - b0, false, stream);
  });
}

int snfft_dgemm(const double* a_dev, const double* b_dev, double* c_dev, int64_t m, int64_t n, int64_t k,
                void* stream) {
  return guarded([&] {
    if (!a_dev || !b_dev || !c_dev) throw InvalidArg("null device buffer");
    if (m < 0 || n < 0 || k < 0 || m > 0x7fffffff || n > 0x7fffffff || k > 0x7fffffff)
      throw InvalidArg("bad GEMM shape");
    ensure_kernels_prepared();
    CUDA_OK(launch_dgemm(a_dev, b_dev, c_dev, (int)m, (int)n, (int)k, (cudaStream_t)stream));
  });
}

}  // extern "C"

// -----------------------------------------------------------------------------------------------
// wreath-product transform as one C object (snfft_wreath_forward)
// -----------------------------------------------------------------------------------------------
struct snfft_wreath {
  int n = 0, device = -1, m = 0;  // m = digits of the orientation DFT (n - 1 for zero-sum orientations, n for all)
  int alpha[3] = {0, 0, 0};
  int64_t N = 0, nH = 0, block = 0, D = 0, nfact = 0, n_ori = 0;
  int32_t* freq = nullptr;
  int64_t *idx = nullptr, *inv_idx = nullptr;
  int32_t* idx_t = nullptr;                                // idx transposed to [h][pair], 32-bit (fused stage 2), or null
  // Outside the fused kernel's range (few cosets or large blocks): stage 2 is the Fourier transform
  // over S_a1 x S_a2 x S_a3 of h -> F_i(t_i h t_j^-1) at the irrep (parts_1, parts_2, parts_3),
  // evaluated factor by factor with the S_a level recursion (forward_impl / inverse_impl of the
  // factor's plan), last factor first.  alpha = (n, 0, 0) up to order is the one-factor, one-coset
  // instance: stage 2 IS the S_n Fourier transform of the character sum.
  struct Factor {
    snfft_plan* plan;
    int a;
    int64_t fact, d, off;  // |S_a|, dimension and packed offset of the irrep parts_k
  };
  std::vector<Factor> ffts;  // the factors with a >= 2, in segment order
  int seg_d[3] = {1, 1, 1};
  bool factor_fft = false;
  double *Y = nullptr, *Z = nullptr;  // with G: the three rotating buffers of the factor transforms
  double *R = nullptr, *RT = nullptr;                      // [nH][b*b] and its transpose
  double *Fre = nullptr, *Fim = nullptr, *G = nullptr, *P = nullptr;  // workspace
  std::vector<void*> allocs;
  std::vector<snfft_plan*> factor_plans;
  std::mutex mu;  // one transform at a time per plan (the workspace is shared)

  template <class T>
  T* alloc(size_t count) {
    void* p = nullptr;
    CUDA_OK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    allocs.push_back(p);
    return (T*)p;
  }
  template <class T>
  T* upload(const std::vector<T>& v) {
    T* p = alloc<T>(v.size());
    if (!v.empty()) CUDA_OK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
  }
};

extern "C" {

int snfft_wreath_create(int n, const int32_t* alpha, const int32_t* parts, int all_orientations, int device,
                        snfft_wreath** out) {
  if (!out) {
    g_last_error = "out is null";
    return SNFFT_ERR_INVALID;
  }
  *out = nullptr;
  snfft_wreath* w = nullptr;
  int rc = guarded([&] {
    if (!alpha || !parts) throw InvalidArg("null argument");
    if (n < 1 || n > 9) throw InvalidArg("n must be in 1..9");
    if (alpha[0] < 0 || alpha[1] < 0 || alpha[2] < 0 || alpha[0] + alpha[1] + alpha[2] != n)
      throw InvalidArg("alpha must be a weak partition of n into three parts");
    if (device < 0) throw NoDevice("the wreath transform needs a CUDA device; there is no CPU fallback");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device >= count)
      throw NoDevice("CUDA device " + std::to_string(device) + " is not available");
    DeviceGuard guard(device);
    ensure_kernels_prepared();
    w = new snfft_wreath();
    w->n = n;
    w->device = device;
    w->m = all_orientations ? n : n - 1;
    if (w->m < 1) throw InvalidArg("the zero-sum orientation group of C_3 wr S_1 is trivial");
    if (w->m > 8) throw InvalidArg("at most 3^8 orientations (the orientation grid lives in shared memory)");
    for (int k = 0; k < 3; ++k) w->alpha[k] = alpha[k];
    w->nfact = 1;
    for (int i = 2; i <= n; ++i) w->nfact *= i;
    w->n_ori = 1;
    for (int i = 0; i < w->m; ++i) w->n_ori *= 3;
    // ---- coset representatives t_i of S_alpha: letters ascending inside every alpha-segment, in
    // lexicographic order (coset_utils.py:72-100) ----
    std::vector<std::vector<int8_t>> reps;
    {
      std::vector<int8_t> cur;
      std::vector<char> used(n + 1, 0);
      std::function<void(int, int, int)> rec = [&](int seg, int filled, int last) {
        if (seg == 3) {
          reps.push_back(cur);
          return;
        }
        if (filled == alpha[seg]) {
          rec(seg + 1, 0, 0);
          return;
        }
        for (int v = last + 1; v <= n; ++v)
          if (!used[v]) {
            used[v] = 1;
            cur.push_back((int8_t)v);
            rec(seg, filled + 1, v);
            cur.pop_back();
            used[v] = 0;
          }
      };
      rec(0, 0, 0);
      std::sort(reps.begin(), reps.end());
    }
    w->N = (int64_t)reps.size();
    std::vector<int8_t> T((size_t)w->N * n), Tinv((size_t)w->N * n);
    std::vector<int32_t> freq((size_t)w->N);
    int seg_of[kMaxN];
    for (int p = 0, k = 0, c = 0; p < n; ++p) {
      while (c == alpha[k]) {
        ++k;
        c = 0;
      }
      seg_of[p] = k;
      ++c;
    }
    for (int64_t i = 0; i < w->N; ++i) {
      int e[kMaxN];
      for (int p = 0; p < n; ++p) {
        T[i * n + p] = reps[i][p];
        Tinv[i * n + reps[i][p] - 1] = (int8_t)(p + 1);
        e[reps[i][p] - 1] = seg_of[p];  // chi_i(o) = omega^{sum_q e(q) o_q}, e(t_i(p)) = segment of p
      }
      int32_t f = 0;
      for (int q = 0; q < w->m; ++q) f = f * 3 + (all_orientations ? e[q] : ((e[q] - e[n - 1]) % 3 + 3) % 3);
      freq[i] = f;
    }
    // ---- Young subgroup S_alpha in itertools.product order, and R[h] = kron_k rho_{parts_k}(h_k) ----
    std::vector<std::vector<std::vector<int>>> seg_perms(3);
    std::vector<std::vector<double>> seg_mats(3);  // [a_k!][d_k * d_k] on the host
    std::vector<int> seg_dim(3, 1);
    int nontrivial = 0;
    double* single_dev = nullptr;
    for (int k = 0; k < 3; ++k) {
      const int a = alpha[k];
      std::vector<int> p(a);
      for (int i = 0; i < a; ++i) p[i] = i + 1;
      do seg_perms[k].push_back(p); while (std::next_permutation(p.begin(), p.end()));
      if (a == 0) {
        seg_mats[k] = {1.0};
        continue;
      }
      ++nontrivial;
    }
    w->nH = (int64_t)seg_perms[0].size() * seg_perms[1].size() * seg_perms[2].size();
    if (w->N * w->nH != w->nfact) throw std::runtime_error("coset count times subgroup order != n!");
    for (int k = 0; k < 3; ++k) {
      const int a = alpha[k];
      if (a == 0) continue;
      snfft_plan* fp = nullptr;
      if (snfft_plan_create(a, device, &fp) != SNFFT_OK) throw CudaError(g_last_error);
      w->factor_plans.push_back(fp);
      const auto& shapes = fp->host.levels[a].shapes;
      int irrep = -1;
      for (size_t i = 0; i < shapes.size(); ++i) {
        bool same = true;
        for (int q = 0; q < n; ++q) {
          const int want = parts[k * n + q];
          const int have = q < (int)shapes[i].rows.size() ? shapes[i].rows[q] : 0;
          if (want != have) same = false;
        }
        if (same) irrep = (int)i;
      }
      if (irrep < 0) throw InvalidArg("parts[" + std::to_string(k) + "] is not a partition of alpha[" + std::to_string(k) + "]");
      const int64_t d = shapes[irrep].dim, cnt = (int64_t)seg_perms[k].size();
      seg_dim[k] = (int)d;
      w->seg_d[k] = (int)d;
      if (a >= 2) w->ffts.push_back({fp, a, cnt, d, (int64_t)shapes[irrep].off});
      std::vector<int32_t> flat((size_t)cnt * a);
      for (int64_t r = 0; r < cnt; ++r)
        for (int q = 0; q < a; ++q) flat[r * a + q] = seg_perms[k][r][q];
      int32_t* perms_dev = w->upload(flat);
      double* mats_dev = w->alloc<double>((size_t)cnt * d * d);
      if (snfft_yor(fp, irrep, perms_dev, cnt, mats_dev, nullptr) != SNFFT_OK) throw CudaError(g_last_error);
      CUDA_OK(cudaDeviceSynchronize());
      if (nontrivial == 1) {
        single_dev = mats_dev;  // R is this factor's table itself
      } else {
        seg_mats[k].resize((size_t)cnt * d * d);
        CUDA_OK(cudaMemcpy(seg_mats[k].data(), mats_dev, seg_mats[k].size() * sizeof(double), cudaMemcpyDeviceToHost));
      }
    }
    w->block = (int64_t)seg_dim[0] * seg_dim[1] * seg_dim[2];
    w->D = w->N * w->block;
    const int64_t bb = w->block * w->block;
    if (nontrivial == 1) {
      w->R = single_dev;
    } else {
      std::vector<double> R((size_t)w->nH * bb);
      const int d0 = seg_dim[0], d1 = seg_dim[1], d2 = seg_dim[2];
      int64_t h = 0;
      for (size_t x = 0; x < seg_perms[0].size(); ++x)
        for (size_t y = 0; y < seg_perms[1].size(); ++y)
          for (size_t z = 0; z < seg_perms[2].size(); ++z, ++h) {
            const double* A = &seg_mats[0][x * d0 * d0];
            const double* B = &seg_mats[1][y * d1 * d1];
            const double* Cm = &seg_mats[2][z * d2 * d2];
            double* dst = &R[h * bb];
            for (int a0 = 0; a0 < d0; ++a0)
              for (int a1 = 0; a1 < d1; ++a1)
                for (int a2 = 0; a2 < d2; ++a2)
                  for (int c0 = 0; c0 < d0; ++c0)
                    for (int c1 = 0; c1 < d1; ++c1)
                      for (int c2 = 0; c2 < d2; ++c2)
                        dst[(((int64_t)a0 * d1 + a1) * d2 + a2) * w->block + ((int64_t)c0 * d1 + c1) * d2 + c2] =
                            A[a0 * d0 + c0] * B[a1 * d1 + c1] * Cm[a2 * d2 + c2];
          }
      w->R = w->upload(R);
    }
    w->RT = w->alloc<double>((size_t)w->nH * bb);
    CUDA_OK(launch_transpose(w->R, w->RT, w->nH, bb, nullptr));
    // ---- gather index sigma = t_i h t_j^-1 and its inverse, on the device ----
    std::vector<int8_t> H((size_t)w->nH * n);
    {
      int64_t h = 0;
      for (const auto& x : seg_perms[0])
        for (const auto& y : seg_perms[1])
          for (const auto& z : seg_perms[2]) {
            int at = 0;
            for (int v : x) H[h * n + at++] = (int8_t)v;
            for (int v : y) H[h * n + at++] = (int8_t)(alpha[0] + v);
            for (int v : z) H[h * n + at++] = (int8_t)(alpha[0] + alpha[1] + v);
            ++h;
          }
    }
    int8_t* T_dev = w->upload(T);
    int8_t* Tinv_dev = w->upload(Tinv);
    int8_t* H_dev = w->upload(H);
    w->freq = w->upload(freq);
    w->idx = w->alloc<int64_t>((size_t)w->nfact * w->N);
    w->inv_idx = w->alloc<int64_t>((size_t)w->nfact * w->N);
    CUDA_OK(launch_wreath_index(n, (int)w->N, (int)w->nH, T_dev, Tinv_dev, H_dev, w->idx, w->inv_idx, nullptr));
    if (wreath_stage2_fits((int)w->N, (int)w->block, (int)w->nH)) {
      if (w->nfact * w->N > 0x7fffffffLL) throw std::runtime_error("gather index does not fit 32 bits");
      w->idx_t = w->alloc<int32_t>((size_t)w->nfact * w->N);
      CUDA_OK(launch_transpose_i64(w->idx, w->idx_t, w->N * w->N, w->nH, nullptr));
    }
    // ---- workspace ----
    // F_i(sigma): two planes, or - for the fused stage 2 - the same memory as interleaved complex128
    w->Fre = w->alloc<double>(2 * (size_t)w->nfact * w->N);
    w->Fim = w->Fre + (size_t)w->nfact * w->N;
    w->G = w->alloc<double>((size_t)w->nfact * w->N);
    w->P = w->alloc<double>(2 * (size_t)w->N * w->N * bb);  // one plane of blocks (GEMM path) or complex blocks (fused adjoint)
    int a_max = 0;
    for (const auto& f : w->ffts) a_max = std::max(a_max, f.a);
    w->factor_fft = !w->idx_t && a_max >= 2 && !std::getenv("SNFFT_WREATH_NO_FACTOR_FFT");
    if (w->factor_fft) {
      w->Y = w->alloc<double>((size_t)w->nfact * w->N);
      w->Z = w->alloc<double>((size_t)w->nfact * w->N);
    }
    CUDA_OK(cudaDeviceSynchronize());
  });
  if (rc != SNFFT_OK) {
    if (w) snfft_wreath_destroy(w);
    return rc;
  }
  *out = w;
  return SNFFT_OK;
}

int snfft_wreath_destroy(snfft_wreath* w) {
  if (!w) return SNFFT_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  if (w->device >= 0) cudaSetDevice(w->device);
  for (void* p : w->allocs) cudaFree(p);
  for (snfft_plan* p : w->factor_plans) snfft_plan_destroy(p);
  if (prev >= 0) cudaSetDevice(prev);
  delete w;
  return SNFFT_OK;
}

int snfft_wreath_dims(const snfft_wreath* w, int64_t* dims) {
  return guarded([&] {
    if (!w || !dims) throw InvalidArg("null argument");
    dims[0] = w->D;
    dims[1] = w->N;
    dims[2] = w->block;
    dims[3] = w->nH;
    dims[4] = w->nfact;
    dims[5] = w->n_ori;
  });
}

// Stage 2 through the level recursion of the factors (see snfft_wreath::Factor).  Forward: one plane
// of F -> one plane of the D x D matrix; inverse: one plane of the matrix -> one plane of F with
// E[h] = <block, R[h]> (the inverse transforms return (d_k / a_k!) <rho_k, .> per factor: the
// product of a_k! / d_k = |S_alpha| / b rides on the block regrouping).
static void wreath_factor_forward(snfft_wreath* w, const double* F_plane, double* out_plane, cudaStream_t stream) {
  const int64_t pairs = w->N * w->N;
  double* bufs[3] = {w->G, w->Y, w->Z};
  CUDA_OK(launch_gather(F_plane, w->idx, bufs[0], pairs * w->nH, 1, stream));  // [pair][h1][h2][h3]
  int cur = 0;
  int64_t outer = pairs * w->nH, inner = 1;
  for (int q = (int)w->ffts.size() - 1; q >= 0; --q) {
    const snfft_wreath::Factor& f = w->ffts[q];
    outer /= f.fact;  // [outer][S_a][inner]
    int in = cur;
    if (inner > 1) {
      in = (cur + 1) % 3;
      CUDA_OK(launch_transpose_batched(bufs[cur], bufs[in], outer, f.fact, inner, stream));  // [outer][inner][S_a]
    }
    const int spec = (in + 1) % 3, work = (in + 2) % 3;
    forward_impl(f.plan, f.a, bufs[in], bufs[spec], bufs[work], outer * inner, true, stream);
    // the block of parts_k out of every packed spectrum: [outer][inner][d * d]
    CUDA_OK(cudaMemcpy2DAsync(bufs[in], (size_t)(f.d * f.d) * sizeof(double), bufs[spec] + f.off,
                              (size_t)f.fact * sizeof(double), (size_t)(f.d * f.d) * sizeof(double),
                              (size_t)(outer * inner), cudaMemcpyDeviceToDevice, stream));
    cur = in;
    inner *= f.d * f.d;
  }
  CUDA_OK(launch_wreath_kron_permute(pairs, w->seg_d[0], w->seg_d[1], w->seg_d[2], bufs[cur], w->P, true, stream));
  CUDA_OK(launch_wreath_permute((int)w->N, (int)w->block, out_plane, w->P, 1.0, true, stream));
}

static void wreath_factor_inverse(snfft_wreath* w, const double* mat_plane, double* F_plane, cudaStream_t stream) {
  const int64_t pairs = w->N * w->N, bb = w->block * w->block;
  double* bufs[3] = {w->G, w->Y, w->Z};
  CUDA_OK(launch_wreath_permute((int)w->N, (int)w->block, const_cast<double*>(mat_plane), w->P,
                                (double)w->nH / (double)w->block, false, stream));
  CUDA_OK(launch_wreath_kron_permute(pairs, w->seg_d[0], w->seg_d[1], w->seg_d[2], bufs[0], w->P, false, stream));
  int cur = 0;
  int64_t outer = pairs, inner = bb;  // [pair][e3][e2][e1]
  for (size_t q = 0; q < w->ffts.size(); ++q) {
    const snfft_wreath::Factor& f = w->ffts[q];
    inner /= f.d * f.d;  // [outer][inner][d * d]
    const int spec = (cur + 1) % 3, fn = (cur + 2) % 3;
    CUDA_OK(cudaMemsetAsync(bufs[spec], 0, (size_t)(outer * inner * f.fact) * sizeof(double), stream));
    CUDA_OK(cudaMemcpy2DAsync(bufs[spec] + f.off, (size_t)f.fact * sizeof(double), bufs[cur],
                              (size_t)(f.d * f.d) * sizeof(double), (size_t)(f.d * f.d) * sizeof(double),
                              (size_t)(outer * inner), cudaMemcpyDeviceToDevice, stream));
    inverse_impl(f.plan, f.a, bufs[spec], bufs[fn], bufs[cur], outer * inner, true, stream);  // [outer][inner][S_a]
    if (inner > 1) {
      CUDA_OK(launch_transpose_batched(bufs[fn], bufs[cur], outer, inner, f.fact, stream));  // [outer][S_a][inner]
    } else {
      cur = fn;
    }
    outer *= f.fact;
  }
  CUDA_OK(launch_gather(bufs[cur], w->inv_idx, F_plane, w->nfact * w->N, 1, stream));
}

int snfft_wreath_forward(snfft_wreath* w, const double* f_dev, double* out_re_dev, double* out_im_dev, void* stream_) {
  return guarded([&] {
    if (!w || !f_dev || !out_re_dev || !out_im_dev) throw InvalidArg("null argument");
    DeviceGuard guard(w->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    std::lock_guard<std::mutex> lock(w->mu);
    const int64_t bb = w->block * w->block, pairs = w->N * w->N;
    // stage 1: character sums over the orientations, F_i(sigma) (one pass over f)
    const bool fused = w->idx_t != nullptr;  // F interleaved (see snfft_wreath_create)
    CUDA_OK(launch_c3_dft(w->m, f_dev, nullptr, w->freq, (int)w->N, w->nfact, w->Fre, fused ? nullptr : w->Fim, false,
                          1.0, stream));
    // stage 2: per coset pair the sum over the Young subgroup.  Small blocks: one fused kernel;
    // otherwise the level recursion of the Young-subgroup factors (one coset, S_alpha = S_n: the S_n
    // transform of F itself); trivial subgroups: gather, one DMMA GEMM and the block placement
    if (fused) {
      CUDA_OK(launch_wreath_stage2((int)w->N, (int)w->block, (int)w->nH, w->nfact, w->idx_t, w->inv_idx, w->R, w->Fre,
                                   out_re_dev, out_im_dev, w->P, false, stream));
      return;
    }
    if (w->factor_fft) {
      wreath_factor_forward(w, w->Fre, out_re_dev, stream);
      wreath_factor_forward(w, w->Fim, out_im_dev, stream);
      return;
    }
    for (int plane = 0; plane < 2; ++plane) {
      CUDA_OK(launch_gather(plane ? w->Fim : w->Fre, w->idx, w->G, pairs * w->nH, 1, stream));
      CUDA_OK(launch_dgemm(w->G, w->R, w->P, (int)pairs, (int)bb, (int)w->nH, stream));
      CUDA_OK(launch_wreath_permute((int)w->N, (int)w->block, plane ? out_im_dev : out_re_dev, w->P, 1.0, true, stream));
    }
  });
}

int snfft_wreath_inverse(snfft_wreath* w, const double* fhat_re_dev, const double* fhat_im_dev, double* out_re_dev,
                         double* out_im_dev, void* stream_) {
  return guarded([&] {
    if (!w || !fhat_re_dev || !fhat_im_dev || !out_re_dev) throw InvalidArg("null argument");
    if (!out_im_dev) require_aligned16(out_re_dev, "out_re_dev (interleaved complex128 output)");
    DeviceGuard guard(w->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    std::lock_guard<std::mutex> lock(w->mu);
    const int64_t bb = w->block * w->block, pairs = w->N * w->N;
    if (w->factor_fft) {
      wreath_factor_inverse(w, fhat_re_dev, w->Fre, stream);
      wreath_factor_inverse(w, fhat_im_dev, w->Fim, stream);
    }
    for (int plane = 0; plane < 2 && !w->idx_t && !w->factor_fft; ++plane) {
      CUDA_OK(launch_wreath_permute((int)w->N, (int)w->block, const_cast<double*>(plane ? fhat_im_dev : fhat_re_dev), w->P,
                                    1.0, false, stream));
      CUDA_OK(launch_dgemm(w->P, w->RT, w->G, (int)pairs, (int)w->nH, (int)bb, stream));  // [pairs][nH]
      CUDA_OK(launch_gather(w->G, w->inv_idx, plane ? w->Fim : w->Fre, w->nfact * w->N, 1, stream));
    }
    if (w->idx_t)  // fused: F interleaved, written coalesced (one thread per element of F)
      CUDA_OK(launch_wreath_stage2((int)w->N, (int)w->block, (int)w->nH, w->nfact, w->idx_t, w->inv_idx, w->R, w->Fre,
                                   const_cast<double*>(fhat_re_dev), const_cast<double*>(fhat_im_dev), w->P, true, stream));
    // d tr(rho(g)^H f_hat) / |G| (cube_ift.py:56-67 divided by the group order, test_inv_fft.py:140):
    // the factor rides on the N values the adjoint DFT scatters onto its grid
    CUDA_OK(launch_c3_dft(w->m, w->Fre, w->idx_t ? nullptr : w->Fim, w->freq, (int)w->N, w->nfact, out_re_dev, out_im_dev, true,
                          (double)w->D / ((double)w->nfact * (double)w->n_ori), stream));
  });
}

int snfft_wreath_inverse_at(snfft_wreath* w, const double* fhat_re_dev, const double* fhat_im_dev,
                            const int64_t* sigma_rank_dev, const int32_t* ori_index_dev, int64_t count,
                            double* out_re_dev, double* out_im_dev, void* stream) {
  return guarded([&] {
    if (!w || !fhat_re_dev || !fhat_im_dev || !out_re_dev || !out_im_dev) throw InvalidArg("null argument");
    if (count < 0) throw InvalidArg("count must be >= 0");
    if (count > 0 && (!sigma_rank_dev || !ori_index_dev)) throw InvalidArg("null state list");
    DeviceGuard guard(w->device);
    // reads the plan's tables only: no workspace, no lock
    CUDA_OK(launch_wreath_inverse_at(w->m, (int)w->N, (int)w->block, (int)w->nH, w->nfact, (int)w->n_ori, w->inv_idx, w->R,
                                     fhat_re_dev, fhat_im_dev, w->freq, sigma_rank_dev, ori_index_dev, count,
                                     (double)w->D / ((double)w->nfact * (double)w->n_ori), out_re_dev, out_im_dev,
                                     (cudaStream_t)stream));
  });
}

int snfft_c3_dft(int m, const double* in_re_dev, const double* in_im_dev, const int32_t* freq_dev, int nfreq,
                 int64_t count, double* out_re_dev, double* out_im_dev, int inverse, void* stream) {
  return guarded([&] {
    if (!in_re_dev || !freq_dev || !out_re_dev) throw InvalidArg("null device buffer");
    if (!out_im_dev) require_aligned16(out_re_dev, "out_re_dev (interleaved complex128 output)");
    if (inverse && !in_im_dev) require_aligned16(in_re_dev, "in_re_dev (interleaved complex128 input)");
    if (m < 1 || m > 8) throw InvalidArg("m must be in 1..8 (the 3^m grid lives in shared memory)");
    if (nfreq < 1 || count < 0) throw InvalidArg("bad shape");
    ensure_kernels_prepared();
    CUDA_OK(launch_c3_dft(m, in_re_dev, in_im_dev, freq_dev, nfreq, count, out_re_dev, out_im_dev, inverse != 0,
                          1.0, (cudaStream_t)stream));
  });
}

int snfft_gather(const double* src_dev, const int64_t* idx_dev, double* dst_dev, int64_t count, int width,
                 void* stream) {
  return guarded([&] {
    if (!src_dev || !idx_dev || !dst_dev) throw InvalidArg("null device buffer");
    if (count < 0 || width < 1) throw InvalidArg("bad gather shape");
    CUDA_OK(launch_gather(src_dev, idx_dev, dst_dev, count, width, (cudaStream_t)stream));
  });
}

int snfft_export_phase(const snfft_plan* plan, int k, int phase, int32_t* sizes, int32_t* panels,
                       int64_t* groups, void* blob_fwd, void* blob_inv) {
  return guarded([&] {
    if (!plan || !sizes) throw InvalidArg("null argument");
    if (k < kPhaseMinLevel || k > plan->host.n) throw InvalidArg("row-group phases exist for levels 9..n");
    const LevelProgram& prog = plan->host.programs[k];
    const int nphases = (int)prog.panels[0].phases.size();
    if (phase < 1 || phase >= nphases) throw InvalidArg("phase must be in 1..phases-1 (phase 0 = the generated stages)");
    // the same builder as the device tables, "uploading" into host memory
    std::vector<std::vector<char>> keep;
    auto up = [&](const auto& v) {
      using T = typename std::decay<decltype(v)>::type::value_type;
      keep.emplace_back((const char*)v.data(), (const char*)(v.data() + v.size()));
      return (T*)keep.back().data();
    };
    std::vector<PhaseHost> devs;
    const LevelLayout lay = plain_layout(plan, k);
    build_phase_devs(prog, lay, up, devs);
    const PhaseDev& D = devs[phase].dev;
    int32_t ngroups = 0;
    for (int p = 0; p < D.npanels; ++p) ngroups += D.panels[p].ngroups;
    int64_t words = 0;
    for (int g = 0; g < ngroups; ++g) words = std::max<int64_t>(words, (int64_t)D.groups[g].blob_off + D.groups[g].blob_len);
    sizes[0] = D.npanels;
    sizes[1] = ngroups;
    sizes[2] = (int32_t)words;
    sizes[3] = D.max_rows;
    sizes[4] = D.max_blob;
    sizes[5] = D.nstages;
    sizes[6] = prog.panels[0].phases[phase].stage_lo;
    sizes[7] = nphases;
    for (int p = 0; p < D.npanels && panels; ++p) {
      const PhasePanelDev& pd = D.panels[p];
      const int32_t rec[5] = {pd.d, pd.tiles, pd.grp_off, pd.ngroups, pd.pack};
      std::copy(rec, rec + 5, panels + 5 * p);
    }
    for (int g = 0; g < ngroups && groups; ++g) {
      groups[4 * g] = D.groups[g].blob_off;
      groups[4 * g + 1] = D.groups[g].blob_len;
      groups[4 * g + 2] = D.groups[g].nrows & 0xffffu;
      groups[4 * g + 3] = D.groups[g].nrows >> 16;
    }
    if (blob_fwd) std::memcpy(blob_fwd, D.blob_f, (size_t)words * 16);
    if (blob_inv) std::memcpy(blob_inv, D.blob_i, (size_t)words * 16);
  });
}

int snfft_export_coef(const snfft_plan* plan, int k, double* coef, int64_t* count) {
  return guarded([&] {
    if (!plan || !count) throw InvalidArg("null argument");
    if (k < 2 || k > plan->host.n) throw InvalidArg("level k must be in 2..n");
    const std::vector<double>& pool = plan->host.programs[k].coef;
    *count = (int64_t)pool.size();
    if (coef) std::copy(pool.begin(), pool.end(), coef);
  });
}

int64_t snfft_plan_device_bytes(const snfft_plan* plan) { return plan ? plan->device_bytes : 0; }

double snfft_level_macs_per_element(const snfft_plan* plan, int k) {
  if (!plan || k < 2 || k > plan->host.n) return 0.0;
  return plan->host.programs[k].macs / (double)plan->host.fact[k];
}

int64_t snfft_launch_count(void) { return g_launch_count; }

int snfft_debug_force_table_driven(int on) {
  g_force_table_driven = on != 0;
  return SNFFT_OK;
}

int snfft_debug_fused_tma(int on) {
  g_fused_tma = on != 0;
  return SNFFT_OK;
}

int snfft_timing_enable(int on) {
  timing_enable(on != 0);
  return SNFFT_OK;
}

int snfft_timing_read(double* ms_by_kind, int64_t* launches_by_kind) {
  return guarded([&] {
    if (!ms_by_kind || !launches_by_kind) throw InvalidArg("null argument");
    timing_read(ms_by_kind, launches_by_kind);
  });
}

int snfft_export_panel(const snfft_plan* plan, int k, int panel, int32_t* n_blocks,
                       int32_t* block_m, int32_t* block_stage, int32_t* block_slots,
                       int32_t* block_final, double* coef_fwd, double* coef_inv, int64_t* dest,
                       int64_t* rowbase) {
  return guarded([&] {
    if (!plan) throw InvalidArg("plan is null");
    if (k < 2 || k > plan->host.n) throw InvalidArg("level k must be in 2..n");
    const LevelProgram& L = plan->host.programs[k];
    if (panel < 0 || panel >= (int)L.panels.size()) throw InvalidArg("panel out of range");
    const PanelProgram& P = L.panels[panel];
    if (n_blocks) *n_blocks = (int32_t)P.blocks.size();
    for (size_t b = 0; b < P.blocks.size(); ++b) {
      const Block& blk = P.blocks[b];
      if (block_m) block_m[b] = blk.m;
      if (block_stage) block_stage[b] = (int32_t)(b / P.d) + 1;
      if (block_final) block_final[b] = blk.final_mask;
      if (block_slots)
        for (int x = 0; x < kMaxBlock; ++x) block_slots[b * kMaxBlock + x] = x < blk.m ? blk.slot[x] : -1;
      for (int x = 0; x < 25; ++x) {
        if (coef_fwd) coef_fwd[b * 25 + x] = x < blk.m * blk.m ? L.coef[blk.coef_f + x] : 0.0;
        if (coef_inv) coef_inv[b * 25 + x] = x < blk.m * blk.m ? L.coef[blk.coef_i + x] : 0.0;
      }
    }
    for (size_t s = 0; s < P.dest.size(); ++s) {
      if (dest) dest[s] = P.dest[s];
      if (rowbase) rowbase[s] = P.rowbase[s];
    }
  });
}

}  // extern "C"
