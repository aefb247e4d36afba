// C ABI of snfft_b200 (include/snfft_b200.h): plan management, table upload, transform drivers.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <map>
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
  uint32_t* lex_table = nullptr;
  std::mutex lazy_mu;
  std::map<int, YorTables> yor_tables;
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

void plan_tiles(const LevelProgram& prog, std::vector<Tile>& tiles, size_t& smem_bytes) {
  // column tiling of every panel: whole panels when they fit in 48 KB, else column strips of up to
  // 96 KB (the widest strip that fits when even one column exceeds that).
  const int64_t kSmall = 6144, kLarge = 12288, kHard = (227 * 1024) / 8;
  int64_t level_elems = kSmall;
  for (size_t p = 0; p < prog.panels.size(); ++p) {
    const PanelProgram& P = prog.panels[p];
    const int64_t rows = (int64_t)P.k * P.d;
    int64_t w;
    if (rows * P.d <= kSmall) {
      w = P.d;
    } else {
      w = std::max<int64_t>(1, std::min<int64_t>(P.d, kLarge / rows));
      const int64_t nt = (P.d + w - 1) / w;
      w = (P.d + nt - 1) / nt;
    }
    if (rows * w > kHard) throw std::runtime_error("panel column does not fit in shared memory");
    level_elems = std::max(level_elems, rows * w);
    for (int64_t c0 = 0; c0 < P.d; c0 += w)
      tiles.push_back(Tile{(int32_t)p, (int32_t)c0, (int32_t)std::min<int64_t>(w, P.d - c0), 1});
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
  plan_tiles(prog, tiles, plan->level_smem[k]);
  plan->ntiles[k] = (int)tiles.size();
  plan->tiles_dev[k] = plan->upload(tiles);
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

const ShapeInfo& shape_of(const snfft_plan* plan, int k, int irrep) {
  if (!plan) throw InvalidArg("plan is null");
  if (k < 1 || k > plan->host.n) throw InvalidArg("k out of range");
  if (irrep < 0 || irrep >= (int)plan->host.levels[k].shapes.size()) throw InvalidArg("irrep out of range");
  return plan->host.levels[k].shapes[irrep];
}

void run_level(const snfft_plan* plan, int k, const double* in, double* out, int64_t groups,
               bool inverse, cudaStream_t stream) {
  CUDA_OK(launch_level(plan->level_dev[k], plan->tiles_dev[k], plan->ntiles[k], plan->level_smem[k],
                       in, out, groups, inverse, plan->num_sms, stream));
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

extern "C" {

int snfft_abi_version(void) { return 1; }

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
      CUDA_OK(prepare_kernels());
      for (int k = 2; k <= n; ++k) upload_level(plan, k);
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

int snfft_forward(const snfft_plan* plan, const double* f_dev, double* fhat_dev, double* work_dev,
                  int64_t batch, int order, void* stream_) {
  return guarded([&] {
    require_device(plan);
    if (batch < 0) throw InvalidArg("batch < 0");
    if (batch == 0) return;
    if (!f_dev || !fhat_dev || !work_dev) throw InvalidArg("null device buffer");
    if (f_dev == fhat_dev || f_dev == work_dev || fhat_dev == work_dev)
      throw InvalidArg("f, fhat and work must be distinct buffers");
    if (order != SNFFT_ORDER_COSET && order != SNFFT_ORDER_LEX) throw InvalidArg("bad order flag");
    DeviceGuard guard(plan->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n = plan->host.n, k0 = plan->fused.k0;
    const int64_t nf = plan->host.fact[n];
    if (n == 1) {
      CUDA_OK(cudaMemcpyAsync(fhat_dev, f_dev, (size_t)batch * sizeof(double), cudaMemcpyDeviceToDevice, stream));
      return;
    }
    // passes: [fused 2..k0 if n >= 4] then one streaming pass per level k0+1..n.  Outputs alternate
    // between work and fhat so that the last one lands in fhat; f is never written.
    const int npasses = (k0 >= 4 ? 1 : 0) + (n - k0);
    auto out_of = [&](int pass) { return ((npasses - 1 - pass) % 2 == 0) ? fhat_dev : work_dev; };
    const double* cur = f_dev;
    if (order == SNFFT_ORDER_LEX) {
      // reorder into the buffer the first pass does not write (a level pass cannot run in place)
      double* tmp = out_of(0) == fhat_dev ? work_dev : fhat_dev;
      CUDA_OK(launch_reorder(n, plan->lex_table, f_dev, tmp, batch, true, stream));
      cur = tmp;
    }
    int pass = 0;
    if (k0 >= 4) {
      double* nxt = out_of(pass++);
      CUDA_OK(launch_fused(plan->fused, cur, nxt, batch * (nf / plan->fused.k0fact), false,
                           plan->num_sms, stream));
      cur = nxt;
    }
    for (int k = k0 + 1; k <= n; ++k) {
      double* nxt = out_of(pass++);
      run_level(plan, k, cur, nxt, batch * (nf / plan->host.fact[k]), false, stream);
      cur = nxt;
    }
  });
}

int snfft_inverse(const snfft_plan* plan, const double* fhat_dev, double* f_dev, double* work_dev,
                  int64_t batch, int order, void* stream_) {
  return guarded([&] {
    require_device(plan);
    if (batch < 0) throw InvalidArg("batch < 0");
    if (batch == 0) return;
    if (!f_dev || !fhat_dev || !work_dev) throw InvalidArg("null device buffer");
    if (f_dev == fhat_dev || f_dev == work_dev || fhat_dev == work_dev)
      throw InvalidArg("f, fhat and work must be distinct buffers");
    if (order != SNFFT_ORDER_COSET && order != SNFFT_ORDER_LEX) throw InvalidArg("bad order flag");
    DeviceGuard guard(plan->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n = plan->host.n, k0 = plan->fused.k0;
    const int64_t nf = plan->host.fact[n];
    if (n == 1) {
      CUDA_OK(cudaMemcpyAsync(f_dev, fhat_dev, (size_t)batch * sizeof(double), cudaMemcpyDeviceToDevice, stream));
      return;
    }
    const bool lex = order == SNFFT_ORDER_LEX;
    // passes: streaming levels n..k0+1, then the fused kernel (levels k0..2).  The last pass writes
    // `target` (work when a reorder to lexicographic order follows, else f); earlier outputs
    // alternate; fhat is never written.
    const int npasses = (k0 >= 4 ? 1 : 0) + (n - k0);
    double* target = lex ? work_dev : f_dev;
    double* other = lex ? f_dev : work_dev;
    auto out_of = [&](int pass) { return ((npasses - 1 - pass) % 2 == 0) ? target : other; };
    const double* cur = fhat_dev;
    int pass = 0;
    for (int k = n; k > k0; --k) {
      double* nxt = out_of(pass++);
      run_level(plan, k, cur, nxt, batch * (nf / plan->host.fact[k]), true, stream);
      cur = nxt;
    }
    if (k0 >= 4) {
      double* nxt = out_of(pass++);
      CUDA_OK(launch_fused(plan->fused, cur, nxt, batch * (nf / plan->fused.k0fact), true,
                           plan->num_sms, stream));
      cur = nxt;
    }
    if (lex) CUDA_OK(launch_reorder(n, plan->lex_table, work_dev, f_dev, batch, false, stream));
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

int snfft_yor(const snfft_plan* plan_, int irrep, const int32_t* perms_dev, int64_t count,
              double* out_dev, void* stream) {
  snfft_plan* plan = const_cast<snfft_plan*>(plan_);  // lazily cached generator tables
  return guarded([&] {
    require_device(plan);
    const ShapeInfo& s = shape_of(plan, plan->host.n, irrep);
    if (count < 0) throw InvalidArg("count < 0");
    if (count == 0) return;
    if (!perms_dev || !out_dev) throw InvalidArg("null device buffer");
    DeviceGuard guard(plan->device);
    if (plan->host.n == 1) {
      std::vector<double> ones((size_t)count, 1.0);
      CUDA_OK(cudaMemcpyAsync(out_dev, ones.data(), ones.size() * sizeof(double),
                              cudaMemcpyHostToDevice, (cudaStream_t)stream));
      CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
      return;
    }
    YorTables& t = yor_tables_for(plan, irrep);
    CUDA_OK(launch_yor(plan->host.n, (int)s.dim, t.partner, t.diag, perms_dev, count, out_dev,
                       (cudaStream_t)stream));
  });
}

int snfft_dft_direct(snfft_plan* plan, const double* f_dev, double* fhat_dev, int64_t batch,
                     void* stream_) {
  return guarded([&] {
    require_device(plan);
    const int n = plan->host.n;
    if (n > 8) throw InvalidArg("snfft_dft_direct supports n <= 8 (the [n!, n!] matrix is 13 GB at n = 8)");
    if (batch < 0) throw InvalidArg("batch < 0");
    if (batch == 0) return;
    if (!f_dev || !fhat_dev || f_dev == fhat_dev) throw InvalidArg("bad buffers");
    DeviceGuard guard(plan->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t nf = plan->host.fact[n];
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
      int32_t* perms_dev = nullptr;
      double* tmp = nullptr;
      CUDA_OK(cudaMalloc((void**)&perms_dev, perms.size() * sizeof(int32_t)));
      CUDA_OK(cudaMemcpy(perms_dev, perms.data(), perms.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      double* R = (double*)plan->alloc((size_t)nf * nf * sizeof(double));
      int64_t dmax = 0;
      for (const ShapeInfo& s : plan->host.levels[n].shapes) dmax = std::max(dmax, s.dim);
      CUDA_OK(cudaMalloc((void**)&tmp, (size_t)nf * dmax * dmax * sizeof(double)));
      for (size_t irrep = 0; irrep < plan->host.levels[n].shapes.size(); ++irrep) {
        const ShapeInfo& s = plan->host.levels[n].shapes[irrep];
        const int rc = snfft_yor(plan, (int)irrep, perms_dev, nf, tmp, stream);
        if (rc != SNFFT_OK) throw CudaError(g_last_error);
        CUDA_OK(cudaMemcpy2DAsync(R + s.off, (size_t)nf * sizeof(double), tmp,
                                  (size_t)(s.dim * s.dim) * sizeof(double),
                                  (size_t)(s.dim * s.dim) * sizeof(double), (size_t)nf,
                                  cudaMemcpyDeviceToDevice, stream));
      }
      CUDA_OK(cudaStreamSynchronize(stream));
      cudaFree(tmp);
      cudaFree(perms_dev);
      plan->dft_matrix = R;
    }
    CUDA_OK(launch_dgemm(f_dev, plan->dft_matrix, fhat_dev, (int)batch, (int)nf, (int)nf, stream));
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
