// sm_100a kernels of the S_n transform.
//
// Every level of the recursion (fft.py:96-112) is executed as the staged block program built by
// plan.cpp: rows of a column block are combined by small dense blocks (2x2 .. 5x5), columns are
// a passive, contiguous dimension.  Two drivers share the block executor:
//   * fused_kernel  : levels 2..k0 (k0 <= 7) of one S_{k0} block entirely in shared memory, one
//                     HBM read and one HBM write for k0-1 levels;
//   * level_kernel  : one level k > k0, streaming column tiles of one panel through shared memory.
// fp64 has no tcgen05 kind, so the arithmetic is DFMA on the CUDA cores; the kernels are sized
// against HBM (16 B per element per pass) and shared-memory bandwidth.
#include "kernels.cuh"

#include <mutex>
#include <vector>

namespace snfft {

int64_t g_launch_count = 0;

namespace {
struct TimedLaunch {
  int kind;
  cudaEvent_t start, stop;
};
bool g_timing = false;
std::vector<TimedLaunch> g_timed;
std::vector<cudaEvent_t> g_event_pool;
std::mutex g_timing_mu;

cudaEvent_t get_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

// brackets one launch with events when timing is on
struct LaunchScope {
  cudaStream_t stream;
  bool on;
  TimedLaunch t;
  LaunchScope(int kind, cudaStream_t s) : stream(s), on(g_timing) {
    ++g_launch_count;
    if (on) {
      std::lock_guard<std::mutex> lock(g_timing_mu);
      t.kind = kind;
      t.start = get_event();
      t.stop = get_event();
      cudaEventRecord(t.start, stream);
    }
  }
  ~LaunchScope() {
    if (on) {
      cudaEventRecord(t.stop, stream);
      std::lock_guard<std::mutex> lock(g_timing_mu);
      g_timed.push_back(t);
    }
  }
};
}  // namespace

void timing_enable(bool on) { g_timing = on; }

void timing_read(double* ms_by_kind, int64_t* launches_by_kind) {
  std::lock_guard<std::mutex> lock(g_timing_mu);
  for (int i = 0; i < kNumKinds; ++i) {
    ms_by_kind[i] = 0.0;
    launches_by_kind[i] = 0;
  }
  for (TimedLaunch& t : g_timed) {
    cudaEventSynchronize(t.stop);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t.start, t.stop) == cudaSuccess) {
      ms_by_kind[t.kind] += ms;
      launches_by_kind[t.kind] += 1;
    }
    g_event_pool.push_back(t.start);
    g_event_pool.push_back(t.stop);
  }
  g_timed.clear();
}

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------
// block executor: y = C x (forward) or y = C^T x (inverse) on m rows addressed by p[b]
// ---------------------------------------------------------------------------------------------
template <int M, bool INV>
__device__ __forceinline__ void apply_block(double* __restrict__ base, const int* off,
                                            const double* __restrict__ cm) {
  double x[M], y[M];
#pragma unroll
  for (int b = 0; b < M; ++b) x[b] = base[off[b]];
#pragma unroll
  for (int a = 0; a < M; ++a) {
    double acc = 0.0;
#pragma unroll
    for (int b = 0; b < M; ++b) {
      const double c = INV ? __ldg(cm + b * M + a) : __ldg(cm + a * M + b);
      acc = fma(c, x[b], acc);
    }
    y[a] = acc;
  }
#pragma unroll
  for (int a = 0; a < M; ++a) base[off[a]] = y[a];
}

template <bool INV>
__device__ __forceinline__ void dispatch_block(int m, double* base, const int* off,
                                               const double* cm) {
  switch (m) {
    case 2: apply_block<2, INV>(base, off, cm); break;
    case 3: apply_block<3, INV>(base, off, cm); break;
    case 4: apply_block<4, INV>(base, off, cm); break;
    default: apply_block<5, INV>(base, off, cm); break;
  }
}

// ---------------------------------------------------------------------------------------------
// streaming level kernel
// grid.x = tiles of the level, grid.y = group chunks (grid-stride over groups)
// ---------------------------------------------------------------------------------------------
template <bool INV>
__global__ void __launch_bounds__(kThreads)
level_kernel(LevelDev L, const Tile* __restrict__ tiles, const double* __restrict__ in,
             double* __restrict__ out, int64_t groups) {
  extern __shared__ double buf[];
  const Tile t = tiles[blockIdx.x];
  const PanelDev P = L.panels[t.panel];
  const int d = P.d, rows = P.rows, w = t.w, k = L.k;
  const uint32_t* __restrict__ dest = L.dest + P.slot_off;
  const double* __restrict__ fscale = L.fscale + P.slot_off;
  const double* __restrict__ iscale = L.iscale + P.slot_off;
  const int tile_elems = rows * w;
  // forward reads the k packed S_{k-1} spectra of a group and writes the packed S_k spectrum;
  // the inverse swaps the roles of the two buffers.
  const double* __restrict__ src = in;
  double* __restrict__ dst = out;

  for (int64_t g0 = (int64_t)blockIdx.y * t.cap; g0 < groups; g0 += (int64_t)gridDim.y * t.cap) {
    const int ng = (int)min((int64_t)t.cap, groups - g0);
    // ---- load --------------------------------------------------------------------------------
    for (int e = threadIdx.x; e < ng * tile_elems; e += kThreads) {
      const int gb = e / tile_elems;
      const int r = e - gb * tile_elems;
      const int s = r / w;
      const int c = r - s * w;
      const int64_t gbase = (g0 + gb) * (int64_t)L.kfact;
      double v;
      if (!INV) {
        const int i = s / d;
        const int u = s - i * d;
        v = src[gbase + (int64_t)i * L.km1fact + P.in_off + u * d + t.c0 + c];
      } else {
        v = iscale[s] * src[gbase + dest[s] + t.c0 + c];
      }
      buf[e] = v;
    }
    __syncthreads();
    // ---- stages ------------------------------------------------------------------------------
    for (int st = 0; st < k - 1; ++st) {
      const int stage = INV ? (k - 2 - st) : st;
      const Block* __restrict__ blocks = L.blocks + P.blk_off + stage * d;
      const int items = ng * d * w;
      for (int e = threadIdx.x; e < items; e += kThreads) {
        const int c = e % w;
        const int tq = e / w;
        const int q = tq % d;
        const int gb = tq / d;
        const Block b = blocks[q];
        int off[kMaxBlock];
#pragma unroll
        for (int x = 0; x < kMaxBlock; ++x) off[x] = (int)b.slot[x] * w;
        dispatch_block<INV>(b.m, buf + gb * tile_elems + c, off, L.coef + b.coef);
      }
      __syncthreads();
    }
    // ---- store -------------------------------------------------------------------------------
    for (int e = threadIdx.x; e < ng * tile_elems; e += kThreads) {
      const int gb = e / tile_elems;
      const int r = e - gb * tile_elems;
      const int s = r / w;
      const int c = r - s * w;
      const int64_t gbase = (g0 + gb) * (int64_t)L.kfact;
      if (!INV) {
        dst[gbase + dest[s] + t.c0 + c] = fscale[s] * buf[e];
      } else {
        const int i = s / d;
        const int u = s - i * d;
        dst[gbase + (int64_t)i * L.km1fact + P.in_off + u * d + t.c0 + c] = buf[e];
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// fused low levels: one CTA per S_{k0} block, two ping-pong buffers of k0! doubles
// ---------------------------------------------------------------------------------------------
template <bool INV>
__device__ __forceinline__ void fused_stage(const LevelDev& L, int stage, double* __restrict__ A,
                                            int nsub) {
  const int items = nsub * L.km1fact;
  for (int e = threadIdx.x; e < items; e += kThreads) {
    const int sub = e / L.km1fact;
    const int r = e - sub * L.km1fact;
    const uint32_t it = __ldg(L.item_tab + r);
    const PanelDev P = L.panels[it >> 24];
    const int q = (it >> 12) & 0xfff;
    const int c = it & 0xfff;
    const Block b = L.blocks[P.blk_off + stage * P.d + q];
    const uint32_t* __restrict__ rowbase = L.rowbase + P.slot_off;
    int off[kMaxBlock];
#pragma unroll
    for (int x = 0; x < kMaxBlock; ++x) off[x] = x < b.m ? (int)__ldg(rowbase + b.slot[x]) : 0;
    dispatch_block<INV>(b.m, A + sub * L.kfact + c, off, L.coef + b.coef);
  }
}

template <bool INV>
__global__ void __launch_bounds__(kThreads)
fused_kernel(FusedDev F, const double* in, double* out, int64_t groups) {  // in may alias out
  extern __shared__ double sm[];
  const int n0 = F.k0fact;
  double* A = sm;
  double* B = sm + n0;
  for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
    const double* gin = in + g * n0;
    double* gout = out + g * n0;
    for (int e = threadIdx.x; e < n0; e += kThreads) A[e] = gin[e];
    __syncthreads();
    if (!INV) {
      for (int k = 2; k <= F.k0; ++k) {
        const LevelDev& L = F.level[k];
        const int nsub = n0 / L.kfact;
        for (int st = 0; st < k - 1; ++st) {
          fused_stage<false>(L, st, A, nsub);
          __syncthreads();
        }
        // packed layout of level k: B[sub][copy_dest[r]] = fscale[r] * A[sub][r]
        for (int e = threadIdx.x; e < n0; e += kThreads) {
          const int sub = e / L.kfact;
          const int r = e - sub * L.kfact;
          B[sub * L.kfact + __ldg(L.copy_dest + r)] = __ldg(L.copy_fscale + r) * A[e];
        }
        __syncthreads();
        double* tmp = A;
        A = B;
        B = tmp;
      }
    } else {
      for (int k = F.k0; k >= 2; --k) {
        const LevelDev& L = F.level[k];
        const int nsub = n0 / L.kfact;
        for (int e = threadIdx.x; e < n0; e += kThreads) {
          const int sub = e / L.kfact;
          const int r = e - sub * L.kfact;
          B[e] = __ldg(L.copy_iscale + r) * A[sub * L.kfact + __ldg(L.copy_dest + r)];
        }
        __syncthreads();
        for (int st = k - 2; st >= 0; --st) {
          fused_stage<true>(L, st, B, nsub);
          __syncthreads();
        }
        double* tmp = A;
        A = B;
        B = tmp;
      }
    }
    for (int e = threadIdx.x; e < n0; e += kThreads) gout[e] = A[e];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// element order
// ---------------------------------------------------------------------------------------------
struct Facts {
  int64_t f[kMaxN + 1];
};

__device__ __forceinline__ int64_t coset_to_lex_dev(int n, int64_t flat, const Facts& F) {
  int p[kMaxN + 1];
#pragma unroll
  for (int x = 1; x <= kMaxN; ++x) p[x] = x;
  for (int k = n; k >= 2; --k) {
    const int i = (int)((flat / F.f[k - 1]) % k) + 1;
    const int first = p[i];
    for (int x = i; x < k; ++x) p[x] = p[x + 1];
    p[k] = first;
  }
  int64_t rank = 0;
  for (int a = 1; a <= n; ++a) {
    int smaller = 0;
    for (int b = a + 1; b <= n; ++b) smaller += p[b] < p[a];
    rank += smaller * F.f[n - a];
  }
  return rank;
}

__global__ void build_lex_table_kernel(int n, Facts F, uint32_t* __restrict__ table) {
  const int64_t total = F.f[n];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x)
    table[t] = (uint32_t)coset_to_lex_dev(n, t, F);
}

template <bool TO_COSET>
__global__ void reorder_kernel(int n, Facts F, const uint32_t* __restrict__ table,
                               const double* __restrict__ src, double* __restrict__ dst,
                               int64_t batch) {
  const int64_t total = F.f[n];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lex = table ? (int64_t)table[t] : coset_to_lex_dev(n, t, F);
    for (int64_t b = blockIdx.y; b < batch; b += gridDim.y) {
      if (TO_COSET)
        dst[b * total + t] = src[b * total + lex];
      else
        dst[b * total + lex] = src[b * total + t];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// rho_lambda(sigma): apply the generator word to the identity, one CTA per permutation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
yor_kernel(int n, int d, const int32_t* __restrict__ partner, const double* __restrict__ diag,
           const int32_t* __restrict__ perms, double* __restrict__ out) {
  __shared__ int word[kMaxN * kMaxN];
  __shared__ int wlen;
  const int64_t pid = blockIdx.x;
  double* __restrict__ M = out + pid * (int64_t)d * d;
  if (threadIdx.x == 0) {
    // bubble-sort word of the permutation map, as yor.py:19-38 does for a cycle
    int p[kMaxN];
    for (int i = 0; i < n; ++i) p[i] = perms[pid * n + i];
    int len = 0;
    for (int i = 0; i < n; ++i)
      for (int j = n - 1; j > i; --j)
        if (p[j] < p[j - 1]) {
          const int tmp = p[j];
          p[j] = p[j - 1];
          p[j - 1] = tmp;
          word[len++] = j;  // the transposition (j, j+1), 1-based
        }
    wlen = len;
  }
  for (int64_t e = threadIdx.x; e < (int64_t)d * d; e += kThreads) M[e] = (e / d == e % d) ? 1.0 : 0.0;
  __syncthreads();
  // rho(sigma) = rho(t_L) ... rho(t_1) read in the order the swaps were recorded reversed, i.e.
  // the product "reversed(factors)" left to right; applied to I from the rightmost factor, which
  // is the first recorded swap.
  for (int t = 0; t < wlen; ++t) {
    const int j = word[t];
    const int32_t* __restrict__ pr = partner + (int64_t)(j - 1) * d;
    const double* __restrict__ dg = diag + (int64_t)(j - 1) * d;
    for (int64_t e = threadIdx.x; e < (int64_t)d * d; e += kThreads) {
      const int r = (int)(e / d);
      const int c = (int)(e - (int64_t)r * d);
      const int q = pr[r];
      const double a = dg[r];
      if (q < 0) {
        M[e] = a * M[e];
      } else if (q > r) {
        const double b = sqrt(1 - a * a);
        const double x = M[(int64_t)r * d + c], y = M[(int64_t)q * d + c];
        M[(int64_t)r * d + c] = a * x + b * y;
        M[(int64_t)q * d + c] = b * x - a * y;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// fp64 GEMM on the DMMA pipe: C[M,N] = A[M,K] B[K,N], row-major.  Each warp owns a 16x16 tile
// of C as 2x2 mma.m8n8k4 fragments; a CTA of 8 warps covers 64x32... (simple, oracle-sized).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

constexpr int kGemmBM = 64, kGemmBN = 64, kGemmBK = 16;

__global__ void __launch_bounds__(256)
dgemm_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
             int M, int N, int K) {
  __shared__ double sA[kGemmBM][kGemmBK + 1];
  __shared__ double sB[kGemmBK][kGemmBN + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;
  // 8 warps: 4 along M (16 rows each) x 2 along N (32 cols each); per warp 2 x 4 fragments
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int grp = lane >> 2, tig = lane & 3;  // fragment coordinates of mma.m8n8k4
  for (int k0 = 0; k0 < K; k0 += kGemmBK) {
    for (int e = threadIdx.x; e < kGemmBM * kGemmBK; e += 256) {
      const int r = e / kGemmBK, c = e % kGemmBK;
      sA[r][c] = (m0 + r < M && k0 + c < K) ? A[(int64_t)(m0 + r) * K + k0 + c] : 0.0;
    }
    for (int e = threadIdx.x; e < kGemmBK * kGemmBN; e += 256) {
      const int r = e / kGemmBN, c = e % kGemmBN;
      sB[r][c] = (k0 + r < K && n0 + c < N) ? B[(int64_t)(k0 + r) * N + n0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGemmBK; kk += 4) {
      double a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = sA[wm + i * 8 + grp][kk + tig];  // A frag: row grp, col tig
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk + tig][wn + j * 8 + grp];  // B frag: row tig, col grp
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
  // C frag: row grp, cols 2*tig, 2*tig+1
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = m0 + wm + i * 8 + grp;
      const int c = n0 + wn + j * 8 + 2 * tig;
      if (r < M) {
        if (c < N) C[(int64_t)r * N + c] = acc[i][j][0];
        if (c + 1 < N) C[(int64_t)r * N + c + 1] = acc[i][j][1];
      }
    }
}

Facts make_facts() {
  Facts F;
  F.f[0] = 1;
  for (int i = 1; i <= kMaxN; ++i) F.f[i] = F.f[i - 1] * i;
  return F;
}

}  // namespace

cudaError_t prepare_kernels() {
  cudaError_t e;
  const int big = 227 * 1024;
  if ((e = cudaFuncSetAttribute(level_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(level_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t launch_level(const LevelDev& L, const Tile* tiles_dev, int ntiles, size_t smem_bytes,
                         const double* in, double* out, int64_t groups, bool inverse, int num_sms,
                         cudaStream_t stream) {
  if (groups <= 0 || ntiles <= 0) return cudaSuccess;
  // enough CTAs to fill the machine a few times over; each CTA strides over group chunks
  int64_t y = (8LL * num_sms + ntiles - 1) / ntiles;
  if (y > groups) y = groups;
  if (y < 1) y = 1;
  if (y > 65535) y = 65535;
  dim3 grid((unsigned)ntiles, (unsigned)y);
  LaunchScope scope(inverse ? kKindLevelInv : kKindLevelFwd, stream);
  if (inverse)
    level_kernel<true><<<grid, kThreads, smem_bytes, stream>>>(L, tiles_dev, in, out, groups);
  else
    level_kernel<false><<<grid, kThreads, smem_bytes, stream>>>(L, tiles_dev, in, out, groups);
  return cudaGetLastError();
}

cudaError_t launch_fused(const FusedDev& F, const double* in, double* out, int64_t groups,
                         bool inverse, int num_sms, cudaStream_t stream) {
  if (groups <= 0) return cudaSuccess;
  const size_t smem = 2 * (size_t)F.k0fact * sizeof(double);
  int64_t per_sm = smem > 0 ? (200 * 1024) / (int64_t)smem : 8;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int64_t grid = (int64_t)num_sms * per_sm;
  if (grid > groups) grid = groups;
  LaunchScope scope(inverse ? kKindFusedInv : kKindFusedFwd, stream);
  if (inverse)
    fused_kernel<true><<<(unsigned)grid, kThreads, smem, stream>>>(F, in, out, groups);
  else
    fused_kernel<false><<<(unsigned)grid, kThreads, smem, stream>>>(F, in, out, groups);
  return cudaGetLastError();
}

cudaError_t launch_build_lex_table(int n, uint32_t* table, cudaStream_t stream) {
  const Facts F = make_facts();
  const int64_t total = F.f[n];
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  LaunchScope scope(kKindOther, stream);
  build_lex_table_kernel<<<blocks, 256, 0, stream>>>(n, F, table);
  return cudaGetLastError();
}

cudaError_t launch_reorder(int n, const uint32_t* table, const double* src, double* dst,
                           int64_t batch, bool to_coset, cudaStream_t stream) {
  if (batch <= 0) return cudaSuccess;
  const Facts F = make_facts();
  const int64_t total = F.f[n];
  int64_t bx = (total + 255) / 256;
  if (bx > 2048) bx = 2048;
  int64_t by = batch < 64 ? batch : 64;
  dim3 grid((unsigned)bx, (unsigned)by);
  LaunchScope scope(kKindReorder, stream);
  if (to_coset)
    reorder_kernel<true><<<grid, 256, 0, stream>>>(n, F, table, src, dst, batch);
  else
    reorder_kernel<false><<<grid, 256, 0, stream>>>(n, F, table, src, dst, batch);
  return cudaGetLastError();
}

cudaError_t launch_yor(int n, int d, const int32_t* partner, const double* diag,
                       const int32_t* perms, int64_t count, double* out, cudaStream_t stream) {
  if (count <= 0) return cudaSuccess;
  LaunchScope scope(kKindYor, stream);
  yor_kernel<<<(unsigned)count, kThreads, 0, stream>>>(n, d, partner, diag, perms, out);
  return cudaGetLastError();
}

cudaError_t launch_dgemm(const double* A, const double* B, double* C, int M, int N, int K,
                         cudaStream_t stream) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  dim3 grid((N + kGemmBN - 1) / kGemmBN, (M + kGemmBM - 1) / kGemmBM);
  LaunchScope scope(kKindDgemm, stream);
  dgemm_kernel<<<grid, 256, 0, stream>>>(A, B, C, M, N, K);
  return cudaGetLastError();
}

}  // namespace snfft
