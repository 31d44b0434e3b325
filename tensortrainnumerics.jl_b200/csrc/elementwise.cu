// Bandwidth-bound helpers (kernel families F7/F8 of SURVEY.md §2.1): strided copies that stand in for
// the reference's permutedims / zero-padding / slicing (dmrg.jl:39-43,198,222; tdvp.jl:54-55,148;
// mals.jl:104-109), diagonal scalings by singular values (tt_tools.jl:754-757, dmrg.jl:203-206,227-230),
// and the vector operations + reductions that surround the effective-operator matvec inside the Krylov
// drivers (dmrg.jl:170,245; tdvp.jl:75).  All of them are coalesced along the fastest index and sized
// as a multiple of the SM count with a grid-stride loop.
#include "ttn_internal.h"
#include <cstring>

namespace ttn {

// ------------------------------------------------------------------------------------------------
// context + stream-ordered allocator
// ------------------------------------------------------------------------------------------------
// One context per HOST THREAD: every thread that calls the library owns a compute stream, a copy stream, an allocation cache
// and its own counters (ttn_init must be called in each of them).  Two threads working on independent trains therefore
// overlap on the GPU: the latency-bound eigensolver kernels of one chunk run beside the DMMA GEMMs of another (bench.py).
// Handles must be used and released by the thread that created them; read-only sharing of uploaded operators is fine.
Context& ctx() {
  static thread_local Context c;
  return c;
}

// Exact-size cache in front of the stream-ordered pool.  Sweeps allocate the same buffer sizes over and over (cores,
// Theta, workspaces); for large blocks the driver pool can answer by mapping fresh physical memory when its free space is
// fragmented (measured: up to 1 s of idle GPU per cfg5 chunk).  Everything runs on ONE stream, so a block released here
// may be handed out again at once: its previous users are ahead in the same stream.  (One cache per host thread / stream.)
namespace {
struct BlockCache {
  std::vector<std::pair<size_t, void*>> free_blocks;   // small, linear search (a few dozen entries)
  size_t cached_bytes = 0;
  static constexpr size_t kMinBytes = 256 * 1024;
  static constexpr size_t kMaxCached = (size_t)48 << 30;
  void* take(size_t b) {
    for (size_t i = 0; i < free_blocks.size(); ++i)
      if (free_blocks[i].first == b) {
        void* p = free_blocks[i].second;
        free_blocks[i] = free_blocks.back();
        free_blocks.pop_back();
        cached_bytes -= b;
        return p;
      }
    return nullptr;
  }
  void give(size_t b, void* p) {
    if (cached_bytes + b > kMaxCached || free_blocks.size() >= 512) trim();
    free_blocks.emplace_back(b, p);
    cached_bytes += b;
  }
  void trim() {
    for (auto& e : free_blocks) cudaFreeAsync(e.second, ctx().stream);
    free_blocks.clear();
    cached_bytes = 0;
  }
};
BlockCache& cache() {
  static thread_local BlockCache c;   // per host thread, like the stream it is ordered on
  return c;
}
}  // namespace

void devbuf_cache_trim() { cache().trim(); }

void read_back(void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return;
  if (bytes > ((size_t)1 << 20)) {                   // bulk results: straight into the caller's memory
    TTN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx().stream));
    TTN_CUDA(cudaStreamSynchronize(ctx().stream));
    return;
  }
  void* st = host_stage(bytes);
  TTN_CUDA(cudaMemcpyAsync(st, src, bytes, cudaMemcpyDeviceToHost, ctx().stream));
  TTN_CUDA(cudaStreamSynchronize(ctx().stream));
  memcpy(dst, st, bytes);
}

void* host_stage(size_t bytes) {
  Context& c = ctx();
  if (bytes > c.host_stage_bytes) {
    if (c.host_stage) cudaFreeHost(c.host_stage);
    c.host_stage = nullptr;
    c.host_stage_bytes = 0;
    size_t want = bytes < (size_t)(1 << 16) ? (size_t)(1 << 16) : bytes * 2;
    TTN_CUDA(cudaHostAlloc(&c.host_stage, want, cudaHostAllocPortable));
    c.host_stage_bytes = want;
  }
  return c.host_stage;
}

void DevBuf::alloc(size_t b) {
  release();
  bytes = b;
  if (b == 0) { p = nullptr; return; }
  if (b >= BlockCache::kMinBytes) {
    p = cache().take(b);
    if (p) return;
  }
  cudaError_t e = cudaMallocAsync(&p, b, ctx().stream);
  if (e == cudaErrorMemoryAllocation) {      // give the cached blocks back and retry once
    cudaGetLastError();
    cache().trim();
    e = cudaMallocAsync(&p, b, ctx().stream);
  }
  TTN_CUDA(e);
}
void DevBuf::release() {
  if (p) {
    if (bytes >= BlockCache::kMinBytes && ctx().inited) cache().give(bytes, p);
    else cudaFreeAsync(p, ctx().stream);
  }
  p = nullptr;
  bytes = 0;
}

namespace {

inline int grid_for(int64_t n, int threads) {
  int64_t blocks = (n + threads - 1) / threads;
  int64_t cap = (int64_t)ctx().sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <class T>
__global__ void copy4_kernel(const T* __restrict__ src, T* __restrict__ dst, Copy4 c, int64_t total) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx;
    const int64_t i0 = r % c.n0; r /= c.n0;
    const int64_t i1 = r % c.n1; r /= c.n1;
    const int64_t i2 = r % c.n2; r /= c.n2;
    const int64_t i3 = r;
    T v;
    const bool keep = (c.tri == 0) || (c.tri == 1 && i0 <= i1) || (c.tri == 2 && i0 >= i1);
    if (keep) {
      v = src[i0 * c.s0 + i1 * c.s1 + i2 * c.s2 + i3 * c.s3];
      if (c.conj) v = t_conj(v);
      if (c.alpha != 1.0) v = t_scale(v, c.alpha);
    } else {
      v = t_zero<T>();
    }
    dst[i0 * c.d0 + i1 * c.d1 + i2 * c.d2 + i3 * c.d3] = v;
  }
}

template <class T>
__global__ void fill_kernel(T* dst, int64_t n, T v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = v;
}

template <class T>
__global__ void identity_kernel(T* dst, int64_t m, int64_t n, int64_t ld) {
  const int64_t total = m * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx % m, j = idx / m;
    dst[i + j * ld] = (i == j) ? t_one<T>() : t_zero<T>();
  }
}

__global__ void r2c_kernel(const double* src, zc* dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = make_cuDoubleComplex(src[i], 0.0);
}
__global__ void c2r_kernel(const zc* src, double* dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i].x;
}

__device__ __forceinline__ double scale_fun(double v, int mode) {
  const double tiny = 1e-300;
  switch (mode) {
    case 0: return v;
    case 1: return sqrt(v > 0.0 ? v : 0.0);
    case 2: return v > tiny ? 1.0 / sqrt(v) : 0.0;
    default: return v > tiny ? 1.0 / v : 0.0;
  }
}

template <class T>
__global__ void diag_scale_kernel(T* A, int64_t m, int64_t n, int64_t rs, int64_t cs, const double* vec, int mode, int axis,
                                  int64_t bA, int64_t bvec) {
  const int64_t total = m * n;
  T* Ab = A + blockIdx.y * bA;
  const double* vb = vec + blockIdx.y * bvec;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    // iterate along the memory-fast index of A
    int64_t i, j;
    if (rs <= cs) { i = idx % m; j = idx / m; } else { j = idx % n; i = idx / n; }
    const double s = scale_fun(vb[axis == 0 ? i : j], mode);
    T* p = Ab + i * rs + j * cs;
    *p = t_scale(*p, s);
  }
}

template <class T>
__global__ void axpy_kernel(int64_t n, T a, const T* __restrict__ x, T* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T v = y[i];
    t_fma(v, a, x[i]);
    y[i] = v;
  }
}
template <class T>
__global__ void scal_kernel(int64_t n, T a, T* x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = t_mul(a, x[i]);
}

// ---- reductions: stage 1 writes one partial per (block, vector); stage 2 sums partials in fixed order ----
constexpr int RED_T = 256;
constexpr int MAXV = 32;

template <class T>
__device__ __forceinline__ T warp_sum(T v);
template <>
__device__ __forceinline__ double warp_sum<double>(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <>
__device__ __forceinline__ zc warp_sum<zc>(zc v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}

template <class T>
__global__ void multi_dot_stage1(int64_t n, int nv, const T* __restrict__ X, int64_t ldx, const T* __restrict__ y,
                                 T* __restrict__ partial /* [nv][gridDim.x] */) {
  __shared__ T sh[RED_T / 32];
  for (int j = 0; j < nv; ++j) {
    T acc = t_zero<T>();
    const T* xj = X + (int64_t)j * ldx;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      t_fma(acc, t_conj(xj[i]), y[i]);
    acc = warp_sum<T>(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      T s = sh[0];
      for (int w = 1; w < RED_T / 32; ++w) s = t_add(s, sh[w]);
      partial[(int64_t)j * gridDim.x + blockIdx.x] = s;
    }
    __syncthreads();
  }
}
template <class T>
__global__ void multi_dot_stage2(int nparts, const T* __restrict__ partial, T* __restrict__ out) {
  // one warp per vector; fixed summation order
  const int j = blockIdx.x;
  T acc = t_zero<T>();
  for (int i = threadIdx.x; i < nparts; i += 32) acc = t_add(acc, partial[(int64_t)j * nparts + i]);
  acc = warp_sum<T>(acc);
  if (threadIdx.x == 0) out[j] = acc;
}

template <class T>
struct HVec { T h[MAXV]; };

template <class T>
__global__ void multi_axpy_kernel(int64_t n, int nv, const T* __restrict__ X, int64_t ldx, HVec<T> h, T* __restrict__ y,
                                  double sign) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T acc = t_zero<T>();
    for (int j = 0; j < nv; ++j) t_fma(acc, h.h[j], X[i + (int64_t)j * ldx]);
    y[i] = t_add(y[i], t_scale(acc, sign));
  }
}

}  // namespace

template <class T>
void copy4(const T* src, T* dst, const Copy4& c) {
  const int64_t total = c.n0 * c.n1 * c.n2 * c.n3;
  if (total <= 0) return;
  ProfScope prof_scope_(KF_COPY);
  copy4_kernel<T><<<grid_for(total, 256), 256, 0, ctx().stream>>>(src, dst, c, total);
  TTN_CHECK_LAUNCH();
  ctx().launches++;
}
template <class T>
void fill(T* dst, int64_t n, T v) {
  if (n <= 0) return;
  fill_kernel<T><<<grid_for(n, 256), 256, 0, ctx().stream>>>(dst, n, v);
  TTN_CHECK_LAUNCH();
  ctx().launches++;
}
template <class T>
void set_identity(T* dst, int64_t m, int64_t n, int64_t ld) {
  if (m * n <= 0) return;
  identity_kernel<T><<<grid_for(m * n, 256), 256, 0, ctx().stream>>>(dst, m, n, ld);
  TTN_CHECK_LAUNCH();
  ctx().launches++;
}
void real_to_cplx(const double* src, zc* dst, int64_t n) {
  if (n <= 0) return;
  r2c_kernel<<<grid_for(n, 256), 256, 0, ctx().stream>>>(src, dst, n);
  TTN_CHECK_LAUNCH();
  ctx().launches++;
}
void cplx_to_real(const zc* src, double* dst, int64_t n) {
  if (n <= 0) return;
  c2r_kernel<<<grid_for(n, 256), 256, 0, ctx().stream>>>(src, dst, n);
  TTN_CHECK_LAUNCH();
  ctx().launches++;
}
template <class T>
void diag_scale(T* A, int64_t m, int64_t n, int64_t rs, int64_t cs, const double* vec, int mode, int axis, int batch,
                int64_t bA, int64_t bvec) {
  if (m * n <= 0 || batch <= 0) return;
  dim3 grid(grid_for(m * n, 256), batch);
  diag_scale_kernel<T><<<grid, 256, 0, ctx().stream>>>(A, m, n, rs, cs, vec, mode, axis, bA, bvec);
  TTN_CHECK_LAUNCH();
  ctx().launches++;
}
template <class T>
void axpy(int64_t n, T a, const T* x, T* y) {
  if (n <= 0) return;
  axpy_kernel<T><<<grid_for(n, 256), 256, 0, ctx().stream>>>(n, a, x, y);
  TTN_CHECK_LAUNCH();
  ctx().launches++;
}
template <class T>
void scal(int64_t n, T a, T* x) {
  if (n <= 0) return;
  scal_kernel<T><<<grid_for(n, 256), 256, 0, ctx().stream>>>(n, a, x);
  TTN_CHECK_LAUNCH();
  ctx().launches++;
}

template <class T>
void multi_dot(int64_t n, int nv, const T* X, int64_t ldx, const T* y, T* host_out) {
  if (nv <= 0) return;
  ProfScope prof_scope_(KF_REDUCE);
  int nblocks = (int)std::min<int64_t>(std::max<int64_t>(1, (n + RED_T * 4 - 1) / (RED_T * 4)), (int64_t)ctx().sm_count * 4);
  DevBuf partial(sizeof(T) * (size_t)nv * nblocks), out(sizeof(T) * (size_t)nv);
  multi_dot_stage1<T><<<nblocks, RED_T, 0, ctx().stream>>>(n, nv, X, ldx, y, partial.as<T>());
  TTN_CHECK_LAUNCH();
  multi_dot_stage2<T><<<nv, 32, 0, ctx().stream>>>(nblocks, partial.as<T>(), out.as<T>());
  TTN_CHECK_LAUNCH();
  ctx().launches += 2;
  read_back(host_out, out.p, sizeof(T) * nv);
}

template <class T>
void multi_axpy(int64_t n, int nv, const T* X, int64_t ldx, const T* h_host, T* y, double sign) {
  ProfScope prof_scope_(KF_REDUCE);
  for (int j0 = 0; j0 < nv; j0 += MAXV) {
    const int c = std::min(MAXV, nv - j0);
    HVec<T> h;
    for (int j = 0; j < c; ++j) h.h[j] = h_host[j0 + j];
    multi_axpy_kernel<T><<<grid_for(n, 256), 256, 0, ctx().stream>>>(n, c, X + (int64_t)j0 * ldx, ldx, h, y, sign);
    TTN_CHECK_LAUNCH();
    ctx().launches++;
  }
}

template <class T>
double nrm2(int64_t n, const T* x) {
  T r;
  multi_dot<T>(n, 1, x, n, x, &r);
  double v = t_real(r);
  return std::sqrt(v > 0 ? v : 0.0);
}
template <class T>
T dotc(int64_t n, const T* x, const T* y) {
  T r;
  multi_dot<T>(n, 1, x, n, y, &r);
  return r;
}

#define INST(T)                                                                                                  \
  template void copy4<T>(const T*, T*, const Copy4&);                                                            \
  template void fill<T>(T*, int64_t, T);                                                                         \
  template void set_identity<T>(T*, int64_t, int64_t, int64_t);                                                  \
  template void diag_scale<T>(T*, int64_t, int64_t, int64_t, int64_t, const double*, int, int, int, int64_t, int64_t); \
  template void axpy<T>(int64_t, T, const T*, T*);                                                               \
  template void scal<T>(int64_t, T, T*);                                                                         \
  template void multi_dot<T>(int64_t, int, const T*, int64_t, const T*, T*);                                     \
  template void multi_axpy<T>(int64_t, int, const T*, int64_t, const T*, T*, double);                            \
  template double nrm2<T>(int64_t, const T*);                                                                    \
  template T dotc<T>(int64_t, const T*, const T*);
INST(double)
INST(zc)
#undef INST

}  // namespace ttn
