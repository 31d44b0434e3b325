// TTO x TTV core contraction (kernel family F6).
//
// Replaces the per-site @tensoropt of src/tt_operations.jl:105-108:
//     y_k[i, a, nu, b, mu] = sum_j A_k[i, j, a, b] * x_k[j, nu, mu]
// with the result stored as the (n, R_l*r_l, R_r*r_r) core whose fused bonds have the MPO index fastest
// (tt_operations.jl:106).  The inner dimension is n_in (2 for QTT), so this is pure memory traffic:
// the MPO core lives in shared memory, each x element is read once per CTA pass through L1, and every
// thread writes consecutive output elements (i fastest) so the stores are fully coalesced.  The
// reference's zero-fill of the output (tt_operations.jl:103) is not needed and not done.
#include "ttn_internal.h"

namespace ttn {
namespace {

__device__ __forceinline__ double shfl_elem(double v, unsigned src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ zc shfl_elem(zc v, unsigned src) {
  return make_cuDoubleComplex(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

// QTT fast path (n_out = n_in = 2, power-of-two bonds): one thread per (a, nu, b) produces the two outputs i = 0, 1 (32
// contiguous bytes for ComplexF64), consecutive threads continue at the next address.  Index split by shifts and masks,
// everything unrolled, 32-bit offsets inside one (mu, batch) slab.  grid: x = chunks of (a, nu, b), y = mu chunks, z = batch.
template <class T>
__global__ void __launch_bounds__(256) apply_qtt_kernel(const T* __restrict__ A, const T* __restrict__ x, T* __restrict__ y,
                                                        int Rl, int Rr, int rl, int rr, int sh_a, int sh_nu, int64_t bx,
                                                        int64_t by, int staged) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const T* As = A;                                 // [ (a,b) ][ j ][ i ]: one thread reads 4 consecutive values A[i, j, a, b]
  if (staged) {
    T* S = reinterpret_cast<T*>(smem_raw);
    for (int t = threadIdx.x; t < 4 * Rl * Rr; t += blockDim.x) S[t] = A[t];
    __syncthreads();
    As = S;
  }
  const unsigned inner = (unsigned)Rl * rl * Rr;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned mu = blockIdx.y; mu < (unsigned)rr; mu += gridDim.y) {
    const T* xb = x + blockIdx.z * bx + (int64_t)2 * rl * mu;
    T* yb = y + blockIdx.z * by + (int64_t)2 * inner * mu;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < inner; idx += stride) {
      const unsigned a = idx & (Rl - 1), nu = (idx >> sh_a) & (rl - 1), b = idx >> (sh_a + sh_nu);
      const T x0 = xb[2 * nu], x1 = xb[2 * nu + 1];
      const T* ap = As + 4 * (a + Rl * b);
      const T a00 = ap[0], a10 = ap[1], a01 = ap[2], a11 = ap[3];     // a_ij
      T y0 = t_mul(a00, x0), y1 = t_mul(a10, x0);
      t_fma(y0, a01, x1);
      t_fma(y1, a11, x1);
      // lane L of a full warp holds outputs 2L, 2L+1 of a 64-element run; re-distribute by shuffles so that each of the two
      // store instructions writes one contiguous 32-element half of the run (full sectors) instead of every other 16 bytes
      const unsigned lane = threadIdx.x & 31;
      const unsigned wbase = idx - lane;                         // warp-uniform
      if (wbase + 31 < inner) {
        const unsigned src0 = lane >> 1, src1 = 16 + (lane >> 1);
        const T lo0 = shfl_elem(y0, src0), hi0 = shfl_elem(y1, src0);
        const T lo1 = shfl_elem(y0, src1), hi1 = shfl_elem(y1, src1);
        yb[2 * wbase + lane] = (lane & 1) ? hi0 : lo0;
        yb[2 * wbase + 32 + lane] = (lane & 1) ? hi1 : lo1;
      } else {
        yb[2 * idx] = y0;
        yb[2 * idx + 1] = y1;
      }
    }
  }
}

// generic path: one thread per (a, nu, b) produces the n_out outputs; divisions for the index split
template <class T>
__global__ void __launch_bounds__(256) apply_kernel(const T* __restrict__ A, const T* __restrict__ x, T* __restrict__ y,
                                                    int n_out, int n_in, int Rl, int Rr, int rl, int rr, int64_t bx, int64_t by,
                                                    int staged) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const T* As = A;
  if (staged) {
    T* S = reinterpret_cast<T*>(smem_raw);
    const int na = n_out * n_in * Rl * Rr;
    for (int i = threadIdx.x; i < na; i += blockDim.x) S[i] = A[i];
    __syncthreads();
    As = S;
  }
  const unsigned inner = (unsigned)Rl * rl * Rr;
  for (unsigned mu = blockIdx.y; mu < (unsigned)rr; mu += gridDim.y) {
    const T* xb = x + blockIdx.z * bx + (int64_t)n_in * rl * mu;
    T* yb = y + blockIdx.z * by + (int64_t)inner * n_out * mu;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < inner; idx += gridDim.x * blockDim.x) {
      const unsigned t1 = idx / Rl, a = idx - t1 * Rl, b = t1 / rl, nu = t1 - b * rl;
      const T* xp = xb + (size_t)n_in * nu;
      const T* ap = As + (size_t)n_out * n_in * (a + Rl * b);
      T* yp = yb + (size_t)idx * n_out;
      for (int i = 0; i < n_out; ++i) {
        T acc = t_zero<T>();
        for (int j = 0; j < n_in; ++j) t_fma(acc, ap[i + n_out * j], xp[j]);
        yp[i] = acc;
      }
    }
  }
}

}  // namespace

template <class T>
void apply_core(const T* A, const T* x, T* y, int n_out, int n_in, int Rl, int Rr, int rl, int rr, int batch, int64_t bx,
                int64_t by) {
  const int64_t inner = (int64_t)Rl * rl * Rr;
  if (inner <= 0 || n_out <= 0 || rr <= 0 || batch <= 0) return;
  ttn_assert(inner * n_out < (1LL << 31), 2, "apply: core too large for 32-bit sub-index arithmetic");
  ProfScope prof_scope_(KF_APPLY);
  // the MPO core is staged in shared memory when it fits (up to 200 KB: e.g. the rank-51 ComplexF64 QFT operator of
  // examples/dft.jl needs 166 KB); larger cores are read through L1 / L2 from global memory
  size_t smem = sizeof(T) * (size_t)n_out * n_in * Rl * Rr;
  const int staged = smem <= 200 * 1024;
  if (!staged) smem = 0;
  const bool qtt = n_out == 2 && n_in == 2 && (Rl & (Rl - 1)) == 0 && (rl & (rl - 1)) == 0;
  int sh_a = 0, sh_nu = 0;
  while ((1 << sh_a) < Rl) ++sh_a;
  while ((1 << sh_nu) < rl) ++sh_nu;
  const unsigned bxg = (unsigned)std::min<int64_t>((inner + 255) / 256, 64);
  ttn_assert(rr <= 65535, 2, "apply: right bond too large for grid.y");
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = std::min(65535, batch - b0);
    // enough CTAs for ~16 per SM, each looping over several mu so that the MPO-core staging is amortised
    const int64_t want = (int64_t)ctx().sm_count * 16;
    const unsigned gy = (unsigned)std::max<int64_t>(1, std::min<int64_t>(rr, want / std::max<int64_t>(1, (int64_t)bxg * nb)));
    dim3 grid(bxg, gy, (unsigned)nb);
    if (qtt) {
      if (smem > 48 * 1024)
        TTN_CUDA(cudaFuncSetAttribute(apply_qtt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      apply_qtt_kernel<T><<<grid, 256, smem, ctx().stream>>>(A, x + (int64_t)b0 * bx, y + (int64_t)b0 * by, Rl, Rr, rl, rr, sh_a,
                                                             sh_nu, bx, by, staged);
    } else {
      if (smem > 48 * 1024)
        TTN_CUDA(cudaFuncSetAttribute(apply_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      apply_kernel<T><<<grid, 256, smem, ctx().stream>>>(A, x + (int64_t)b0 * bx, y + (int64_t)b0 * by, n_out, n_in, Rl, Rr, rl,
                                                         rr, bx, by, staged);
    }
    TTN_CHECK_LAUNCH();
    ctx().launches++;
  }
}

template void apply_core<double>(const double*, const double*, double*, int, int, int, int, int, int, int, int64_t, int64_t);
template void apply_core<zc>(const zc*, const zc*, zc*, int, int, int, int, int, int, int, int64_t, int64_t);

}  // namespace ttn
