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

template <class T>
__global__ void apply_kernel(const T* __restrict__ A, const T* __restrict__ x, T* __restrict__ y, int n_out, int n_in,
                             int Rl, int Rr, int rl, int rr, int64_t bx, int64_t by) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* As = reinterpret_cast<T*>(smem_raw);
  const int na = n_out * n_in * Rl * Rr;
  for (int i = threadIdx.x; i < na; i += blockDim.x) As[i] = A[i];
  __syncthreads();
  const T* xb = x + blockIdx.y * bx;
  T* yb = y + blockIdx.y * by;
  const int64_t total = (int64_t)n_out * Rl * rl * Rr * rr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx;
    const int i = (int)(r % n_out); r /= n_out;
    const int a = (int)(r % Rl); r /= Rl;
    const int nu = (int)(r % rl); r /= rl;
    const int b = (int)(r % Rr); r /= Rr;
    const int mu = (int)r;
    const T* xp = xb + (int64_t)n_in * (nu + (int64_t)rl * mu);
    const T* ap = As + i + (int64_t)n_out * n_in * (a + Rl * b);
    T acc = t_zero<T>();
    for (int j = 0; j < n_in; ++j) t_fma(acc, ap[(int64_t)n_out * j], xp[j]);
    yb[idx] = acc;
  }
}

}  // namespace

template <class T>
void apply_core(const T* A, const T* x, T* y, int n_out, int n_in, int Rl, int Rr, int rl, int rr, int batch, int64_t bx,
                int64_t by) {
  const int64_t total = (int64_t)n_out * Rl * rl * Rr * rr;
  if (total <= 0 || batch <= 0) return;
  ProfScope prof_scope_(KF_APPLY);
  const size_t smem = sizeof(T) * (size_t)n_out * n_in * Rl * Rr;
  ttn_assert(smem <= 48 * 1024, 2, "apply: MPO core does not fit in shared memory");
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = std::max<int64_t>(1, (int64_t)ctx().sm_count * 8 / std::max(1, std::min(batch, 8)));
  if (blocks > cap) blocks = cap;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = std::min(65535, batch - b0);
    dim3 grid((unsigned)blocks, (unsigned)nb);
    apply_kernel<T><<<grid, 256, smem, ctx().stream>>>(A, x + (int64_t)b0 * bx, y + (int64_t)b0 * by, n_out, n_in, Rl, Rr, rl,
                                                       rr, bx, by);
    TTN_CHECK_LAUNCH();
    ctx().launches++;
  }
}

template void apply_core<double>(const double*, const double*, double*, int, int, int, int, int, int, int, int64_t, int64_t);
template void apply_core<zc>(const zc*, const zc*, zc*, int, int, int, int, int, int, int, int64_t, int64_t);

}  // namespace ttn
