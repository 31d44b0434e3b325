// One-sided (Hestenes) Jacobi SVD (kernel family F5).
//
// Replaces the LAPACK gesdd calls behind the reference's truncations: `_svdtrunc`
// (src/tt_cross_interpolation.jl:150, used by tt_compress! src/tt_tools.jl:752 and tdvp2
// src/solvers/tdvp.jl:250,278), the MALS core moves (src/solvers/mals.jl:99,126) and the DMRG core moves
// (src/solvers/dmrg.jl:189,213).  The bond-sized matrix is first reduced to a square triangular factor by
// the Householder QR of qr.cu (svd.cu), then its columns are orthogonalised in place by plane rotations:
//       X <- X V,   X^H X diagonal,   sigma_j = ||x_j||,   u_j = x_j / sigma_j.
// One-sided Jacobi delivers small singular values with high *relative* accuracy, which the reference's
// tolerance-driven rank rules (mals.jl:42-56 `sv_trunc`, dmrg.jl:179-185 `cut_off_index`) depend on.
//
// Kernel: a CTA stages a group of columns in shared memory (conflict-free: lanes walk down a column),
// one warp per column pair and round (round-robin tournament, all pairs of a round are disjoint), three
// dot products by warp shuffles, rotation applied in place.  If all n columns fit in one SM's shared
// memory the CTA iterates sweeps to convergence without leaving the kernel; otherwise the host walks a
// block round-robin (one launch per step, CTAs = independent block pairs, data stays in L2).
#include "ttn_internal.h"

namespace ttn {
namespace {

constexpr int JAC_MAX_SWEEPS = 40;

__device__ __forceinline__ double wsumd(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Rotates the column pair (xp, xq) of length m (shared memory) with a group of GL lanes (GL = 16: two pairs per
// warp in flight).  Returns true when a rotation was applied, i.e. |x_p^H x_q| > tol·||x_p||·||x_q||.
template <class T, int GL>
__device__ __forceinline__ bool rotate_pair(T* xp, T* xq, int m, int gl, double tol2, bool active) {
  // every lane of the warp must reach the shuffles below: inactive groups contribute zeros and never rotate
  double a = 0.0, b = 0.0;
  T c = t_zero<T>();
  if (active)
  for (int i = gl; i < m; i += GL) {
    const T p = xp[i], q = xq[i];
    a += t_abs2(p);
    b += t_abs2(q);
    t_fma(c, t_conj(p), q);
  }
  double cr = t_real(c), ci = t_imag(c);
#pragma unroll
  for (int o = GL / 2; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    cr += __shfl_xor_sync(0xffffffffu, cr, o);
    if (is_cplx<T>::value) ci += __shfl_xor_sync(0xffffffffu, ci, o);
  }
  const double c2 = cr * cr + ci * ci;
  if (!(c2 > tol2 * a * b)) return false;   // also covers zero columns and NaNs
  double cs, sn, phr = 1.0, phi = 0.0;
  if (is_cplx<T>::value) {
    const double inv = rsqrt(c2);           // 1/|c|
    phr = cr * inv; phi = ci * inv;
    const double zeta = 0.5 * (b - a) * inv;
    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    cs = rsqrt(1.0 + t * t);
    sn = cs * t;
  } else {
    const double zeta = 0.5 * (b - a) / cr;  // real case: the sign of c is carried by zeta
    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    cs = rsqrt(1.0 + t * t);
    sn = cs * t;
  }
  const bool swap = a < b;  // keep the larger column first (de Rijk ordering)
  const T ph = t_from<T>(phr, phi);
  const T phc = t_from<T>(phr, -phi);
  for (int i = gl; i < m; i += GL) {
    const T p = xp[i], q = xq[i];
    const T np = t_sub(t_scale(p, cs), t_scale(t_mul(phc, q), sn));
    const T nq = t_add(t_scale(t_mul(ph, p), sn), t_scale(q, cs));
    xp[i] = swap ? nq : np;
    xq[i] = swap ? np : nq;
  }
  return true;
}

constexpr int JGL = 16;  // lanes per column pair

// mode 0: all pairs among the na+nb columns; mode 1: cross pairs (one column from each group) only.
// full != 0: iterate sweeps until a whole sweep applies no rotation (single-CTA problem), else run `sweeps` sweeps.
template <class T>
__global__ void __launch_bounds__(1024) jacobi_kernel(T* __restrict__ X, int m, int64_t ldx, int64_t bX,
                                                      const int* __restrict__ grpA, const int* __restrict__ grpB, int bsz,
                                                      int n, int mode, int full, int sweeps, double tol,
                                                      unsigned int* __restrict__ d_rotated, int* __restrict__ d_sweeps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Xs = reinterpret_cast<T*>(smem_raw);
  __shared__ int s_rot;
  T* Xb = X + blockIdx.y * bX;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int grp = tid / JGL, ngrp = blockDim.x / JGL, gl = tid % JGL;
  const double tol2 = tol * tol;

  const int a0 = grpA[blockIdx.x] * bsz;
  const int na = min(bsz, n - a0);
  const int gb = grpB[blockIdx.x];
  const int b0 = gb >= 0 ? gb * bsz : 0;
  const int nb = gb >= 0 ? min(bsz, n - b0) : 0;
  const int nc = na + nb;

  for (int c = warp; c < nc; c += nwarps) {
    const T* src = Xb + (int64_t)(c < na ? a0 + c : b0 + (c - na)) * ldx;
    T* dst = Xs + (size_t)c * m;
    for (int i = lane; i < m; i += 32) dst[i] = src[i];
  }
  if (tid == 0) s_rot = 0;
  __syncthreads();

  int sw = 0;
  const int max_sw = full ? JAC_MAX_SWEEPS : sweeps;
  bool any_rot = false;
  for (; sw < max_sw; ++sw) {
    bool myrot = false;
    if (mode == 0) {
      const int ne = nc + (nc & 1);          // even number of players (last may be a dummy)
      const int half = ne / 2;
      for (int r = 0; r < ne - 1; ++r) {
        for (int i0 = 0; i0 < half; i0 += ngrp) {   // warp-uniform trip count
          const int i = i0 + grp;
          int p = 0, q = 0;
          if (i == 0) { p = ne - 1; q = r; }
          else if (i < half) { p = (r + i) % (ne - 1); q = (r - i + (ne - 1)) % (ne - 1); }
          const bool act = i < half && p < nc && q < nc;
          if (p > q) { const int t = p; p = q; q = t; }
          myrot |= rotate_pair<T, JGL>(Xs + (size_t)(act ? p : 0) * m, Xs + (size_t)(act ? q : 0) * m, m, gl, tol2, act);
        }
        __syncthreads();
      }
    } else {
      const int bm = max(na, nb);
      for (int r = 0; r < bm; ++r) {
        for (int i0 = 0; i0 < bm; i0 += ngrp) {     // warp-uniform trip count
          const int p = i0 + grp, q = (p + r) % bm;
          const bool act = p < na && q < nb;
          myrot |= rotate_pair<T, JGL>(Xs + (size_t)(act ? p : 0) * m, Xs + (size_t)(act ? na + q : 0) * m, m, gl, tol2, act);
        }
        __syncthreads();
      }
    }
    if (myrot) s_rot = 1;       // benign race: every writer stores the same value
    __syncthreads();
    const int rot = s_rot;
    __syncthreads();
    if (tid == 0) s_rot = 0;
    any_rot |= (rot != 0);
    if (full && !rot) { ++sw; break; }
  }

  __syncthreads();
  for (int c = warp; c < nc; c += nwarps) {
    T* dst = Xb + (int64_t)(c < na ? a0 + c : b0 + (c - na)) * ldx;
    const T* src = Xs + (size_t)c * m;
    for (int i = lane; i < m; i += 32) dst[i] = src[i];
  }
  if (tid == 0) {
    if (d_rotated && any_rot) atomicOr(d_rotated, 1u);
    if (d_sweeps && full) d_sweeps[blockIdx.y] = sw;
  }
}

template <class T>
__global__ void colnorm_kernel(const T* __restrict__ X, int m, int n, int64_t ldx, int64_t bX, double* __restrict__ norms,
                               int64_t bnorms) {
  // one warp per column
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const T* x = X + blockIdx.y * bX + (int64_t)warp * ldx;
  double a = 0.0;
  for (int i = lane; i < m; i += 32) a += t_abs2(x[i]);
  a = wsumd(a);
  if (lane == 0) norms[blockIdx.y * bnorms + warp] = sqrt(a);
}

template <class T>
__global__ void gather_kernel(const T* __restrict__ X, int m, int64_t ldx, const int* __restrict__ perm,
                              const double* __restrict__ scale, int r, T* __restrict__ dst, int64_t rs, int64_t cs,
                              int64_t bX, int64_t bperm, int64_t bdst) {
  const int64_t total = (int64_t)m * r;
  const T* Xb = X + blockIdx.y * bX;
  const int* pb = perm + blockIdx.y * bperm;
  const double* sb = scale + blockIdx.y * bperm;
  T* db = dst + blockIdx.y * bdst;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % m), j = (int)(idx / m);
    db[i * rs + j * cs] = t_scale(Xb[i + (int64_t)pb[j] * ldx], sb[j]);
  }
}

}  // namespace

template <class T>
int jacobi_orth(T* X, int m, int n, int64_t ldx, double* norms, int batch, int64_t bX, int64_t bnorms) {
  if (n <= 0 || m <= 0 || batch <= 0) return 0;
  // rotation threshold |x_p^H x_q| <= tol ||x_p|| ||x_q||: m·eps is the rounding level of the computed inner product
  const double tol = (double)std::max(m, 8) * 1.1102230246251565e-16;
  const size_t budget = 220 * 1024;
  auto kern = jacobi_kernel<T>;
  static bool attr_done = false;
  if (!attr_done) {
    TTN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(224 * 1024)));
    attr_done = true;
  }
  int sweeps_used = 0;
  const size_t col_bytes = sizeof(T) * (size_t)m;
  ttn_assert(2 * col_bytes <= budget, 2, "jacobi: a column pair does not fit in shared memory");
  if ((size_t)n * col_bytes <= budget) {
    // whole matrix in one SM: iterate to convergence inside the kernel
    int h_grp[2] = {0, -1};
    DevBuf grp(sizeof(int) * 2), dsw(sizeof(int) * batch);
    TTN_CUDA(cudaMemcpyAsync(grp.p, h_grp, sizeof(h_grp), cudaMemcpyHostToDevice, ctx().stream));
    const int pairs = (n + 1) / 2;
    int threads = std::min(1024, std::max(64, ((JGL * pairs + 31) / 32) * 32));
    for (int b0 = 0; b0 < batch; b0 += 65535) {
      const int nb = std::min(65535, batch - b0);
      dim3 grid(1, nb);
      ProfScope prof_scope_(KF_JACOBI);
      kern<<<grid, threads, (size_t)n * col_bytes, ctx().stream>>>(X + (int64_t)b0 * bX, m, ldx, bX, grp.as<int>(),
                                                                 grp.as<int>() + 1, n, n, 0, 1, 0, tol, nullptr,
                                                                 dsw.as<int>() + b0);
      TTN_CHECK_LAUNCH();
      ctx().launches++;
    }
    int hsw = 0;   // sweeps of batch element 0 (diagnostics; one 4-byte D2H, overlapped with the norms read-back sync)
    TTN_CUDA(cudaMemcpyAsync(&hsw, dsw.p, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
    TTN_CUDA(cudaStreamSynchronize(ctx().stream));
    sweeps_used = hsw;
  } else {
    // block Jacobi: blocks of bsz columns, two blocks per CTA
    int bsz = (int)(budget / (2 * col_bytes));
    if (bsz > 32) bsz = 32;
    const int nblk = (n + bsz - 1) / bsz;
    const int ne = nblk + (nblk & 1);
    DevBuf dmax(sizeof(unsigned int));
    // pair lists of every round-robin step, uploaded once: step 0 = pairs inside each block,
    // steps 1..ne-1 = block tournament (cross pairs only)
    std::vector<int> hA, hB, off(ne + 1, 0);
    for (int i = 0; i < nblk; ++i) { hA.push_back(i); hB.push_back(-1); }
    off[1] = nblk;
    for (int r = 0; r < ne - 1; ++r) {
      for (int i = 0; i < ne / 2; ++i) {
        int p, q;
        if (i == 0) { p = ne - 1; q = r; }
        else { p = (r + i) % (ne - 1); q = (r - i + (ne - 1)) % (ne - 1); }
        if (p < nblk && q < nblk) { hA.push_back(std::min(p, q)); hB.push_back(std::max(p, q)); }
      }
      off[r + 2] = (int)hA.size();
    }
    DevBuf gA(sizeof(int) * hA.size()), gB(sizeof(int) * hB.size());
    TTN_CUDA(cudaMemcpyAsync(gA.p, hA.data(), sizeof(int) * hA.size(), cudaMemcpyHostToDevice, ctx().stream));
    TTN_CUDA(cudaMemcpyAsync(gB.p, hB.data(), sizeof(int) * hB.size(), cudaMemcpyHostToDevice, ctx().stream));
    TTN_CUDA(cudaStreamSynchronize(ctx().stream));
    ttn_assert(batch <= 65535, 2, "jacobi: batch too large for the block path");
    const int threads = std::min(1024, std::max(64, ((JGL * bsz + 31) / 32) * 32));
    for (int sw = 0; sw < JAC_MAX_SWEEPS; ++sw) {
      TTN_CUDA(cudaMemsetAsync(dmax.p, 0, sizeof(unsigned int), ctx().stream));
      for (int st = 0; st < ne; ++st) {
        const int cnt = off[st + 1] - off[st];
        if (cnt <= 0) continue;
        dim3 grid(cnt, batch);
        const size_t smem = (size_t)(st == 0 ? 1 : 2) * bsz * col_bytes;
        ProfScope prof_scope_(KF_JACOBI);
        kern<<<grid, threads, smem, ctx().stream>>>(X, m, ldx, bX, gA.as<int>() + off[st], gB.as<int>() + off[st], bsz, n,
                                                   st == 0 ? 0 : 1, 0, 1, tol, dmax.as<unsigned int>(), nullptr);
        TTN_CHECK_LAUNCH();
        ctx().launches++;
      }
      unsigned int rotated = 0;
      TTN_CUDA(cudaMemcpyAsync(&rotated, dmax.p, sizeof(rotated), cudaMemcpyDeviceToHost, ctx().stream));
      TTN_CUDA(cudaStreamSynchronize(ctx().stream));
      sweeps_used = sw + 1;
      if (!rotated) break;
    }
  }
  {
    const int wpb = 8;
    for (int b0 = 0; b0 < batch; b0 += 65535) {
      const int nb = std::min(65535, batch - b0);
      dim3 grid((n + wpb - 1) / wpb, nb);
      ProfScope prof_scope_(KF_GATHER);
      colnorm_kernel<T><<<grid, wpb * 32, 0, ctx().stream>>>(X + (int64_t)b0 * bX, m, n, ldx, bX, norms + (int64_t)b0 * bnorms,
                                                            bnorms);
      TTN_CHECK_LAUNCH();
      ctx().launches++;
    }
  }
  ctx().last_jacobi_sweeps = sweeps_used;
  return sweeps_used;
}

template <class T>
void gather_cols(const T* X, int m, int64_t ldx, const int* perm, const double* scale, int r, T* dst, int64_t rs, int64_t cs,
                 int batch, int64_t bX, int64_t bperm, int64_t bdst) {
  if (r <= 0 || m <= 0 || batch <= 0) return;
  const int64_t total = (int64_t)m * r;
  int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx().sm_count * 4);
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = std::min(65535, batch - b0);
    dim3 grid(blocks, nb);
    ProfScope prof_scope_(KF_GATHER);
    gather_kernel<T><<<grid, 256, 0, ctx().stream>>>(X + (int64_t)b0 * bX, m, ldx, perm + (int64_t)b0 * bperm,
                                                    scale + (int64_t)b0 * bperm, r, dst + (int64_t)b0 * bdst, rs, cs, bX, bperm,
                                                    bdst);
    TTN_CHECK_LAUNCH();
    ctx().launches++;
  }
}

#define INST(T)                                                                                   \
  template int jacobi_orth<T>(T*, int, int, int64_t, double*, int, int64_t, int64_t);              \
  template void gather_cols<T>(const T*, int, int64_t, const int*, const double*, int, T*, int64_t, int64_t, int, int64_t, \
                               int64_t, int64_t);
INST(double)
INST(zc)
#undef INST

}  // namespace ttn
