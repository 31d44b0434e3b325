// FP64 tensor-pipe building blocks shared by the DMMA kernels (gemm.cu, jacobi_gram.cu).
// mma.sync.aligned.m8n8k4.f64 (SASS: DMMA.8x8x4): lane = 4*g8 + t4 holds A[g8][t4], B[t4][g8] and C[g8][2*t4 + {0,1}].
#pragma once
#include "ttn_internal.h"

namespace ttn {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <class T> struct Acc;
template <> struct Acc<double> {
  double c[2];
  __device__ __forceinline__ void zero() { c[0] = c[1] = 0.0; }
  __device__ __forceinline__ void mma(double a, double b) { dmma884(c[0], c[1], a, b); }
  __device__ __forceinline__ double get(int i) const { return c[i]; }
};
template <> struct Acc<zc> {
  double re[2], im[2];
  __device__ __forceinline__ void zero() { re[0] = re[1] = im[0] = im[1] = 0.0; }
  __device__ __forceinline__ void mma(zc a, zc b) {
    dmma884(re[0], re[1], a.x, b.x);
    dmma884(re[0], re[1], -a.y, b.y);
    dmma884(im[0], im[1], a.x, b.y);
    dmma884(im[0], im[1], a.y, b.x);
  }
  __device__ __forceinline__ zc get(int i) const { return make_cuDoubleComplex(re[i], im[i]); }
};


// 8 / 16-byte asynchronous global -> shared copy (LDGSTS); src_bytes = 0 zero-fills the destination (out-of-range elements)
template <class T>
__device__ __forceinline__ void cp_async_elem(T* smem_dst, const T* gsrc, bool valid) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? (int)sizeof(T) : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(dst), "l"(gsrc), "n"(sizeof(T)), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }


}  // namespace ttn
