// Site surgery built on the two-site truncated split (the other users of kernels F3 + F5).
//
//   tt_swap_sites    `_swap_adjacent_sites` (src/qtt_tools.jl:660-694, driver `reorder` :731-774) and `_ttm_swap!`
//                    (src/tt_operations.jl:366-383): contract two neighbouring cores, exchange their physical indices,
//                    re-factorise  Theta' = U (S Vt)  with the rank rule of the caller.
//   tt_merge_diag    `_ttm_contract!` (src/tt_operations.jl:385-397): Pi[s] = A[s] B[s], one site fewer.
//   tt_split_site    the inner step of `to_qtt` (src/qtt_tools.jl:254-310): split one physical index n = coarse * fine
//                    into two sites by an SVD of the reshaped core.
//
// Everything is a strided GEMM, one `split_left` (QR / one-sided Jacobi of svd.cu) and re-layout copies; no new kernel.
#include "tt.h"

namespace ttn {
namespace {

// mode 0: keep sigma_j > tol * sigma_1 (all of them if tol <= 0)     qtt_tools.jl:680-685, :286-289
// mode 1: `_svdtrunc` tail-norm rule with a rank cap                    tt_cross_interpolation.jl:149-166
struct SiteRule {
  int mode; int64_t max_bond; double tol;
  int operator()(const double* s, int k) const {
    if (mode == 1) return rank_tailnorm(s, k, max_bond, tol);
    if (tol <= 0) return k;
    int r = 0;
    for (int j = 0; j < k; ++j) r += s[j] > tol * s[0];
    return std::max(1, r);
  }
};

// core (nf, rn, rr) <- SVt (rn x (nf*rr), ld rn):  core[f, j, r] = SVt[j, f + nf*r]
template <class T>
void svt_to_core(const DevBuf& SVt, int nf, int rn, int rr, DevBuf& core) {
  core.alloc(sizeof(T) * (size_t)nf * rn * rr);
  Copy4 c;
  c.n0 = nf; c.n1 = rn; c.n2 = rr;
  c.s0 = rn; c.s1 = 1; c.s2 = (int64_t)rn * nf;
  c.d0 = 1; c.d1 = nf; c.d2 = (int64_t)nf * rn;
  copy4<T>(SVt.as<T>(), core.as<T>(), c);
}

}  // namespace

template <class T>
void tt_swap_sites(TT<T>& x, int k1, int mode, int64_t max_bond, double tol) {
  ttn_assert(1 <= k1 && k1 < x.d, 2, "k must be in 1:(N-1)");
  ttn_assert(x.batch == 1, 2, "site swaps act on a single train");
  ttn_assert(max_bond >= 1, 2, "max_bond must be >= 1");
  const int k = k1 - 1;
  const int n1 = (int)x.dims[k], n2 = (int)x.dims[k + 1];
  const int rl = (int)x.rks[k], rm = (int)x.rks[k + 1], rr = (int)x.rks[k + 2];
  const int p = n2 * rl, q = n1 * rr;
  DevBuf Theta(sizeof(T) * (size_t)p * q);
  {
    GemmArgs g;  // Theta'[(s2,l),(s1,r)] = sum_m A[s1,l,m] B[s2,m,r]                     (qtt_tools.jl:667-676)
    g.M = rl; g.N = rr; g.K = rm;
    g.A = x.cores[k].p; g.sAm = n1; g.sAk = (int64_t)n1 * rl; g.bA1 = 1; g.bA2 = 0;
    g.B = x.cores[k + 1].p; g.sBk = n2; g.sBn = (int64_t)n2 * rm; g.bB1 = 0; g.bB2 = 1;
    g.C = Theta.p; g.sCm = n2; g.sCn = (int64_t)p * n1; g.bC1 = p; g.bC2 = 1;
    g.batch1 = n1; g.batch2 = n2;
    gemm<T>(g);
  }
  DevBuf U, SVt, newB;
  const int rn = split_left<T>(Theta.as<T>(), p, q, 1, p, false, SiteRule{mode, max_bond, tol}, U, SVt);
  svt_to_core<T>(SVt, n1, rn, rr, newB);                                              // (qtt_tools.jl:690-692)
  TTN_CUDA(cudaStreamSynchronize(ctx().stream));
  x.cores[k] = std::move(U);                                                          // (s2, l, j) is the column-major p x rn
  x.cores[k + 1] = std::move(newB);
  std::swap(x.dims[k], x.dims[k + 1]);
  x.rks[k + 1] = rn;
  x.ot[k] = x.ot[k + 1] = 0;
}

template <class T>
void tt_merge_diag(TT<T>& x, int k1) {
  ttn_assert(1 <= k1 && k1 < x.d, 2, "k must be in 1:(N-1)");
  ttn_assert(x.batch == 1, 2, "site merges act on a single train");
  const int k = k1 - 1;
  ttn_assert(x.dims[k] == x.dims[k + 1], 2, "Incompatible TT dimensions");
  const int n = (int)x.dims[k], rl = (int)x.rks[k], rm = (int)x.rks[k + 1], rr = (int)x.rks[k + 2];
  DevBuf Pi(sizeof(T) * (size_t)n * rl * rr);
  GemmArgs g;  // Pi[s,l,r] = sum_m A[s,l,m] B[s,m,r]                                     (tt_operations.jl:390-393)
  g.M = rl; g.N = rr; g.K = rm;
  g.A = x.cores[k].p; g.sAm = n; g.sAk = (int64_t)n * rl; g.bA1 = 1;
  g.B = x.cores[k + 1].p; g.sBk = n; g.sBn = (int64_t)n * rm; g.bB1 = 1;
  g.C = Pi.p; g.sCm = n; g.sCn = (int64_t)n * rl; g.bC1 = 1;
  g.batch1 = n;
  gemm<T>(g);
  TTN_CUDA(cudaStreamSynchronize(ctx().stream));
  x.cores[k] = std::move(Pi);
  x.cores.erase(x.cores.begin() + k + 1);
  x.dims.erase(x.dims.begin() + k + 1);
  x.rks.erase(x.rks.begin() + k + 1);
  x.ot.erase(x.ot.begin() + k + 1);
  x.ot[k] = 0;
  x.d -= 1;
}

template <class T>
void tt_split_site(TT<T>& x, int k1, int64_t coarse, int mode, int64_t max_bond, double tol) {
  ttn_assert(1 <= k1 && k1 <= x.d, 2, "k must be in 1:N");
  ttn_assert(x.batch == 1, 2, "site splits act on a single train");
  ttn_assert(max_bond >= 1, 2, "max_bond must be >= 1");
  const int k = k1 - 1;
  const int n = (int)x.dims[k], rl = (int)x.rks[k], rr = (int)x.rks[k + 1];
  ttn_assert(coarse >= 1 && n % coarse == 0, 2, "split size must divide the physical dimension");
  const int nc = (int)coarse, nf = n / nc;
  const int p = nc * rl, q = nf * rr;
  // big-endian split s = fine + coarse_index * nf (qtt_tools.jl:274-280):  M[(c,l),(f,r)] = core[f + nf c, l, r]
  DevBuf M(sizeof(T) * (size_t)p * q);
  {
    Copy4 c;
    c.n0 = p; c.n1 = nf; c.n2 = rr;
    c.s0 = nf; c.s1 = 1; c.s2 = (int64_t)n * rl;         // row c + nc*l sits at nf*(c + nc*l)
    c.d0 = 1; c.d1 = p; c.d2 = (int64_t)p * nf;
    copy4<T>(x.core(k), M.as<T>(), c);
  }
  DevBuf U, SVt, right;
  const int rn = split_left<T>(M.as<T>(), p, q, 1, p, false, SiteRule{mode, max_bond, tol}, U, SVt);
  svt_to_core<T>(SVt, nf, rn, rr, right);                                             // (qtt_tools.jl:297)
  TTN_CUDA(cudaStreamSynchronize(ctx().stream));
  x.cores[k] = std::move(U);                                                          // (c, l, j)
  x.cores.insert(x.cores.begin() + k + 1, std::move(right));
  x.dims[k] = nc;
  x.dims.insert(x.dims.begin() + k + 1, nf);
  x.rks.insert(x.rks.begin() + k + 1, rn);
  x.ot[k] = 0;
  x.ot.insert(x.ot.begin() + k + 1, 0);
  x.d += 1;
}

#define INST(T)                                                                  \
  template void tt_swap_sites<T>(TT<T>&, int, int, int64_t, double);             \
  template void tt_merge_diag<T>(TT<T>&, int);                                   \
  template void tt_split_site<T>(TT<T>&, int, int64_t, int, int64_t, double);
INST(double)
INST(zc)
#undef INST

}  // namespace ttn
