// CROWN bounds for ReLU MLPs on the device: the reference's DEFAULT interval method
//   makeIntervalsInfo(..., IntervalsAutoLirpa())             src/Intervals/Intervals.jl:38,44-45
//   intervalsAutoLirpaSliced                                 src/Intervals/intervals_auto_lirpa.jl:44-63
//   exts/auto_lirpa_bridge.py:97-112 -> auto_LiRPA compute_bounds(method="CROWN")  (vendored, float32)
// which costs K Julia->Python round trips through temporary ONNX files.  Algorithm (backward linear
// relaxation, see oracle/nnsdp_oracle.py "CROWN bounds" for the citations into the vendored auto_LiRPA):
// a target (the pre-activation y_t, or the post-activation x_{k+1} of a prefix) is kept as two affine
// functions  lA x + lb <= target <= uA x + ub  of an earlier activation vector and pushed back layer by layer
//   through relu_k :  uA <- uA+ d_u + uA- d_l,  ub += uA+ . b_u ;   lA <- lA+ d_l + lA- d_u,  lb += lA- . b_u
//   through W_k, b_k:  b += A b_k ;  A <- A W_k                          (one FP64 GEMM for the whole batch)
// down to the input box, where  lower = lA c - |lA| r + lb,  upper = uA c + |uA| r + ub.
// FP64 throughout (the reference's values carry float32 precision).  Rows of lA / uA of all queries of a chunk
// are stacked, so the linear step is one GEMM against the shared W_k'.
#include <algorithm>

#include "internal.h"
#include <stdlib.h>

namespace nnsdp {

namespace {

constexpr int CR_THREADS = 128;

__device__ __forceinline__ double cr_block_sum(double v, double* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = CR_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// ReLU relaxation of one layer from its pre-activation bounds (auto_LiRPA operators/activation.py:306-323,387-388)
__global__ void crown_params_kernel(const double* __restrict__ l, const double* __restrict__ u, long long stride,
                                    int n, double* __restrict__ d_u, double* __restrict__ b_u,
                                    double* __restrict__ d_l) {
  const int q = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const long long o = (long long)q * stride + c;
  const double lb_r = fmin(l[o], 0.0);
  double ub_r = fmax(u[o], 0.0);
  ub_r = fmax(ub_r, lb_r + 1e-8);
  const double ud = ub_r / (ub_r - lb_r);
  d_u[o] = ud;
  b_u[o] = -lb_r * ud;
  d_l[o] = ud > 0.5 ? 1.0 : 0.0;
}

// Post-activation bounds from the finished pre-activation bounds.  x_{k+1} = relu(y_k) is bounded by the reference as
// the output of "identity o relu_k o prefix": the backward pass starts from A = I, whose rows leave relu_k as
// d_u[i] e_i (+ bias b_u[i]) for the upper and d_l[i] e_i for the lower bound.  Every later step is positively
// homogeneous per row and d >= 0, so the chain of row i is d[i] times the chain of y_k[i] that already produced
// (prel, preu):  hi(x_{k+1}) = d_u o hi(y_k) + b_u,  lo(x_{k+1}) = d_l o lo(y_k)  -- no second set of chains.
// Then lb = min(lb, ub), ub = max(lb, ub) (intervals_auto_lirpa.jl:37-39).
__global__ void crown_post_kernel(const double* __restrict__ prel, const double* __restrict__ preu,
                                  const double* __restrict__ d_u, const double* __restrict__ b_u,
                                  const double* __restrict__ d_l, long long par_stride, int n, double* __restrict__ out_lo,
                                  double* __restrict__ out_hi, long long out_stride) {
  const int q = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const long long o = (long long)q * par_stride + c;
  double L = __dmul_rn(d_l[o], prel[o]);
  double U = __dadd_rn(__dmul_rn(d_u[o], preu[o]), b_u[o]);
  L = fmin(L, U);
  U = fmax(L, U);
  out_lo[(long long)q * out_stride + c] = L;
  out_hi[(long long)q * out_stride + c] = U;
}

// One chain step for one row of one query: relaxation through relu_k (params of y_k) and the bias part of
// the linear layer k.  src rows may be shared by all queries (q_stride_src = 0: the target's own W rows).
//   rows: [2][Qc][nrows][ld]  (0 = lower, 1 = upper);  bias: [2][Qc][nrows]
__global__ void __launch_bounds__(CR_THREADS)
crown_row_kernel(const double* __restrict__ srcL, const double* __restrict__ srcU, long long src_row_stride,
                 long long src_q_stride, double* __restrict__ dst, long long dst_row_stride, int nrows, int Qc,
                 int n, const double* __restrict__ d_u, const double* __restrict__ b_u,
                 const double* __restrict__ d_l, long long par_stride, const double* __restrict__ bias_k,
                 double* __restrict__ bias) {
  __shared__ double red[CR_THREADS];
  const int r = blockIdx.x, q = blockIdx.y;
  const double* sl = srcL + (long long)q * src_q_stride + (long long)r * src_row_stride;
  const double* su = srcU + (long long)q * src_q_stride + (long long)r * src_row_stride;
  double* dl = dst + ((long long)q * nrows + r) * dst_row_stride;
  double* du = dst + ((long long)(Qc + q) * nrows + r) * dst_row_stride;
  const double* pu = d_u + (long long)q * par_stride;
  const double* pb = b_u + (long long)q * par_stride;
  const double* pl = d_l + (long long)q * par_stride;
  double accl = 0.0, accu = 0.0;
  for (int c = threadIdx.x; c < n; c += CR_THREADS) {
    double al = sl[c], au = su[c];
    if (au > 0.0) {
      accu = fma(au, pb[c], accu);
      au *= pu[c];
    } else {
      au *= pl[c];
    }
    if (al < 0.0) {
      accl = fma(al, pb[c], accl);
      al *= pu[c];
    } else {
      al *= pl[c];
    }
    accl = fma(al, bias_k[c], accl);
    accu = fma(au, bias_k[c], accu);
    dl[c] = al;
    du[c] = au;
  }
  const double tl = cr_block_sum(accl, red), tu = cr_block_sum(accu, red);
  if (threadIdx.x == 0) {
    bias[(long long)q * nrows + r] += tl;
    bias[((long long)Qc + q) * nrows + r] += tu;
  }
}

// One chain step in ONE launch: relaxation through relu_k, the two bias updates and the product with W_k.
//   dst[col] (n_k entries) = W_k' * relax(src[col])          col = (half, q, r): half 0 = lower, 1 = upper
//   bias[col] += sum_c [sign rule] src[col][c] b_u[c] + relax(src[col])[c] b_k[c]
// A SIMT FP64 GEMM (64 x 64 x 16 tiles, 4 x 4 register blocks) whose B operand -- the rows of lA / uA -- is
// relaxed while it is staged into shared memory; the CTAs of the first row-tile also reduce the bias sums of
// their 64 columns (16 threads share a column of the staged tile: half-warp shuffle).  Replaces a row kernel
// plus a GEMM: the chain is launch-bound on deep narrow nets and this halves the launches.
constexpr int SB_M = 64, SB_N = 64, SB_K = 16, SB_THREADS = 256;

__global__ void __launch_bounds__(SB_THREADS)
crown_step_kernel(const double* __restrict__ Wt, int ldT, int M, int Kdim,          // W_k' : M = n_k rows, Kdim = n_{k+1}
                  const double* __restrict__ srcL, const double* __restrict__ srcU, long long src_row_stride,
                  long long src_q_stride, double* __restrict__ dst, long long ld, int nrows, int Rs, int Qc,
                  const double* __restrict__ d_u, const double* __restrict__ b_u, const double* __restrict__ d_l,
                  long long par_stride, const double* __restrict__ bias_k, double* __restrict__ bias) {
  // nrows active rows per (half, query), stored with Rs row slots per (half, query) in dst and bias
  __shared__ double As[SB_K][SB_M];
  __shared__ double Bs[SB_K][SB_N + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * SB_M, n0 = blockIdx.y * SB_N;
  const int N = 2 * Qc * nrows;
  const bool do_bias = (blockIdx.x == 0);
  // the four columns of the staged B tile this thread loads (k = tid % 16, n = tid / 16 + 16 i)
  const int kk = tid & 15;
  const double* colp[4];
  const double *pu[4], *pb[4], *pl[4];
  long long slot[4];  // (half * Qc + q) * Rs + r : row slot of the column in dst / bias
  bool upper[4], valid[4];
  double bsum[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gn = n0 + (tid >> 4) + 16 * i;
    valid[i] = gn < N;
    const int g = valid[i] ? gn : 0;
    const int half = g / (Qc * nrows), rem = g - half * (Qc * nrows);
    const int q = rem / nrows, r = rem - q * nrows;
    upper[i] = half == 1;
    slot[i] = ((long long)half * Qc + q) * Rs + r;
    colp[i] = (half ? srcU : srcL) + (long long)q * src_q_stride + (long long)r * src_row_stride;
    pu[i] = d_u + (long long)q * par_stride;
    pb[i] = b_u + (long long)q * par_stride;
    pl[i] = d_l + (long long)q * par_stride;
  }
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int k0 = 0; k0 < Kdim; k0 += SB_K) {
#pragma unroll
    for (int i = 0; i < (SB_K * SB_M) / SB_THREADS; ++i) {
      const int idx = tid + i * SB_THREADS;
      const int m = idx % SB_M, k = idx / SB_M;
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < Kdim) ? Wt[gm + (long long)gk * ldT] : 0.0;
    }
    const int gk = k0 + kk;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double t = 0.0;
      if (valid[i] && gk < Kdim) {
        const double v = colp[i][gk];
        const bool use_u = upper[i] ? (v > 0.0) : (v < 0.0);  // the entries that take the upper line of the relaxation
        t = v * (use_u ? pu[i][gk] : pl[i][gk]);
        if (do_bias) {
          if (use_u) bsum[i] = fma(v, pb[i][gk], bsum[i]);
          bsum[i] = fma(t, bias_k[gk], bsum[i]);
        }
      }
      Bs[kk][(tid >> 4) + 16 * i] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SB_K; ++k) {
      double a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][tx * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[k][ty * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int gn = n0 + ty * 4 + j;
    if (gn >= N) continue;
    const int half = gn / (Qc * nrows), rem = gn - half * (Qc * nrows);
    const int q = rem / nrows, r = rem - q * nrows;
    double* out = dst + (((long long)half * Qc + q) * Rs + r) * ld;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gm = m0 + tx * 4 + i;
      if (gm < M) out[gm] = acc[i][j];
    }
  }
  if (do_bias) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double v = bsum[i];
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 16);
      if (kk == 0 && valid[i]) bias[slot[i]] += v;
    }
  }
}

// bias[.][q][r] = b_t[r]  (start of a pre-activation target)
__global__ void crown_init_bias_kernel(const double* __restrict__ bt, int nrows, int Qc, double* __restrict__ bias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * Qc * nrows) return;
  bias[i] = bt ? bt[i % nrows] : 0.0;
}

// start of a post-activation target x_{k+1} = relu(y_k): A = I through relu_k and W_k at once:
//   uA row r = d_u[r] W_k[r, :], ub[r] = b_u[r] + d_u[r] b_k[r];   lA row r = d_l[r] W_k[r, :], lb[r] = d_l[r] b_k[r]
__global__ void __launch_bounds__(CR_THREADS)
crown_init_post_kernel(const double* __restrict__ Wt, int ldT, int n_in_k, const double* __restrict__ bias_k,
                       const double* __restrict__ d_u, const double* __restrict__ b_u,
                       const double* __restrict__ d_l, long long par_stride, double* __restrict__ dst,
                       long long dst_row_stride, int Rs, int row0, int Qc, double* __restrict__ bias) {
  // rows [row0, row0 + gridDim.x) of a stack with Rs row slots per (half, query)
  const int r = blockIdx.x, q = blockIdx.y;
  const double du = d_u[(long long)q * par_stride + r], dl = d_l[(long long)q * par_stride + r];
  const double* w = Wt + (long long)r * ldT;  // row r of W_k
  double* ol = dst + ((long long)q * Rs + row0 + r) * dst_row_stride;
  double* ou = dst + ((long long)(Qc + q) * Rs + row0 + r) * dst_row_stride;
  for (int c = threadIdx.x; c < n_in_k; c += CR_THREADS) {
    ol[c] = dl * w[c];
    ou[c] = du * w[c];
  }
  if (threadIdx.x == 0) {
    bias[(long long)q * Rs + row0 + r] = dl * bias_k[r];
    bias[((long long)Qc + q) * Rs + row0 + r] = b_u[(long long)q * par_stride + r] + du * bias_k[r];
  }
}

// concretisation on the input box; optional min/max post-processing of intervals_auto_lirpa.jl:37-39
__global__ void __launch_bounds__(CR_THREADS)
crown_concretize_kernel(const double* __restrict__ rowsL, const double* __restrict__ rowsU, long long row_stride,
                        long long q_stride, int Rs, int row0, int Qc, int n0, const double* __restrict__ x1min,
                        long long s_min, const double* __restrict__ x1max, long long s_max, int q_first,
                        const double* __restrict__ bias, double* __restrict__ out_lo, double* __restrict__ out_hi,
                        long long out_stride, int postprocess) {
  __shared__ double red[CR_THREADS];
  const int r = blockIdx.x, q = blockIdx.y;
  const double* al = rowsL + (long long)q * q_stride + (long long)r * row_stride;
  const double* au = rowsU + (long long)q * q_stride + (long long)r * row_stride;
  const double* lo = x1min + (long long)(q_first + q) * s_min;
  const double* hi = x1max + (long long)(q_first + q) * s_max;
  double sl = 0.0, su = 0.0;
  for (int c = threadIdx.x; c < n0; c += CR_THREADS) {
    const double cc = 0.5 * (lo[c] + hi[c]), rr = 0.5 * (hi[c] - lo[c]);
    sl += al[c] * cc - fabs(al[c]) * rr;
    su += au[c] * cc + fabs(au[c]) * rr;
  }
  const double tl = cr_block_sum(sl, red), tu = cr_block_sum(su, red);
  if (threadIdx.x == 0) {
    double L = tl + bias[(long long)q * Rs + row0 + r], U = tu + bias[((long long)Qc + q) * Rs + row0 + r];
    if (postprocess) {
      L = fmin(L, U);
      U = fmax(L, U);
    }
    out_lo[(long long)q * out_stride + r] = L;
    out_hi[(long long)q * out_stride + r] = U;
  }
}

// Narrow nets (every width <= ~110): the whole backward chain of one pre-activation target y_t in ONE launch.
// A CTA owns a query: the rows of (lA, uA) live in shared memory (two buffers), each step relaxes them through
// relu_k in place (a warp per row, shuffle-reduced bias sums), multiplies by W_k out of L1/L2 and swaps buffers;
// the last step concretises on the input box.  Replaces t fused step launches per target: on W20-D100 the
// pre-activation targets alone were 4,950 launches.
constexpr int CC_THREADS = 256, CC_MAXW = 128, CC_NE = 4;   // widths up to 128 if the buffers fit (see the launcher)

__device__ __forceinline__ void cc_cp16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}

__global__ void __launch_bounds__(CC_THREADS)
crown_chain_kernel(NetDev net, int t_pre, int post, int maxw, int rcap, const double* __restrict__ d_u,
                   const double* __restrict__ b_u, const double* __restrict__ d_l, long long par_stride,
                   const double* __restrict__ x1min, long long s_min, const double* __restrict__ x1max, long long s_max,
                   int q_first, double* __restrict__ out_lo, double* __restrict__ out_hi, long long out_stride) {
  // The rows of a target never mix, so they are also split over gridDim.z CTAs (a few targets x queries would not
  // fill the GPU, and the steps are latency bound: fewer rows per CTA = shorter steps).
  // pre (post == 0): target y_t, t = t_pre, rows = W_t, results at out[q * out_stride + r]
  // post (post == 1): target x_{t+1} = relu(y_t) seen through the relaxation of relu_t, t = blockIdx.y (all K-1 of
  //   them in one launch: they only need the relaxations, which are complete), rows = d_l / d_u scaled W_t, results
  //   min/max post-processed (intervals_auto_lirpa.jl:37-39) at out[q * out_stride + xoff[t + 1] + r]
  const int t = post ? (int)blockIdx.y : t_pre;
  extern __shared__ __align__(16) double csh[];
  const int ldw = maxw + 1, wp = (maxw + 1) & ~1;
  double* wbuf = csh;                                  // [2][maxw][wp]: W_k of this step and of the next one
  double* cur = wbuf + (size_t)2 * maxw * wp;          // [2][rcap][ldw]: half 0 = lower, 1 = upper; rcap >= rows of this CTA
  double* nxt = cur + (size_t)2 * rcap * ldw;
  double* bias = nxt + (size_t)2 * rcap * ldw;         // [2][rcap]
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_in = net.n_in;
  const int rows_per = (net.n[t + 1] + (int)gridDim.z - 1) / (int)gridDim.z, r0 = blockIdx.z * rows_per;
  const int nrows = min(rows_per, net.n[t + 1] - r0);
  if (nrows <= 0) return;
  // operands of a step, requested one step ahead: W_k by cp.async into shared memory, the relaxation of y_k and
  // b_k into registers (lanes c, c + 32, ...: widths are <= 32 CC_NE), so that no step waits on L2
  auto stage_w = [&](int k, int buf) {
    const double* Wt = net.Wt[k];
    const int ldT = net.ldT[k], cols = net.n[k + 1], chunks = (net.n[k] + 1) / 2;
    double* dst = wbuf + (size_t)buf * maxw * wp;
    for (int i = tid; i < cols * chunks; i += CC_THREADS) {
      const int c = i / chunks, j = i % chunks;
      cc_cp16(dst + c * wp + 2 * j, Wt + (long long)c * ldT + 2 * j);   // Wt is zero padded to ldT rows
    }
  };
  double pu[CC_NE], pb[CC_NE], pl[CC_NE], bk[CC_NE];
  auto load_par = [&](int k, double (&u)[CC_NE], double (&bb)[CC_NE], double (&l)[CC_NE], double (&bs)[CC_NE]) {
    const int n = net.n[k + 1];
    const long long po = (long long)q * par_stride + (net.xoff[k + 1] - n_in);
    const double* bsrc = net.M[k] + (long long)net.n[k] * n;
#pragma unroll
    for (int e = 0; e < CC_NE; ++e) {
      const int c = lane + 32 * e;
      const bool ok = c < n;
      u[e] = ok ? d_u[po + c] : 0.0;
      bb[e] = ok ? b_u[po + c] : 0.0;
      l[e] = ok ? d_l[po + c] : 0.0;
      bs[e] = ok ? bsrc[c] : 0.0;
    }
  };
  if (t >= 1) stage_w(t - 1, 0);
  asm volatile("cp.async.commit_group;\n" ::);
#pragma unroll
  for (int e = 0; e < CC_NE; ++e) pu[e] = pb[e] = pl[e] = bk[e] = 0.0;
  if (t >= 1) load_par(t - 1, pu, pb, pl, bk);
  {  // rows of W_t (row r of W_t = column r of Wt_t), bias b_t; post: scaled by the relaxation of relu_t
    const int n = net.n[t];
    const double* Wt = net.Wt[t];
    const int ldT = net.ldT[t];
    const long long pt = (long long)q * par_stride + (net.xoff[t + 1] - n_in);
    for (int i = tid; i < nrows * n; i += CC_THREADS) {
      const int c = i % n, r = i / n;
      const double w = Wt[c + (long long)(r0 + r) * ldT];
      cur[(0 * rcap + r) * ldw + c] = post ? d_l[pt + r0 + r] * w : w;
      cur[(1 * rcap + r) * ldw + c] = post ? d_u[pt + r0 + r] * w : w;
    }
    const double* bt = net.M[t] + (long long)net.n[t] * net.n[t + 1] + r0;
    for (int i = tid; i < 2 * nrows; i += CC_THREADS) {
      const int h = i / nrows, r = i % nrows;
      double v = bt[r];
      if (post) v = h ? b_u[pt + r0 + r] + d_u[pt + r0 + r] * v : d_l[pt + r0 + r] * v;
      bias[h * rcap + r] = v;
    }
  }
  __syncthreads();
  int wb = 0;
  for (int k = t - 1; k >= 0; --k) {
    const int nk1 = net.n[k + 1], nk = net.n[k];
    double pu2[CC_NE], pb2[CC_NE], pl2[CC_NE], bk2[CC_NE];
#pragma unroll
    for (int e = 0; e < CC_NE; ++e) pu2[e] = pb2[e] = pl2[e] = bk2[e] = 0.0;
    if (k > 0) {
      stage_w(k - 1, wb ^ 1);
      load_par(k - 1, pu2, pb2, pl2, bk2);
    }
    asm volatile("cp.async.commit_group;\n" ::);
    // relaxation through relu_k and the bias sums (same rules as crown_row_kernel)
    for (int hr = warp; hr < 2 * nrows; hr += CC_THREADS / 32) {
      const int h = hr / nrows, r = hr % nrows;
      double* a = cur + (h * rcap + r) * ldw;
      double acc = 0.0;
#pragma unroll
      for (int e = 0; e < CC_NE; ++e) {
        const int c = lane + 32 * e;
        if (c < nk1) {
          double v = a[c];
          const bool steep = h ? (v > 0.0) : (v < 0.0);   // the side that takes the chord (upper relaxation)
          if (steep) {
            acc = fma(v, pb[e], acc);
            v *= pu[e];
          } else {
            v *= pl[e];
          }
          acc = fma(v, bk[e], acc);
          a[c] = v;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) bias[h * rcap + r] += acc;
    }
    asm volatile("cp.async.wait_group 1;\n" ::);       // W_k has landed (the group just committed may be in flight)
    __syncthreads();
    // product with W_k: nxt[h][r][i] = sum_c cur[h][r][c] W_k[c, i]
    const double* W = wbuf + (size_t)wb * maxw * wp;
    for (int idx = tid; idx < 2 * nrows * nk; idx += CC_THREADS) {
      const int i = idx % nk, hr = idx / nk;
      const double* a = cur + ((hr / nrows) * rcap + hr % nrows) * ldw;
      double sum = 0.0;
      for (int c = 0; c < nk1; ++c) sum = fma(a[c], W[c * wp + i], sum);
      nxt[((hr / nrows) * rcap + hr % nrows) * ldw + i] = sum;
    }
    __syncthreads();
    double* tmp = cur;
    cur = nxt;
    nxt = tmp;
    wb ^= 1;
#pragma unroll
    for (int e = 0; e < CC_NE; ++e) pu[e] = pu2[e], pb[e] = pb2[e], pl[e] = pl2[e], bk[e] = bk2[e];
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
  // concretise on the input box (crown_concretize_kernel without post-processing)
  const double* lo = x1min + (long long)(q_first + q) * s_min;
  const double* hi = x1max + (long long)(q_first + q) * s_max;
  for (int r = warp; r < nrows; r += CC_THREADS / 32) {
    const double* al = cur + (0 * rcap + r) * ldw;
    const double* au = cur + (1 * rcap + r) * ldw;
    double sl = 0.0, su = 0.0;
    for (int c = lane; c < n_in; c += 32) {
      const double cc = 0.5 * (lo[c] + hi[c]), rr = 0.5 * (hi[c] - lo[c]);
      sl += al[c] * cc - fabs(al[c]) * rr;
      su += au[c] * cc + fabs(au[c]) * rr;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sl += __shfl_xor_sync(0xffffffffu, sl, o);
      su += __shfl_xor_sync(0xffffffffu, su, o);
    }
    if (lane == 0) {
      double L = sl + bias[0 * rcap + r], U = su + bias[1 * rcap + r];
      long long o = (long long)q * out_stride + r0 + r;
      if (post) {
        L = fmin(L, U);
        U = fmax(L, U);
        o += net.xoff[t + 1];
      }
      out_lo[o] = L;
      out_hi[o] = U;
    }
  }
}

}  // namespace

int launch_crown_params(const double* l, const double* u, long long stride, int n, int Qc, double* d_u,
                        double* b_u, double* d_l, cudaStream_t st) {
  crown_params_kernel<<<dim3((n + 127) / 128, Qc), 128, 0, st>>>(l, u, stride, n, d_u, b_u, d_l);
  return 1;
}

int launch_crown_post(const double* prel, const double* preu, const double* d_u, const double* b_u, const double* d_l,
                      long long par_stride, int n, int Qc, double* out_lo, double* out_hi, long long out_stride,
                      cudaStream_t st) {
  if (n <= 0 || Qc <= 0) return 0;
  crown_post_kernel<<<dim3((n + 127) / 128, Qc), 128, 0, st>>>(prel, preu, d_u, b_u, d_l, par_stride, n, out_lo, out_hi,
                                                              out_stride);
  return 1;
}

int launch_crown_row(const double* srcL, const double* srcU, long long src_row_stride, long long src_q_stride,
                     double* dst, long long dst_row_stride, int nrows, int Qc, int n, const double* d_u,
                     const double* b_u, const double* d_l, long long par_stride, const double* bias_k,
                     double* bias, cudaStream_t st) {
  crown_row_kernel<<<dim3(nrows, Qc), CR_THREADS, 0, st>>>(srcL, srcU, src_row_stride, src_q_stride, dst,
                                                          dst_row_stride, nrows, Qc, n, d_u, b_u, d_l, par_stride,
                                                          bias_k, bias);
  return 1;
}

int launch_crown_step(const double* Wt, int ldT, int M, int Kdim, const double* srcL, const double* srcU,
                      long long src_row_stride, long long src_q_stride, double* dst, long long ld, int nrows, int Rs,
                      int Qc, const double* d_u, const double* b_u, const double* d_l, long long par_stride,
                      const double* bias_k, double* bias, cudaStream_t st) {
  const int N = 2 * Qc * nrows;
  dim3 grid((M + SB_M - 1) / SB_M, (N + SB_N - 1) / SB_N);
  crown_step_kernel<<<grid, SB_THREADS, 0, st>>>(Wt, ldT, M, Kdim, srcL, srcU, src_row_stride, src_q_stride, dst, ld,
                                                 nrows, Rs, Qc, d_u, b_u, d_l, par_stride, bias_k, bias);
  return 1;
}

int launch_crown_init_bias(const double* bt, int nrows, int Qc, double* bias, cudaStream_t st) {
  const int n = 2 * Qc * nrows;
  crown_init_bias_kernel<<<(n + 255) / 256, 256, 0, st>>>(bt, nrows, Qc, bias);
  return 1;
}

int launch_crown_init_post(const double* Wt, int ldT, int n_in_k, const double* bias_k, const double* d_u,
                           const double* b_u, const double* d_l, long long par_stride, double* dst,
                           long long dst_row_stride, int nrows, int Rs, int row0, int Qc, double* bias,
                           cudaStream_t st) {
  crown_init_post_kernel<<<dim3(nrows, Qc), CR_THREADS, 0, st>>>(Wt, ldT, n_in_k, bias_k, d_u, b_u, d_l, par_stride,
                                                                dst, dst_row_stride, Rs, row0, Qc, bias);
  return 1;
}

int launch_crown_concretize(const double* rowsL, const double* rowsU, long long row_stride, long long q_stride,
                            int nrows, int Rs, int row0, int Qc, int n0, const double* x1min, long long s_min,
                            const double* x1max, long long s_max, int q_first, const double* bias, double* out_lo,
                            double* out_hi, long long out_stride, int postprocess, cudaStream_t st) {
  crown_concretize_kernel<<<dim3(nrows, Qc), CR_THREADS, 0, st>>>(rowsL, rowsU, row_stride, q_stride, Rs, row0, Qc, n0,
                                                                 x1min, s_min, x1max, s_max, q_first, bias, out_lo,
                                                                 out_hi, out_stride, postprocess);
  return 1;
}

// Narrow nets (widths up to ~110: whatever fits the shared memory).  post == 0: pre-activation bounds of y_t for Qc
// queries in one launch.
// post == 1: the K-1 post-activation targets x_{t+1}, t = 0 .. K-2, of Qc queries in one launch (out = base of
// xmin / xmax of the chunk, stride xtot).  0 = not applicable.
int launch_crown_chain(const NetDev& net, int t, int post, int ntargets, int maxw, int Qc, const double* d_u,
                       const double* b_u, const double* d_l, long long par_stride, const double* x1min, long long s_min,
                       const double* x1max, long long s_max, int q_first, double* out_lo, double* out_hi,
                       long long out_stride, cudaStream_t st) {
  static const bool off = [] { const char* e = getenv("NNSDP_NO_CROWN_CHAIN"); return e && atoi(e) != 0; }();
  if (off || maxw > CC_MAXW || (!post && t < 1) || ntargets < 1) return 0;
  // Row groups: enough CTAs to occupy the SMs when targets x queries are few, and few enough rows per CTA for the
  // two row buffers to fit next to the two W buffers (widths near 100 need >= 8 groups)
  int rgroups = 1;
  const int ctas = Qc * (post ? ntargets : 1);
  if (ctas < 128) rgroups = std::min(8, (128 + ctas - 1) / ctas);
  auto bytes = [&](int rg) {
    const size_t rcap = (size_t)(maxw + rg - 1) / rg;
    return ((size_t)2 * maxw * ((maxw + 1) & ~1) + (size_t)4 * rcap * (maxw + 1) + 2 * rcap) * sizeof(double);
  };
  while (bytes(rgroups) > (size_t)220 * 1024 && rgroups < 32) rgroups *= 2;
  if (bytes(rgroups) > (size_t)220 * 1024) return 0;
  const size_t smem = bytes(rgroups);
  const int rcap = (maxw + rgroups - 1) / rgroups;
  if (smem > 48 * 1024) cudaFuncSetAttribute(crown_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  crown_chain_kernel<<<dim3(Qc, post ? ntargets : 1, rgroups), CC_THREADS, smem, st>>>(net, t, post, maxw, rcap, d_u, b_u, d_l,
                                                                                      par_stride, x1min, s_min, x1max, s_max,
                                                                                      q_first, out_lo, out_hi, out_stride);
  return 1;
}

}  // namespace nnsdp
