// K1 (interval bound propagation), K2 (sector slopes) and the batched affine-column
// contraction.  FP64 throughout.
//
// Reference semantics:
//   intervalsWorstCase            src/Intervals/intervals_easy.jl:21-33
//   one-step pre-activation IBP   src/Intervals/intervals_auto_lirpa.jl:55-62
//   makeSectorMinMax (ReLU)       src/Qc/activ_sector.jl:63-72
//
// The IBP of one layer for Q boxes is a GEMM-shaped contraction.  The reference computes
//   ymin = W+ xmin + W- xmax + b,   ymax = W+ xmax + W- xmin + b        (intervals_easy.jl:23-24)
// with W+ = max(W,0), W- = min(W,0) re-materialised on every call; the same interval in centre / radius
// form is  y = W c + b -/+ |W| r,  c = (xmin + xmax)/2, r = (xmax - xmin)/2 : two products instead of four
// (|W| formed in registers), M = n_{k+1}, N = Q, K = n_k.  A degenerate box (r = 0) gives ymin == ymax.
#include "internal.h"
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace nnsdp {

namespace {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int GEMM_THREADS = 256;

// MODE 0: IBP layer.  MODE 1: C += A * B (affine column).  MODE 2: C = A * B.
template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_nn_kernel(const double* __restrict__ A, int lda, int M, int Kdim,
               const double* __restrict__ B0, const double* __restrict__ B1, long long ldb, int N,
               const double* __restrict__ bias,
               double* __restrict__ C0, double* __restrict__ C1, long long ldc,       // x_{k+1} / aff
               double* __restrict__ D0, double* __restrict__ D1, long long ldd,       // acx (may be null)
               int relu, int write_x, int* __restrict__ flag_bad) {
  __shared__ double As[BK][BM];
  __shared__ double Bs0[BK][BN + 1];
  __shared__ double Bs1[MODE == 0 ? BK : 1][BN + 1];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  double acc0[4][4], acc1[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc0[i][j] = 0.0, acc1[i][j] = 0.0;

  for (int k0 = 0; k0 < Kdim; k0 += BK) {
    // A tile: BK x BM, m contiguous in global
#pragma unroll
    for (int i = 0; i < (BK * BM) / GEMM_THREADS; ++i) {
      const int idx = tid + i * GEMM_THREADS;
      const int m = idx % BM, k = idx / BM;
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < Kdim) ? A[gm + (long long)gk * lda] : 0.0;
    }
    // B tile: BK x BN, k contiguous in global (one query vector per column)
#pragma unroll
    for (int i = 0; i < (BK * BN) / GEMM_THREADS; ++i) {
      const int idx = tid + i * GEMM_THREADS;
      const int k = idx % BK, n = idx / BK;
      const int gn = n0 + n, gk = k0 + k;
      const bool ok = (gn < N && gk < Kdim);
      if (MODE == 0) {  // centre and radius of the input box: y = W c -/+ |W| r + b
        const double lo = ok ? B0[(long long)gn * ldb + gk] : 0.0, hi = ok ? B1[(long long)gn * ldb + gk] : 0.0;
        Bs0[k][n] = 0.5 * (lo + hi);
        Bs1[k][n] = 0.5 * (hi - lo);
      } else {
        Bs0[k][n] = ok ? B0[(long long)gn * ldb + gk] : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      double a[4], b0[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][tx * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b0[j] = Bs0[k][ty * 4 + j];
      if (MODE == 0) {
        double b1[4], aa[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b1[j] = Bs1[k][ty * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i) aa[i] = fabs(a[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc0[i][j] = fma(a[i], b0[j], acc0[i][j]);    // W c
            acc1[i][j] = fma(aa[i], b1[j], acc1[i][j]);   // |W| r
          }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc0[i][j] = fma(a[i], b0[j], acc0[i][j]);
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int gn = n0 + ty * 4 + j;
    if (gn >= N) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gm = m0 + tx * 4 + i;
      if (gm >= M) continue;
      if (MODE == 0) {
        const double bb = bias[gm];
        const double mid = acc0[i][j] + bb;
        const double ymin = mid - acc1[i][j], ymax = mid + acc1[i][j];
        if (!(ymin <= ymax) && flag_bad) atomicOr(flag_bad, 1);
        if (D0) {
          D0[(long long)gn * ldd + gm] = ymin;
          D1[(long long)gn * ldd + gm] = ymax;
        }
        if (write_x) {
          C0[(long long)gn * ldc + gm] = relu ? fmax(ymin, 0.0) : ymin;
          C1[(long long)gn * ldc + gm] = relu ? fmax(ymax, 0.0) : ymax;
        }
      } else {
        if (MODE == 2) C0[(long long)gn * ldc + gm] = acc0[i][j];
        else C0[(long long)gn * ldc + gm] += acc0[i][j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Few queries (N <= 32): the same contractions as HBM-streaming GEMVs.  One query on a wide layer is bound
// by reading W once (8 MB at width 1000); the 64 x 64 GEMM tiles above would put that on 16 CTAs.
// A cluster of GV_SPLIT CTAs owns 64 rows of A; each CTA streams its slice of the contraction index in
// 512 B column runs (8 k-lanes x 64 rows, GV_UNROLL loads in flight per thread), reduces its k-lanes in
// shared memory, and the cluster's partial sums meet through distributed shared memory in rank order, so
// the result is deterministic.  Queries go in groups of 8 (grid.z; later groups find A in L2) and every
// query accumulates on its own: its result does not depend on how many share the launch.
// ---------------------------------------------------------------------------------------------
constexpr int GV_ROWS = 64, GV_KL = 8, GV_THREADS = GV_ROWS * GV_KL, GV_UNROLL = 16, GV_SPLIT = 8;
constexpr int GV_KC = GV_KL * GV_UNROLL;   // contraction indices staged per pass
constexpr int GV_MAX_N = 32;
static bool gemv_enabled() {   // NNSDP_NO_GEMV=1: A/B switch back to the tiled GEMM for few queries
  static const bool on = [] { const char* e = getenv("NNSDP_NO_GEMV"); return !(e && atoi(e) != 0); }();
  return on;
}

// MODE 0: IBP layer, 1: C += A B, 2: C = A B.  NQT: queries per group held in registers (1, 2, 4, 8).
template <int MODE, int NQT>
__global__ void __cluster_dims__(1, GV_SPLIT, 1) __launch_bounds__(GV_THREADS, 1)
gemv_small_kernel(const double* __restrict__ A, int lda, int M, int Kdim,
                  const double* __restrict__ B0, const double* __restrict__ B1, long long ldb, int N,
                  const double* __restrict__ bias, double* __restrict__ C0, double* __restrict__ C1, long long ldc,
                  double* __restrict__ D0, double* __restrict__ D1, long long ldd, int relu, int write_x,
                  int* __restrict__ flag_bad) {
  constexpr int NACC = (MODE == 0) ? 2 : 1;
  __shared__ double xs[GV_KC][NACC][NQT];               // staged right-hand sides (centre, radius), zero padded
  __shared__ double red[GV_KL][NACC][GV_ROWS];          // one query's k-lane partials
  __shared__ double part[NQT][NACC][GV_ROWS];           // this CTA's slice sums, read by the cluster
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, r = tid & (GV_ROWS - 1), kl = tid / GV_ROWS;
  const int m0 = blockIdx.x * GV_ROWS;
  const int split = blockIdx.y;                         // == rank in the cluster
  const int q0 = blockIdx.z * NQT, nq = min(NQT, N - q0);
  const int kslice = (Kdim + GV_SPLIT - 1) / GV_SPLIT;
  const int kbeg = min(Kdim, split * kslice), kend = min(Kdim, kbeg + kslice);
  const double* Ar = A + min(m0 + r, M - 1);            // rows past M load a valid row and are dropped below
  double acc[NACC][NQT];
#pragma unroll
  for (int w = 0; w < NACC; ++w)
#pragma unroll
    for (int q = 0; q < NQT; ++q) acc[w][q] = 0.0;

  for (int kc = kbeg; kc < kend; kc += GV_KC) {
    // every load of A for this pass is in flight before anything waits on one
    double a[GV_UNROLL];
#pragma unroll
    for (int j = 0; j < GV_UNROLL; ++j) {
      const int k = min(kc + kl + j * GV_KL, kend - 1);
      a[j] = __ldg(Ar + (long long)k * lda);
    }
    __syncthreads();                                    // previous pass has finished with xs
    for (int i = tid; i < GV_KC * NQT; i += GV_THREADS) {
      const int kk = i / NQT, q = i % NQT, k = kc + kk;
      double v0 = 0.0, v1 = 0.0;
      if (k < kend && q < nq) {
        if (MODE == 0) {
          const double lo = B0[(long long)(q0 + q) * ldb + k], hi = B1[(long long)(q0 + q) * ldb + k];
          v0 = 0.5 * (lo + hi);
          v1 = 0.5 * (hi - lo);
        } else {
          v0 = B0[(long long)(q0 + q) * ldb + k];
        }
      }
      xs[kk][0][q] = v0;
      if (MODE == 0) xs[kk][1][q] = v1;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < GV_UNROLL; ++j) {
      const int kk = kl + j * GV_KL;
      const double av = (kc + kk < kend) ? a[j] : 0.0;
      const double aa = fabs(av);
#pragma unroll
      for (int q = 0; q < NQT; ++q) {
        acc[0][q] = fma(av, xs[kk][0][q], acc[0][q]);
        if (MODE == 0) acc[1][q] = fma(aa, xs[kk][1][q], acc[1][q]);
      }
    }
  }
  // k-lanes of this CTA, one query at a time (fixed order)
#pragma unroll
  for (int q = 0; q < NQT; ++q) {
    if (q < nq) {
      __syncthreads();
#pragma unroll
      for (int w = 0; w < NACC; ++w) red[kl][w][r] = acc[w][q];
      __syncthreads();
      if (tid < GV_ROWS * NACC) {
        const int rr = tid & (GV_ROWS - 1), w = tid / GV_ROWS;
        double sum = 0.0;
#pragma unroll
        for (int l = 0; l < GV_KL; ++l) sum += red[l][w][rr];
        part[q][w][rr] = sum;
      }
    }
  }
  cluster.sync();
  // CTA `split` finishes rows [split * 8, split * 8 + 8) of the tile for every query of the group
  constexpr int RPC = GV_ROWS / GV_SPLIT;
  if (tid < RPC * NQT) {
    const int rr = split * RPC + (tid % RPC), q = tid / RPC;
    const int m = m0 + rr;
    if (m < M && q < nq) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int c = 0; c < GV_SPLIT; ++c) {
        const double* rp = cluster.map_shared_rank(&part[0][0][0], c);
        s0 += rp[(q * NACC + 0) * GV_ROWS + rr];
        if (MODE == 0) s1 += rp[(q * NACC + NACC - 1) * GV_ROWS + rr];
      }
      const long long qg = q0 + q;
      if (MODE == 0) {
        const double mid = s0 + bias[m];
        const double ymin = mid - s1, ymax = mid + s1;
        if (!(ymin <= ymax) && flag_bad) atomicOr(flag_bad, 1);
        if (D0) {
          D0[qg * ldd + m] = ymin;
          D1[qg * ldd + m] = ymax;
        }
        if (write_x) {
          C0[qg * ldc + m] = relu ? fmax(ymin, 0.0) : ymin;
          C1[qg * ldc + m] = relu ? fmax(ymax, 0.0) : ymax;
        }
      } else if (MODE == 1) {
        C0[qg * ldc + m] += s0;
      } else {
        C0[qg * ldc + m] = s0;
      }
    }
  }
  cluster.sync();   // peers may still be reading this CTA's partial sums
}

template <int MODE>
static void gemv_launch(const double* A, int lda, int M, int Kdim, const double* B0, const double* B1, long long ldb,
                        int N, const double* bias, double* C0, double* C1, long long ldc, double* D0, double* D1,
                        long long ldd, int relu, int write_x, int* flag_bad, cudaStream_t st) {
  const int nqt = N == 1 ? 1 : N == 2 ? 2 : N <= 4 ? 4 : 8;
  const dim3 grid((M + GV_ROWS - 1) / GV_ROWS, GV_SPLIT, (N + nqt - 1) / nqt);
#define NNSDP_GEMV(T) gemv_small_kernel<MODE, T><<<grid, GV_THREADS, 0, st>>>(A, lda, M, Kdim, B0, B1, ldb, N, bias, C0, C1, ldc, D0, D1, ldd, relu, write_x, flag_bad)
  if (nqt == 1) NNSDP_GEMV(1);
  else if (nqt == 2) NNSDP_GEMV(2);
  else if (nqt == 4) NNSDP_GEMV(4);
  else NNSDP_GEMV(8);
#undef NNSDP_GEMV
}

// ---------------------------------------------------------------------------------------------
// Whole-net variants for few queries.  A launch (and a cluster barrier) costs more than streaming one
// 8 MB layer, so (a) the affine-column GEMVs of all layers, which are independent, go in ONE launch and
// (b) interval propagation, whose layers depend on each other, runs as ONE cooperative kernel with a grid
// barrier between layers.  Both use the tile routine below: ROWS x KL threads, each thread keeps
// UNROLL loads of A in flight, right-hand sides are staged in shared memory (zero padded), k-lanes
// are summed by warp shuffles and then across warps in a fixed order.
// ---------------------------------------------------------------------------------------------
constexpr int GVT_THREADS = 256;

template <int NACC, int NQT, int ROWS, int UNROLL>
struct GvtSmem {
  static constexpr int KL = GVT_THREADS / ROWS, KC = KL * UNROLL;
  double xs[KC][NACC][NQT];
  double red[GVT_THREADS / 32][NACC][NQT][ROWS];
};

// Sums over k of A[m0 + r, k] * x_q[k] (and |A| * rad_q[k] when NACC == 2) for the rows of one tile.
// On return thread tid < ROWS * NQT holds row m0 + tid % ROWS of query tid / ROWS in (s0, s1).
// COHERENT: the right-hand side was written earlier in this kernel by other CTAs (no read-only path).
template <int NACC, int NQT, int ROWS, int UNROLL, bool COHERENT>
__device__ __forceinline__ void gemv_tile(GvtSmem<NACC, NQT, ROWS, UNROLL>& sm, const double* __restrict__ A, int lda, int M,
                                          int Kdim, int m0, const double* B0, const double* B1, long long ldb, int nq,
                                          double& s0, double& s1) {
  constexpr int KL = GVT_THREADS / ROWS, KC = KL * UNROLL;
  const int tid = threadIdx.x, r = tid % ROWS, kl = tid / ROWS;
  const double* Ar = A + min(m0 + r, M - 1);            // rows past M load a valid row and are dropped by the caller
  double acc[NACC][NQT];
#pragma unroll
  for (int w = 0; w < NACC; ++w)
#pragma unroll
    for (int q = 0; q < NQT; ++q) acc[w][q] = 0.0;
  for (int kc = 0; kc < Kdim; kc += KC) {
    double a[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) a[j] = __ldg(Ar + (long long)min(kc + kl + j * KL, Kdim - 1) * lda);
    __syncthreads();                                    // the previous pass / tile has finished with xs and red
    for (int i = tid; i < KC * NQT; i += GVT_THREADS) {
      const int kk = i / NQT, q = i % NQT, k = kc + kk;
      double v0 = 0.0, v1 = 0.0;
      if (k < Kdim && q < nq) {
        const double* p0 = B0 + (long long)q * ldb + k;
        if (NACC == 2) {
          const double* p1 = B1 + (long long)q * ldb + k;
          const double lo = COHERENT ? __ldcg(p0) : *p0, hi = COHERENT ? __ldcg(p1) : *p1;
          v0 = 0.5 * (lo + hi);
          v1 = 0.5 * (hi - lo);
        } else {
          v0 = COHERENT ? __ldcg(p0) : *p0;
        }
      }
      sm.xs[kk][0][q] = v0;
      if (NACC == 2) sm.xs[kk][NACC - 1][q] = v1;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const int kk = kl + j * KL;
      const double av = (kc + kk < Kdim) ? a[j] : 0.0;
      const double aa = fabs(av);
#pragma unroll
      for (int q = 0; q < NQT; ++q) {
        acc[0][q] = fma(av, sm.xs[kk][0][q], acc[0][q]);
        if (NACC == 2) acc[NACC - 1][q] = fma(aa, sm.xs[kk][NACC - 1][q], acc[NACC - 1][q]);
      }
    }
  }
  // k-lanes inside a warp (lane bits above log2(ROWS)), then the warps in order
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int w = 0; w < NACC; ++w)
#pragma unroll
    for (int q = 0; q < NQT; ++q) {
      double v = acc[w][q];
#pragma unroll
      for (int o = ROWS; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane < ROWS) sm.red[warp][w][q][lane] = v;
    }
  __syncthreads();
  s0 = s1 = 0.0;
  if (tid < ROWS * NQT) {
    const int rr = tid % ROWS, q = tid / ROWS;
#pragma unroll
    for (int w = 0; w < GVT_THREADS / 32; ++w) {
      s0 += sm.red[w][0][q][rr];
      if (NACC == 2) s1 += sm.red[w][NACC - 1][q][rr];
    }
  }
}

// (a) aff[off[blk] + m] += sum_j Wt_blk[m, j] u[noff(blk + 1) + j] for every block: grid (row tiles, K - 1, query groups)
constexpr int AFA_ROWS = 16, AFA_UNROLL = 16;
template <int NQT>
__global__ void __launch_bounds__(GVT_THREADS)
affine_all_kernel(NetDev net, const double* __restrict__ u, long long u_stride, double* __restrict__ aff,
                  long long aff_stride, int N) {
  __shared__ GvtSmem<1, NQT, AFA_ROWS, AFA_UNROLL> sm;
  const int blk = blockIdx.y, M = net.n[blk], m0 = blockIdx.x * AFA_ROWS;
  if (m0 >= M) return;
  const int q0 = blockIdx.z * NQT, nq = min(NQT, N - q0);
  double s0, s1;
  gemv_tile<1, NQT, AFA_ROWS, AFA_UNROLL, false>(sm, net.Wt[blk], net.ldT[blk], M, net.n[blk + 1], m0,
                                     u + (net.off[blk + 1] - net.n_in) + (long long)q0 * u_stride, nullptr, u_stride, nq,
                                     s0, s1);
  const int tid = threadIdx.x;
  if (tid < AFA_ROWS * NQT) {
    const int m = m0 + tid % AFA_ROWS, q = tid / AFA_ROWS;
    if (m < M && q < nq) aff[(long long)(q0 + q) * aff_stride + net.off[blk] + m] += s0;
  }
}

// (b) interval propagation through every layer (intervals_easy.jl:2-37) + sector slopes (activ_sector.jl:63-72)
constexpr int IBA_ROWS = 8, IBA_MAX_N = 8;
template <int NQT> struct IbaUnroll { static constexpr int value = NQT <= 4 ? 16 : 8; };   // staged right-hand sides stay under 48 KB
template <int NQT>
__global__ void __launch_bounds__(GVT_THREADS)
ibp_all_kernel(NetDev net, const double* __restrict__ x1min, long long s_min, const double* __restrict__ x1max,
               long long s_max, double* xmin, double* xmax, long long x_stride, double* acxmin, double* acxmax,
               double* smin, double* smax, long long acx_stride, int N, int* flag_bad) {
  __shared__ GvtSmem<2, NQT, IBA_ROWS, IbaUnroll<NQT>::value> sm;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, K = net.K, n_in = net.n_in;
  const int groups = (N + NQT - 1) / NQT;
  for (long long i = blockIdx.x * (long long)GVT_THREADS + tid; i < (long long)n_in * N; i += (long long)gridDim.x * GVT_THREADS) {
    const long long q = i / n_in, r = i % n_in;
    xmin[q * x_stride + r] = x1min[q * s_min + r];
    xmax[q * x_stride + r] = x1max[q * s_max + r];
  }
  grid.sync();
  for (int k = 0; k < K; ++k) {
    const int M = net.n[k + 1], Kdim = net.n[k];
    const double* A = net.M[k];
    const double* bias = A + (long long)Kdim * M;
    const bool last = (k == K - 1);
    const int row_tiles = (M + IBA_ROWS - 1) / IBA_ROWS;
    for (int t = blockIdx.x; t < row_tiles * groups; t += gridDim.x) {
      const int m0 = (t % row_tiles) * IBA_ROWS, q0 = (t / row_tiles) * NQT, nq = min(NQT, N - q0);
      const double* b0 = xmin + net.xoff[k] + (long long)q0 * x_stride;
      const double* b1 = xmax + net.xoff[k] + (long long)q0 * x_stride;
      double s0, s1;
      gemv_tile<2, NQT, IBA_ROWS, IbaUnroll<NQT>::value, true>(sm, A, M, M, Kdim, m0, b0, b1, x_stride, nq, s0, s1);
      if (tid < IBA_ROWS * NQT) {
        const int m = m0 + tid % IBA_ROWS, q = tid / IBA_ROWS;
        if (m < M && q < nq) {
          const long long qg = q0 + q;
          const double mid = s0 + bias[m];
          const double ymin = mid - s1, ymax = mid + s1;
          if (!(ymin <= ymax) && flag_bad) atomicOr(flag_bad, 1);
          if (!last) {
            const long long o = qg * acx_stride + (net.off[k + 1] - n_in) + m;
            acxmin[o] = ymin;
            acxmax[o] = ymax;
            const double eps = 1e-4;  // activ_sector.jl:65
            smin[o] = (ymin > eps) ? 1.0 : 0.0;
            smax[o] = (ymax < -eps) ? 0.0 : 1.0;
          }
          xmin[qg * x_stride + net.xoff[k + 1] + m] = last ? ymin : fmax(ymin, 0.0);
          xmax[qg * x_stride + net.xoff[k + 1] + m] = last ? ymax : fmax(ymax, 0.0);
        }
      }
    }
    if (!last) grid.sync();
  }
}

// (c) narrow nets (every width <= 256), any number of queries: queries are independent, so a CTA takes ICH_NQ of
// them through ALL layers by itself -- no barrier between CTAs, one launch for the whole propagation.  Thread m owns
// output row m of the current layer (W is read column by column, coalesced, from L2 after the first CTA), the
// (centre, radius) of the layer input sits in shared memory, double buffered.
constexpr int ICH_NQ = 8, ICH_MAXW = 256, ICH_UNROLL = 16;

__global__ void __launch_bounds__(ICH_MAXW)
ibp_chain_kernel(NetDev net, int maxw, const double* __restrict__ x1min, long long s_min, const double* __restrict__ x1max,
                 long long s_max, double* __restrict__ xmin, double* __restrict__ xmax, long long x_stride,
                 double* __restrict__ acxmin, double* __restrict__ acxmax, double* __restrict__ smin,
                 double* __restrict__ smax, long long acx_stride, int N, int* flag_bad) {
  extern __shared__ double ich[];                       // [2 buffers][maxw][2][ICH_NQ]
  const int tid = threadIdx.x, K = net.K, n_in = net.n_in;
  const int q0 = blockIdx.x * ICH_NQ, nq = min(ICH_NQ, N - q0);
  double* cur = ich;
  double* nxt = ich + (size_t)maxw * 2 * ICH_NQ;
  for (int i = tid; i < n_in * ICH_NQ; i += blockDim.x) {
    const int r = i / ICH_NQ, q = i % ICH_NQ;
    double lo = 0.0, hi = 0.0;
    if (q < nq) {
      lo = x1min[(long long)(q0 + q) * s_min + r];
      hi = x1max[(long long)(q0 + q) * s_max + r];
      xmin[(long long)(q0 + q) * x_stride + r] = lo;
      xmax[(long long)(q0 + q) * x_stride + r] = hi;
    }
    cur[(r * 2 + 0) * ICH_NQ + q] = 0.5 * (lo + hi);
    cur[(r * 2 + 1) * ICH_NQ + q] = 0.5 * (hi - lo);
  }
  __syncthreads();
  for (int k = 0; k < K; ++k) {
    const int M = net.n[k + 1], Kdim = net.n[k];
    const double* A = net.M[k];
    const bool last = (k == K - 1);
    {
      double ac[ICH_NQ], ar[ICH_NQ];
#pragma unroll
      for (int q = 0; q < ICH_NQ; ++q) ac[q] = ar[q] = 0.0;
      const double* Ar = A + min(tid, M - 1);            // idle threads load a valid row: the barrier below is uniform
      for (int k0 = 0; k0 < Kdim; k0 += ICH_UNROLL) {     // ICH_UNROLL loads of W in flight per thread
        double a[ICH_UNROLL];
#pragma unroll
        for (int j = 0; j < ICH_UNROLL; ++j) a[j] = __ldcg(Ar + (long long)min(k0 + j, Kdim - 1) * M);   // not .nc: stays above the barrier
        __syncthreads();                                  // all ICH_UNROLL loads are issued before the first is consumed
#pragma unroll
        for (int j = 0; j < ICH_UNROLL; ++j) {
          const double av = (k0 + j < Kdim) ? a[j] : 0.0;
          const double aa = fabs(av);
          const double* x = cur + min(k0 + j, Kdim - 1) * 2 * ICH_NQ;
#pragma unroll
          for (int q = 0; q < ICH_NQ; ++q) {
            ac[q] = fma(av, x[q], ac[q]);
            ar[q] = fma(aa, x[ICH_NQ + q], ar[q]);
          }
        }
      }
      const double bias = A[(long long)Kdim * M + min(tid, M - 1)];
#pragma unroll
      for (int q = 0; q < ICH_NQ && tid < M; ++q) {
        const double mid = ac[q] + bias;
        const double ymin = mid - ar[q], ymax = mid + ar[q];
        const double lo = last ? ymin : fmax(ymin, 0.0), hi = last ? ymax : fmax(ymax, 0.0);
        nxt[(tid * 2 + 0) * ICH_NQ + q] = 0.5 * (lo + hi);
        nxt[(tid * 2 + 1) * ICH_NQ + q] = 0.5 * (hi - lo);
        if (q < nq) {
          const long long qg = q0 + q;
          if (!(ymin <= ymax) && flag_bad) atomicOr(flag_bad, 1);
          if (!last) {
            const long long o = qg * acx_stride + (net.off[k + 1] - n_in) + tid;
            acxmin[o] = ymin;
            acxmax[o] = ymax;
            const double eps = 1e-4;  // activ_sector.jl:65
            smin[o] = (ymin > eps) ? 1.0 : 0.0;
            smax[o] = (ymax < -eps) ? 0.0 : 1.0;
          }
          xmin[qg * x_stride + net.xoff[k + 1] + tid] = lo;
          xmax[qg * x_stride + net.xoff[k + 1] + tid] = hi;
        }
      }
    }
    __syncthreads();
    double* t = cur;
    cur = nxt;
    nxt = t;
  }
}

__global__ void sector_minmax_kernel(long long n, const double* __restrict__ lo,
                                     const double* __restrict__ hi, double* __restrict__ smin,
                                     double* __restrict__ smax) {
  const double eps = 1e-4;  // activ_sector.jl:65
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    smin[i] = (lo[i] > eps) ? 1.0 : 0.0;
    smax[i] = (hi[i] < -eps) ? 0.0 : 1.0;
  }
}

// x_1 bounds of every query into the stacked x layout (stride 0 = one box shared by all queries)
__global__ void place_x1_kernel(const double* __restrict__ x1min, long long s_min,
                                const double* __restrict__ x1max, long long s_max,
                                double* __restrict__ xmin, double* __restrict__ xmax,
                                long long xtot, int n_in, int Q) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)n_in * Q) return;
  const long long q = i / n_in, r = i % n_in;
  xmin[q * xtot + r] = x1min[q * s_min + r];
  xmax[q * xtot + r] = x1max[q * s_max + r];
}

// Wt[r + j*ldT] = W[j, r] : input-contiguous copy of W_k (rows r >= n_in_k stay zero).
__global__ void transpose_w_kernel(const double* __restrict__ Mk, int n_out_k, int n_in_k,
                                   double* __restrict__ Wt, int ldT) {
  __shared__ double tile[32][33];
  const int j0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {  // i: input index r, x: neuron j
    const int j = j0 + threadIdx.x, r = r0 + i;
    tile[i][threadIdx.x] = (j < n_out_k && r < n_in_k) ? Mk[j + (long long)r * n_out_k] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {  // i: neuron j, x: input index r
    const int j = j0 + i, r = r0 + threadIdx.x;
    if (j < n_out_k && r < n_in_k) Wt[r + (long long)j * ldT] = tile[threadIdx.x][i];
  }
}

}  // namespace

int launch_place_x1(const double* x1min, long long s_min, const double* x1max, long long s_max,
                    double* xmin, double* xmax, long long xtot, int n_in, int Q, cudaStream_t st) {
  const long long n = (long long)n_in * Q;
  place_x1_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(x1min, s_min, x1max, s_max, xmin, xmax,
                                                          xtot, n_in, Q);
  return 1;
}

int launch_transpose_w(const double* Mk, int n_out_k, int n_in_k, double* Wt, int ldT,
                       cudaStream_t st) {
  dim3 grid((n_out_k + 31) / 32, (n_in_k + 31) / 32), block(32, 8);
  transpose_w_kernel<<<grid, block, 0, st>>>(Mk, n_out_k, n_in_k, Wt, ldT);
  return 1;
}

int ibp_layer_launch(const double* Mk, int n_out_k, int n_in_k, const double* xin_min,
                     const double* xin_max, long long x_stride, double* xout_min, double* xout_max,
                     double* acx_min, double* acx_max, long long acx_stride, int Q, int relu,
                     int write_x, int* flag_bad, cudaStream_t st) {
  if (Q <= GV_MAX_N && gemv_enabled()) {
    gemv_launch<0>(Mk, n_out_k, n_out_k, n_in_k, xin_min, xin_max, x_stride, Q, Mk + (long long)n_in_k * n_out_k, xout_min,
        xout_max, x_stride, acx_min, acx_max, acx_stride, relu, write_x, flag_bad, st);
    return 1;
  }
  if (ibp_dmma_launch(Mk, n_out_k, n_in_k, xin_min, xin_max, x_stride, xout_min, xout_max, acx_min, acx_max, acx_stride, Q,
                      relu, write_x, flag_bad, st))
    return 1;
  dim3 grid((n_out_k + BM - 1) / BM, (Q + BN - 1) / BN);
  gemm_nn_kernel<0><<<grid, GEMM_THREADS, 0, st>>>(
      Mk, n_out_k, n_out_k, n_in_k, xin_min, xin_max, x_stride, Q, Mk + (long long)n_in_k * n_out_k,
      xout_min, xout_max, x_stride, acx_min, acx_max, acx_stride, relu, write_x, flag_bad);
  return 1;
}

int affine_layer_launch(const double* Wt, int ldT, int n_rows, int n_neurons, const double* u,
                        long long u_stride, double* aff, long long aff_stride, int Q,
                        cudaStream_t st) {
  if (Q <= GV_MAX_N && gemv_enabled()) {
    gemv_launch<1>(Wt, ldT, n_rows, n_neurons, u, nullptr, u_stride, Q, nullptr, aff, nullptr, aff_stride, nullptr, nullptr, 0, 0, 0,
        nullptr, st);
    return 1;
  }
  if (dgemm_dmma_launch(Wt, ldT, n_rows, n_neurons, u, u_stride, aff, aff_stride, Q, 1, st)) return 1;
  dim3 grid((n_rows + BM - 1) / BM, (Q + BN - 1) / BN);
  gemm_nn_kernel<1><<<grid, GEMM_THREADS, 0, st>>>(Wt, ldT, n_rows, n_neurons, u, nullptr, u_stride,
                                                  Q, nullptr, aff, nullptr, aff_stride, nullptr,
                                                  nullptr, 0, 0, 0, nullptr);
  return 1;
}

int gemm_set_launch(const double* A, int lda, int M, int Kdim, const double* B, long long ldb,
                    double* C, long long ldc, int N, cudaStream_t st) {
  if (N <= GV_MAX_N && gemv_enabled()) {
    gemv_launch<2>(A, lda, M, Kdim, B, nullptr, ldb, N, nullptr, C, nullptr, ldc, nullptr, nullptr, 0, 0, 0, nullptr,
                   st);
    return 1;
  }
  if (dgemm_dmma_launch(A, lda, M, Kdim, B, ldb, C, ldc, N, 0, st)) return 1;
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
  gemm_nn_kernel<2><<<grid, GEMM_THREADS, 0, st>>>(A, lda, M, Kdim, B, nullptr, ldb, N, nullptr, C, nullptr,
                                                  ldc, nullptr, nullptr, 0, 0, 0, nullptr);
  return 1;
}

int gemm_acc_launch(const double* A, int lda, int M, int Kdim, const double* B, long long ldb,
                    double* C, long long ldc, int N, cudaStream_t st) {
  if (N <= GV_MAX_N && gemv_enabled()) {
    gemv_launch<1>(A, lda, M, Kdim, B, nullptr, ldb, N, nullptr, C, nullptr, ldc, nullptr, nullptr, 0, 0, 0, nullptr,
                   st);
    return 1;
  }
  if (dgemm_dmma_launch(A, lda, M, Kdim, B, ldb, C, ldc, N, 1, st)) return 1;
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
  gemm_nn_kernel<1><<<grid, GEMM_THREADS, 0, st>>>(A, lda, M, Kdim, B, nullptr, ldb, N, nullptr, C, nullptr,
                                                  ldc, nullptr, nullptr, 0, 0, 0, nullptr);
  return 1;
}

int launch_sector_minmax(long long n, const double* acxmin, const double* acxmax, double* smin,
                         double* smax, cudaStream_t st) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sector_minmax_kernel<<<(int)blocks, 256, 0, st>>>(n, acxmin, acxmax, smin, smax);
  return 1;
}

// All affine-column GEMVs of a batch of few queries in one launch; 0 = not applicable (caller loops over layers).
int affine_all_launch(const NetDev& nd, int K, int max_rows, const double* u, long long u_stride, double* aff,
                      long long aff_stride, int Q, cudaStream_t st) {
  // few queries, or narrow nets (the per-layer launches cost more than their products)
  if ((Q > GV_MAX_N && max_rows > ICH_MAXW) || !gemv_enabled() || K < 2) return 0;
  const int nqt = Q == 1 ? 1 : Q == 2 ? 2 : Q <= 4 ? 4 : 8;
  const dim3 grid((max_rows + AFA_ROWS - 1) / AFA_ROWS, K - 1, (Q + nqt - 1) / nqt);
  if (nqt == 1) affine_all_kernel<1><<<grid, GVT_THREADS, 0, st>>>(nd, u, u_stride, aff, aff_stride, Q);
  else if (nqt == 2) affine_all_kernel<2><<<grid, GVT_THREADS, 0, st>>>(nd, u, u_stride, aff, aff_stride, Q);
  else if (nqt == 4) affine_all_kernel<4><<<grid, GVT_THREADS, 0, st>>>(nd, u, u_stride, aff, aff_stride, Q);
  else affine_all_kernel<8><<<grid, GVT_THREADS, 0, st>>>(nd, u, u_stride, aff, aff_stride, Q);
  return 1;
}

template <int NQT>
static int ibp_all_launch_t(const NetDev& nd, int max_out, const double* x1min, long long s_min, const double* x1max,
                            long long s_max, double* xmin, double* xmax, long long x_stride, double* acxmin,
                            double* acxmax, double* smin, double* smax, long long acx_stride, int Q, int* flag_bad,
                            cudaStream_t st) {
  static std::atomic<int> resident{0};   // co-resident CTAs of this kernel (the devices of a context are alike)
  int cap = resident.load(std::memory_order_relaxed);
  if (cap == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ibp_all_kernel<NQT>, GVT_THREADS, 0) != cudaSuccess) return -1;
    if (per_sm < 1 || sms < 1) return -1;
    cap = per_sm * sms;
    resident.store(cap, std::memory_order_relaxed);
  }
  const int groups = (Q + NQT - 1) / NQT;
  const long long tiles = (long long)((max_out + IBA_ROWS - 1) / IBA_ROWS) * groups;
  const int grid = (int)std::max<long long>(1, std::min<long long>(cap, tiles));
  NetDev ndv = nd;
  void* args[] = {&ndv, &x1min, &s_min, &x1max, &s_max, &xmin, &xmax, &x_stride, &acxmin, &acxmax,
                  &smin, &smax, &acx_stride, &Q, &flag_bad};
  if (cudaLaunchCooperativeKernel((const void*)ibp_all_kernel<NQT>, dim3(grid), dim3(GVT_THREADS), args, 0, st) != cudaSuccess)
    return -1;
  return 1;
}

// Interval propagation through the whole net (+ sector slopes) for few queries as one cooperative kernel;
// 0 = not applicable (caller launches layer by layer), < 0 = CUDA error.
int ibp_all_launch(const NetDev& nd, int max_out, const double* x1min, long long s_min, const double* x1max,
                   long long s_max, double* xmin, double* xmax, long long x_stride, double* acxmin, double* acxmax,
                   double* smin, double* smax, long long acx_stride, int Q, int* flag_bad, cudaStream_t st) {
  if (Q > IBA_MAX_N || !gemv_enabled()) return 0;   // more query groups: per-layer cluster GEMVs are faster
#define NNSDP_IBA(T) ibp_all_launch_t<T>(nd, max_out, x1min, s_min, x1max, s_max, xmin, xmax, x_stride, acxmin, acxmax, smin, smax, acx_stride, Q, flag_bad, st)
  if (Q == 1) return NNSDP_IBA(1);
  if (Q == 2) return NNSDP_IBA(2);
  if (Q <= 4) return NNSDP_IBA(4);
  return NNSDP_IBA(8);
#undef NNSDP_IBA
}

// Interval propagation of narrow nets (max width <= 256) for any number of queries: one launch, a CTA per 8 queries
// walks every layer.  0 = not applicable.
int ibp_chain_launch(const NetDev& nd, int max_w, const double* x1min, long long s_min, const double* x1max,
                     long long s_max, double* xmin, double* xmax, long long x_stride, double* acxmin, double* acxmax,
                     double* smin, double* smax, long long acx_stride, int Q, int* flag_bad, cudaStream_t st) {
  static const bool off = [] { const char* e = getenv("NNSDP_NO_IBP_CHAIN"); return e && atoi(e) != 0; }();
  if (off || max_w > ICH_MAXW || !gemv_enabled()) return 0;
  const int threads = max_w <= 64 ? 64 : max_w <= 128 ? 128 : 256;
  const size_t smem = (size_t)2 * max_w * 2 * ICH_NQ * sizeof(double);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(ibp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ibp_chain_kernel<<<(Q + ICH_NQ - 1) / ICH_NQ, threads, smem, st>>>(nd, max_w, x1min, s_min, x1max, s_max, xmin, xmax,
                                                                   x_stride, acxmin, acxmax, smin, smax, acx_stride, Q,
                                                                   flag_bad);
  return 1;
}

}  // namespace nnsdp
