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
  dim3 grid((n_out_k + BM - 1) / BM, (Q + BN - 1) / BN);
  gemm_nn_kernel<0><<<grid, GEMM_THREADS, 0, st>>>(
      Mk, n_out_k, n_out_k, n_in_k, xin_min, xin_max, x_stride, Q, Mk + (long long)n_in_k * n_out_k,
      xout_min, xout_max, x_stride, acx_min, acx_max, acx_stride, relu, write_x, flag_bad);
  return 1;
}

int affine_layer_launch(const double* Wt, int ldT, int n_rows, int n_neurons, const double* u,
                        long long u_stride, double* aff, long long aff_stride, int Q,
                        cudaStream_t st) {
  dim3 grid((n_rows + BM - 1) / BM, (Q + BN - 1) / BN);
  gemm_nn_kernel<1><<<grid, GEMM_THREADS, 0, st>>>(Wt, ldT, n_rows, n_neurons, u, nullptr, u_stride,
                                                  Q, nullptr, aff, nullptr, aff_stride, nullptr,
                                                  nullptr, 0, 0, 0, nullptr);
  return 1;
}

int gemm_set_launch(const double* A, int lda, int M, int Kdim, const double* B, long long ldb,
                    double* C, long long ldc, int N, cudaStream_t st) {
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
  gemm_nn_kernel<2><<<grid, GEMM_THREADS, 0, st>>>(A, lda, M, Kdim, B, nullptr, ldb, N, nullptr, C, nullptr,
                                                  ldc, nullptr, nullptr, 0, 0, 0, nullptr);
  return 1;
}

int gemm_acc_launch(const double* A, int lda, int M, int Kdim, const double* B, long long ldb,
                    double* C, long long ldc, int N, cudaStream_t st) {
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

}  // namespace nnsdp
