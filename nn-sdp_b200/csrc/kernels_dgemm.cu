// FP64 tensor-core GEMM for the plain products of the path:  C (+)= A B,  all column-major.
// Used by the CROWN chain steps on wide layers (A = W_k', B = the stack of relaxed rows: N = 2 x queries x rows),
// by the affine-column product of the QC preparation and by the factored matvec of the lambda_max check when
// many queries share a launch.  tcgen05 has no f64 kind, so this is the DMMA path (mma.sync.m8n8k4.f64), the same
// building block as the Gram kernel (kernels_gram.cu): 128 x 64 tiles, 8 warps of 32 x 32, two CTAs per SM, 3 cp.async stages.
//   A tile: As[k][m]   (columns of A are contiguous in m: 16 B chunks along m)
//   B tile: Bs[n][k]   (columns of B are contiguous in k: 16 B chunks along k; leading dimension 20 doubles puts
//                       the 4 x 4 (k, n) addresses of a half warp's fragment load in 16 distinct 8-byte banks)
// Edges are zero-filled by the copies (src-size 0 / 8 / 16).  Operands must allow 16 B copies: A, B 16 B aligned and
// lda, ldb even; otherwise the caller keeps the SIMT kernel (kernels_bounds.cu).
#include "internal.h"
#include <stdint.h>
#include <stdlib.h>

namespace nnsdp {

namespace {

constexpr int DT = 128;         // tile side
#ifndef NNSDP_DGEMM_DK
#define NNSDP_DGEMM_DK 16
#endif
constexpr int DK = NNSDP_DGEMM_DK;  // contraction indices per stage
constexpr int DSTAGES = 3;
#ifndef NNSDP_DGEMM_LDA_PAD
#define NNSDP_DGEMM_LDA_PAD 4
#endif
#ifndef NNSDP_DGEMM_WN
#define NNSDP_DGEMM_WN 32
#endif
constexpr int DLDA = DT + NNSDP_DGEMM_LDA_PAD;  // As leading dimension: 4 (mod 16) puts the 4 x 4 (k, m) addresses of a half
                                                // warp's fragment load in 16 distinct 8-byte banks (8 gives 2-way conflicts)
constexpr int DLDB = DK + 4;    // Bs leading dimension
constexpr int DWN = NNSDP_DGEMM_WN;             // columns per warp tile (rows: 32)
constexpr int DNJ = DWN / 8;
#ifndef NNSDP_DGEMM_TN
#define NNSDP_DGEMM_TN 64
#endif
#ifndef NNSDP_DGEMM_MINB
#define NNSDP_DGEMM_MINB 2
#endif
constexpr int DTN = NNSDP_DGEMM_TN;             // tile columns (rows: DT).  128 x 64 tiles, 8 warps, two CTAs per SM: one CTA's
                                                // barrier and pipeline fill hide under the other's DMMA stream (CROWN W1000-D20
                                                // 68.3 -> 66.1 ms per 2 queries, affine launch 2.49 -> 2.31 ms against 128 x 128 x 1)
constexpr int DTHREADS = 4 * (DTN / DWN) * 32;   // 16 warps of 32 x 32: the DMMA issue cadence of a warp leaves the pipe half
                                                // idle with 2 warps per scheduler (8 warps of 32 x 64)

struct DgemmSmem {
  double A[DSTAGES][DK][DLDA];
  double B[DSTAGES][DTN][DLDB];
};

__device__ __forceinline__ void cp16(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int ACC>  // 0: C = A B, 1: C += A B
__device__ __forceinline__ void dgemm_tile(const double* __restrict__ A, int lda, int M, int K, const double* __restrict__ B,
                                           long long ldb, double* __restrict__ C, long long ldc, int N) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DgemmSmem& sm = *reinterpret_cast<DgemmSmem*>(smem_raw);
  const int m0 = blockIdx.x * DT, n0 = blockIdx.y * DTN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * DWN;
  const int nsteps = (K + DK - 1) / DK;

  auto load_stage = [&](int stage, int step) {
    const int k0 = step * DK;
    // A: 16 columns x 64 chunks of two rows
    for (int c = tid; c < DK * (DT / 2); c += DTHREADS) {
      const int kk = c / (DT / 2), ch = c % (DT / 2);
      const int m = m0 + ch * 2, k = k0 + kk;
      const int rows = (k < K) ? max(0, min(2, M - m)) : 0;      // 0 bytes read: the chunk is zero-filled
      cp16(&sm.A[stage][kk][ch * 2], rows ? A + (long long)k * lda + m : A, rows * 8);
    }
    // B: 128 columns x 8 chunks of two contraction indices
    for (int c = tid; c < DTN * (DK / 2); c += DTHREADS) {
      const int nn = c / (DK / 2), ch = c % (DK / 2);
      const int n = n0 + nn, k = k0 + ch * 2;
      const int cnt = (n < N) ? max(0, min(2, K - k)) : 0;
      cp16(&sm.B[stage][nn][ch * 2], cnt ? B + (long long)n * ldb + k : B, cnt * 8);
    }
  };

  double acc[4][DNJ][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < DNJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int s = 0; s < DSTAGES - 1; ++s) {
    if (s < nsteps) load_stage(s, s);
    cp_commit();
  }
  for (int step = 0; step < nsteps; ++step) {
    cp_wait<DSTAGES - 2>();
    __syncthreads();
    {
      const int nxt = step + DSTAGES - 1;
      if (nxt < nsteps) load_stage(nxt % DSTAGES, nxt);
      cp_commit();
    }
    const int st = step % DSTAGES;
    const double(*As)[DLDA] = sm.A[st];
    const double(*Bs)[DLDB] = sm.B[st];
#pragma unroll
    for (int k4 = 0; k4 < DK; k4 += 4) {
      const int k = k4 + (lane & 3);
      double af[4], bf[DNJ];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = As[k][wm + i * 8 + (lane >> 2)];
#pragma unroll
      for (int j = 0; j < DNJ; ++j) bf[j] = Bs[wn + j * 8 + (lane >> 2)][k];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < DNJ; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_wait<0>();

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + wm + i * 8 + (lane >> 2);
#pragma unroll
    for (int j = 0; j < DNJ; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = n0 + wn + j * 8 + (lane & 3) * 2 + e;
        if (r < M && c < N) {
          double* p = C + r + (long long)c * ldc;
          *p = ACC ? *p + acc[i][j][e] : acc[i][j][e];
        }
      }
    }
  }
}

template <int ACC>
__global__ void __launch_bounds__(DTHREADS, NNSDP_DGEMM_MINB)
dgemm_dmma_kernel(const double* __restrict__ A, int lda, int M, int K, const double* __restrict__ B, long long ldb,
                  double* __restrict__ C, long long ldc, int N) {
  dgemm_tile<ACC>(A, lda, M, K, B, ldb, C, ldc, N);
}

// The affine-column products of ALL layers in one launch (blockIdx.z = layer): aff[off_b ..] += Wt_b u_{b+1} for
// b = 0 .. K-2.  The products are independent; one layer alone is 8 x 8 tiles at width 1000 and Q = 1024 -- 64 CTAs on
// 148 SMs -- so launched layer by layer they run at 10 TFLOP/s.
__global__ void __launch_bounds__(DTHREADS, NNSDP_DGEMM_MINB)
dgemm_dmma_affine_layers_kernel(NetDev net, int b0, const double* __restrict__ u, long long u_stride,
                                double* __restrict__ aff, long long aff_stride, int Q) {
  const int b = b0 + blockIdx.z;
  const int M = net.n[b], Kd = net.n[b + 1];
  if ((int)blockIdx.x * DT >= M) return;
  dgemm_tile<1>(net.Wt[b], net.ldT[b], M, Kd, u + (net.off[b + 1] - net.n_in), u_stride, aff + net.off[b], aff_stride, Q);
}

// ---------------------------------------------------------------------------------------------
// Interval propagation of one wide layer for many boxes (intervalsWorstCase,
// /root/reference/src/Intervals/intervals_easy.jl:21-33) on the FP64 tensor cores.  In centre / radius form
//   ymin = W c + b - |W| r,   ymax = W c + b + |W| r,   c = (xmin + xmax) / 2,  r = (xmax - xmin) / 2
// the two products share their A operand: a CTA owns 128 neurons x 64 boxes with TWO accumulator sets, loads the
// fragment of W once and feeds it to the second product through fabs.  The tiles of xmin / xmax are staged as they
// are (cp.async cannot transform) and turned into c / r when the B fragments are read.  The epilogue is the one of
// the SIMT kernel (kernels_bounds.cu): bias, the ordering check, pre-activation bounds, ReLU'd post-activation bounds.
// 16 warps of 32 x 16: at width 1000 and 1024 boxes that is 8 x 16 = 128 CTAs for 148 SMs in a single wave.
// ---------------------------------------------------------------------------------------------
constexpr int IT_N = 64;                     // boxes per CTA

// WM = warps along the neurons: the CTA owns 32 WM neurons x 64 boxes with 4 WM warps of 32 x 16.  WM = 2 at the stress
// size (256 CTAs for 1000 neurons x 1024 boxes, two per SM); batches of a few hundred boxes take WM = 1 so that the launch
// still covers the SMs (a CTA is bound by its own DMMA stream: 1000 neurons x 128 boxes as 16 CTAs of WM = 4 take
// 0.19 ms per layer whatever the rest of the GPU does).
template <int WM>
struct IbpSmem {
  static constexpr int TM = 32 * WM, LDA = TM + 4;
  double A[DSTAGES][DK][LDA];
  double Blo[DSTAGES][IT_N][DLDB];
  double Bhi[DSTAGES][IT_N][DLDB];
};

template <int WM>
__global__ void __launch_bounds__(128 * WM, 1)
ibp_dmma_kernel(const double* __restrict__ A, int lda, int M, int K, const double* __restrict__ Xlo,
                const double* __restrict__ Xhi, long long ldb, int N, const double* __restrict__ bias,
                double* __restrict__ C0, double* __restrict__ C1, long long ldc, double* __restrict__ D0,
                double* __restrict__ D1, long long ldd, int relu, int write_x, int* __restrict__ flag_bad) {
  constexpr int TM = IbpSmem<WM>::TM, NTHREADS = 128 * WM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  IbpSmem<WM>& sm = *reinterpret_cast<IbpSmem<WM>*>(smem_raw);
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * IT_N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp % WM) * 32, wn = (warp / WM) * 16;
  const int nsteps = (K + DK - 1) / DK;

  auto load_stage = [&](int stage, int step) {
    const int k0 = step * DK;
    for (int c = tid; c < DK * (TM / 2); c += NTHREADS) {
      const int kk = c / (TM / 2), ch = c % (TM / 2);
      const int m = m0 + ch * 2, k = k0 + kk;
      const int rows = (k < K) ? max(0, min(2, M - m)) : 0;
      cp16(&sm.A[stage][kk][ch * 2], rows ? A + (long long)k * lda + m : A, rows * 8);
    }
    for (int c = tid; c < 2 * IT_N * (DK / 2); c += NTHREADS) {
      const int which = c / (IT_N * (DK / 2)), cc = c % (IT_N * (DK / 2));
      const int nn = cc / (DK / 2), ch = cc % (DK / 2);
      const int n = n0 + nn, k = k0 + ch * 2;
      const int cnt = (n < N) ? max(0, min(2, K - k)) : 0;
      const double* X = which ? Xhi : Xlo;
      cp16(which ? &sm.Bhi[stage][nn][ch * 2] : &sm.Blo[stage][nn][ch * 2], cnt ? X + (long long)n * ldb + k : X, cnt * 8);
    }
  };

  double acc0[4][2][2], acc1[4][2][2];   // W c and |W| r
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc0[i][j][0] = acc0[i][j][1] = acc1[i][j][0] = acc1[i][j][1] = 0.0;

  for (int s = 0; s < DSTAGES - 1; ++s) {
    if (s < nsteps) load_stage(s, s);
    cp_commit();
  }
  for (int step = 0; step < nsteps; ++step) {
    cp_wait<DSTAGES - 2>();
    __syncthreads();
    {
      const int nxt = step + DSTAGES - 1;
      if (nxt < nsteps) load_stage(nxt % DSTAGES, nxt);
      cp_commit();
    }
    const int st = step % DSTAGES;
    const double(*As)[IbpSmem<WM>::LDA] = sm.A[st];
    const double(*Bl)[DLDB] = sm.Blo[st];
    const double(*Bh)[DLDB] = sm.Bhi[st];
#pragma unroll
    for (int k4 = 0; k4 < DK; k4 += 4) {
      const int k = k4 + (lane & 3);
      double af[4], aa[4], bc[2], br[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        af[i] = As[k][wm + i * 8 + (lane >> 2)];
        aa[i] = fabs(af[i]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const double lo = Bl[wn + j * 8 + (lane >> 2)][k], hi = Bh[wn + j * 8 + (lane >> 2)][k];
        bc[j] = 0.5 * (lo + hi);
        br[j] = 0.5 * (hi - lo);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dmma(acc0[i][j][0], acc0[i][j][1], af[i], bc[j]);
          dmma(acc1[i][j][0], acc1[i][j][1], aa[i], br[j]);
        }
    }
  }
  cp_wait<0>();

  bool bad = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + wm + i * 8 + (lane >> 2);
    if (r >= M) continue;
    const double bb = bias[r];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = n0 + wn + j * 8 + (lane & 3) * 2 + e;
        if (c >= N) continue;
        const double mid = acc0[i][j][e] + bb;
        const double ymin = mid - acc1[i][j][e], ymax = mid + acc1[i][j][e];
        bad |= !(ymin <= ymax);
        if (D0) {
          D0[(long long)c * ldd + r] = ymin;
          D1[(long long)c * ldd + r] = ymax;
        }
        if (write_x) {
          C0[(long long)c * ldc + r] = relu ? fmax(ymin, 0.0) : ymin;
          C1[(long long)c * ldc + r] = relu ? fmax(ymax, 0.0) : ymax;
        }
      }
    }
  }
  if (bad && flag_bad) atomicOr(flag_bad, 1);
}

template <int WM>
static void ibp_dmma_launch_t(const double* Mk, int n_out_k, int n_in_k, const double* xin_min, const double* xin_max,
                              long long x_stride, double* xout_min, double* xout_max, double* acx_min, double* acx_max,
                              long long acx_stride, int Q, int relu, int write_x, int* flag_bad, cudaStream_t st) {
  cudaFuncSetAttribute(ibp_dmma_kernel<WM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IbpSmem<WM>));
  const dim3 grid((n_out_k + 32 * WM - 1) / (32 * WM), (Q + IT_N - 1) / IT_N);
  ibp_dmma_kernel<WM><<<grid, 128 * WM, sizeof(IbpSmem<WM>), st>>>(Mk, n_out_k, n_out_k, n_in_k, xin_min, xin_max, x_stride, Q,
                                                                 Mk + (long long)n_in_k * n_out_k, xout_min, xout_max,
                                                                 x_stride, acx_min, acx_max, acx_stride, relu, write_x,
                                                                 flag_bad);
}

}  // namespace

// Layers b0 .. b0 + nb - 1; the caller has checked what dgemm_dmma_launch checks, for every one of them.
int dgemm_dmma_affine_layers_launch(const NetDev& nd, int b0, int nb, int max_rows, const double* u, long long u_stride,
                                    double* aff, long long aff_stride, int Q, cudaStream_t st) {
  static const bool off = [] { const char* e = getenv("NNSDP_NO_DMMA_GEMM"); return e && atoi(e) != 0; }();
  static const bool off2 = [] { const char* e = getenv("NNSDP_NO_AFFINE_LAYERS"); return e && atoi(e) != 0; }();
  if (off || off2 || nb < 2 || Q < 128) return 0;
  cudaFuncSetAttribute(dgemm_dmma_affine_layers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DgemmSmem));
  const dim3 grid((max_rows + DT - 1) / DT, (Q + DTN - 1) / DTN, nb);
  dgemm_dmma_affine_layers_kernel<<<grid, DTHREADS, sizeof(DgemmSmem), st>>>(nd, b0, u, u_stride, aff, aff_stride, Q);
  return 1;
}



// Interval propagation of one layer: Mk = [W_k b_k] (n_out x (n_in + 1), ld = n_out).  0 = operands not suitable.
int ibp_dmma_launch(const double* Mk, int n_out_k, int n_in_k, const double* xin_min, const double* xin_max,
                    long long x_stride, double* xout_min, double* xout_max, double* acx_min, double* acx_max,
                    long long acx_stride, int Q, int relu, int write_x, int* flag_bad, cudaStream_t st) {
  static const bool off = [] {
    const char* e = getenv("NNSDP_NO_DMMA_GEMM"); const char* f = getenv("NNSDP_NO_DMMA_IBP");
    return (e && atoi(e) != 0) || (f && atoi(f) != 0);
  }();
  if (off || n_out_k < 192 || Q < 128 || n_in_k < 128) return 0;
  if (((uintptr_t)Mk | (uintptr_t)xin_min | (uintptr_t)xin_max) & 15) return 0;
  if ((n_out_k & 1) || (x_stride & 1)) return 0;
  // the largest tile whose grid still covers the SMs
  const long long nt = (Q + IT_N - 1) / IT_N;
  static const int wm_env = [] { const char* e = getenv("NNSDP_IBP_WM"); return e ? atoi(e) : 0; }();
  // 64-neuron tiles run two CTAs per SM (one CTA's barrier hides under the other's DMMA stream): 1000 neurons x 1024 boxes
  // 3.94 ms per 20 layers against 4.14 ms with 128-neuron tiles and 4.50 ms with 32-neuron ones (NNSDP_IBP_WM forces one)
  const int wm = wm_env ? wm_env : (((n_out_k + 63) / 64) * nt >= 120 ? 2 : 1);
#define NNSDP_IBP(W) ibp_dmma_launch_t<W>(Mk, n_out_k, n_in_k, xin_min, xin_max, x_stride, xout_min, xout_max, acx_min, acx_max, \
                                          acx_stride, Q, relu, write_x, flag_bad, st)
  if (wm == 4) NNSDP_IBP(4);
  else if (wm == 2) NNSDP_IBP(2);
  else NNSDP_IBP(1);
#undef NNSDP_IBP
  return 1;
}

// Returns 1 when the product was launched, 0 when the operands do not fit this kernel (the caller falls back).
int dgemm_dmma_launch(const double* A, int lda, int M, int K, const double* B, long long ldb, double* C, long long ldc,
                      int N, int accumulate, cudaStream_t st) {
  static const bool off = [] { const char* e = getenv("NNSDP_NO_DMMA_GEMM"); return e && atoi(e) != 0; }();
  if (off || M < 192 || N < 128 || K < 128) return 0;   // smaller products: the 64 x 64 SIMT tiles waste less
  if (((uintptr_t)A | (uintptr_t)B) & 15) return 0;
  if ((lda & 1) || (ldb & 1)) return 0;
  cudaFuncSetAttribute(dgemm_dmma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DgemmSmem));
  cudaFuncSetAttribute(dgemm_dmma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DgemmSmem));
  const dim3 grid((M + DT - 1) / DT, (N + DTN - 1) / DTN);
  if (accumulate) dgemm_dmma_kernel<1><<<grid, DTHREADS, sizeof(DgemmSmem), st>>>(A, lda, M, K, B, ldb, C, ldc, N);
  else dgemm_dmma_kernel<0><<<grid, DTHREADS, sizeof(DgemmSmem), st>>>(A, lda, M, K, B, ldb, C, ldc, N);
  return 1;
}

}  // namespace nnsdp
