// K3: Gram contractions  G_b = W_b' diag(d11) W_b  restricted to the Gram-active neurons
// (d11 = -2 smin smax lambda != 0, i.e. stably-active ReLUs with a non-zero multiplier).
// This is the A' Q11 A part of  Zac = R' Q R  (reference: src/Qc/activ.jl:40 with
// Q11 from src/Qc/activ_sector.jl:42).
//
// FP64 tensor cores: tcgen05 has no f64 kind, so the contraction uses the DMMA path
// (mma.sync.aligned.m8n8k4.f64).  Operands come from the input-contiguous copy Wt of W, so a
// gathered neuron j is one contiguous column (cp.async 16 B chunks, fully coalesced whatever
// the active set looks like).  Only upper-triangular 128x128 tile pairs are computed; every tile
// is written to G[r,c] and mirrored to G[c,r] from the same accumulators, so G is bit-symmetric.
#include "internal.h"
#include <stdlib.h>

namespace nnsdp {

namespace {

// Tile side GT is 128, or 64 when the work list is too short to fill the GPU with 128 x 128 tiles (a single query
// on a wide net: 36 tiles of one active layer on 148 SMs).  Every entry accumulates its neurons in the same order
// (four per DMMA, stage after stage) whatever the tile side, so the result does not depend on the choice.
constexpr int GK = 16;         // neurons per pipeline stage
constexpr int GSTAGES = 3;
#ifndef NNSDP_GRAM_WARPS_N
#define NNSDP_GRAM_WARPS_N 4
#endif
#ifndef NNSDP_GRAM_LD_PAD
#define NNSDP_GRAM_LD_PAD 4
#endif
// 128 x 128 tiles: 16 warps, 4 (rows) x 4 (cols) of 32 x 32 (the DMMA cadence of one warp leaves the pipe half idle with
// two warps per scheduler); 64 x 64 tiles: 8 warps, 4 x 2 of 16 x 32
template <int GT>
struct GramCfg {
  static constexpr int WARPS_N = GT >= 128 ? NNSDP_GRAM_WARPS_N : 2;
  static constexpr int THREADS = 4 * WARPS_N * 32;
};

template <int GT>
struct GramSmem {
  // smem leading dimension: 4 (mod 16) puts the 4 x 8 (k, row) addresses of a fragment load in distinct 8-byte banks
  // per half warp (GT + 8 gave two-way conflicts: k and k + 2 met in the same banks)
  static constexpr int GLD = GT + NNSDP_GRAM_LD_PAD;
  double A[GSTAGES][GK][GLD];
  double B[GSTAGES][GK][GLD];
  double d[GSTAGES][GK];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;  // src-size 0 zero-fills the 16 bytes
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int GT>
__global__ void __launch_bounds__(GramCfg<GT>::THREADS, 1)
gram_kernel(NetDev net, BatchDev b, GramDev g, int q0, const int* __restrict__ pairs) {
  constexpr int GTHREADS = GramCfg<GT>::THREADS;
  constexpr int GLD = GramSmem<GT>::GLD, WM = GT / 4, WN = GT / GramCfg<GT>::WARPS_N, NI = WM / 8, NJ = WN / 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GramSmem<GT>& sm = *reinterpret_cast<GramSmem<GT>*>(smem_raw);

  // work list: only (query, block) pairs with active neurons; Gram of block blk uses layer matrix M[blk]
  // (pairs on grid.x: up to 2^31 - 1; tile pairs on grid.y)
  const int q = pairs[2 * blockIdx.x], blk = pairs[2 * blockIdx.x + 1], slot = q - q0;
  const int cnt = b.cnt[(long long)q * net.K + blk];
  if (cnt == 0) return;
  const int nb = net.n[blk];
  const int ntile = (nb + GT - 1) / GT;
  // decode the upper-triangular tile pair (ti <= tj) from blockIdx.y
  int ti = 0, rem = blockIdx.y;
  while (ti < ntile && rem >= ntile - ti) {
    rem -= ntile - ti;
    ++ti;
  }
  if (ti >= ntile) return;
  const int tj = ti + rem;
  const bool diag = (ti == tj);
  const int m0 = ti * GT, n0 = tj * GT;

  const double* Wt = net.Wt[blk];
  const int ldT = net.ldT[blk];
  const int L0 = net.off[blk + 1] - net.n_in;
  const int* act = b.act + (long long)q * net.acdim + L0;
  const double* d11 = b.d11 + (long long)q * net.acdim + L0;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 3) * WM, wn = (warp >> 2) * WN;
  const int nsteps = (cnt + GK - 1) / GK;

  auto load_stage = [&](int stage, int step) {
    const int k0 = step * GK;
    // 16 neurons x GT rows = 16 x GT/2 chunks of 16 B per operand
    for (int c = tid; c < GK * (GT / 2); c += GTHREADS) {
      const int kk = c / (GT / 2), ch = c % (GT / 2);
      const bool valid = (k0 + kk) < cnt;
      const int jl = valid ? act[k0 + kk] : 0;
      const double* col = Wt + (long long)jl * ldT;
      cp_async16(&sm.A[stage][kk][ch * 2], col + m0 + ch * 2, valid);
      if (!diag) cp_async16(&sm.B[stage][kk][ch * 2], col + n0 + ch * 2, valid);
    }
    if (tid < GK) sm.d[stage][tid] = (k0 + tid < cnt) ? d11[act[k0 + tid]] : 0.0;
  };

  double acc[NI][NJ][2];
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int s = 0; s < GSTAGES - 1; ++s) {
    if (s < nsteps) load_stage(s, s);
    cp_async_commit();
  }
  for (int step = 0; step < nsteps; ++step) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();
    {  // prefetch into the stage consumed in the previous iteration
      const int nxt = step + GSTAGES - 1;
      if (nxt < nsteps) load_stage(nxt % GSTAGES, nxt);
      cp_async_commit();
    }
    const int st = step % GSTAGES;
    const double(*As)[GLD] = sm.A[st];
    const double(*Bs)[GLD] = diag ? sm.A[st] : sm.B[st];
#pragma unroll
    for (int k4 = 0; k4 < GK; k4 += 4) {
      const int k = k4 + (lane & 3);
      const double dk = sm.d[st][k];
      double af[NI], bf[NJ];
#pragma unroll
      for (int i = 0; i < NI; ++i) af[i] = As[k][wm + i * 8 + (lane >> 2)] * dk;
#pragma unroll
      for (int j = 0; j < NJ; ++j) bf[j] = Bs[k][wn + j * 8 + (lane >> 2)];
#pragma unroll
      for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // epilogue: G[r, c] and the mirror G[c, r]; on diagonal tiles only r <= c is authoritative.
  double* G = g.scratch + (long long)slot * g.per_query + g.goff[blk];
  const int ldG = g.ldG[blk];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int r = m0 + wm + i * 8 + (lane >> 2);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = n0 + wn + j * 8 + (lane & 3) * 2;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = c + e;
        if (r < nb && cc < nb && (!diag || r <= cc)) {
          const double v = acc[i][j][e];
          G[r + (long long)cc * ldG] = v;
          if (r != cc && g.mirror) G[cc + (long long)r * ldG] = v;
        }
      }
    }
  }
}

}  // namespace

template <int GT>
static void launch_gram_t(const NetDev& net, const BatchDev& b, const GramDev& g, int max_n, int q0, const int* pairs,
                          int npairs, cudaStream_t st) {
  cudaFuncSetAttribute(gram_kernel<GT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GramSmem<GT>));
  const int ntile = (max_n + GT - 1) / GT;
  dim3 grid(npairs, ntile * (ntile + 1) / 2);  // grid.y <= 65535 is checked by launch_gram
  gram_kernel<GT><<<grid, GramCfg<GT>::THREADS, sizeof(GramSmem<GT>), st>>>(net, b, g, q0, pairs);
}

int launch_gram(const NetDev& net, const BatchDev& b, const GramDev& g, int max_n, int q0, const int* pairs,
                int npairs, cudaStream_t st) {
  if (net.K < 2 || npairs <= 0) return 0;
  {  // tile pairs of the widest block must fit grid.y (widths up to ~23,000 with 64 x 64 tiles, ~46,000 with 128 x 128)
    const long long nt = (max_n + 127) / 128;
    if (nt * (nt + 1) / 2 > 65535) return -1;
  }
  static const int force = [] { const char* e = getenv("NNSDP_GRAM_TILE"); return e ? atoi(e) : 0; }();
  const int nt128 = (max_n + 127) / 128;
  const long long items128 = (long long)nt128 * (nt128 + 1) / 2 * npairs;
  const int nt64 = (max_n + 63) / 64;
  const long long pad128 = (long long)nt128 * (nt128 + 1) / 2 * 4, pad64 = (long long)nt64 * (nt64 + 1) / 2;
  // 64 x 64 tiles when there are fewer 128 x 128 tiles than SMs, or when they would be mostly padding (width 100:
  // one 128 tile against three 64 tiles)
  const bool small = force ? force == 64 : (items128 < 148 || pad64 * 100 <= pad128 * 85);
  if (small) launch_gram_t<64>(net, b, g, max_n, q0, pairs, npairs, st);
  else launch_gram_t<128>(net, b, g, max_n, q0, pairs, npairs, st);
  return 1;
}

}  // namespace nnsdp
