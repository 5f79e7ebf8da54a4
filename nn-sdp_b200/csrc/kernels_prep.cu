// K2b: per-query preparation of everything the block emitter needs that is O(Zdim):
// QC diagonals, band multipliers T, the affine column Z[:, a], the small Zin/Zout blocks and
// the compacted sets of Gram-active neurons.  HBM-streaming FP64, one thread per neuron.
//
// Reference semantics (numeric gamma):
//   makeQ(QcActivSector)   src/Qc/activ_sector.jl:23-60   (lambda / v / eta / nu layout :26,34,51-54)
//   makeQ(QcActivBounded)  src/Qc/activ_bounded.jl:13-23
//   makeZac = R' Q R       src/Qc/activ.jl:30-41
//   makeZin (box)          src/Qc/input.jl:22-26,36-40
//   makeZout               src/Qc/output.jl:52-106
// The closed form these kernels evaluate is derived in DESIGN.md ("closed form of Z").
#include "internal.h"

namespace nnsdp {

namespace {

// number of pairs (i', j) with i' < i in the reference's ordering (i ascending, j ascending,
// j - i <= beta, j < acdim): sum_{i' < i} min(beta, acdim-1-i').   activ_sector.jl:29
__device__ __forceinline__ long long pair_base(long long i, long long acdim, long long beta) {
  long long m = acdim - beta;
  if (m < 0) m = 0;
  if (i <= m) return i * beta;
  return m * beta + (i - m) * (acdim - 1) - ((m + i - 1) * (i - m)) / 2;
}

__global__ void __launch_bounds__(PREP_THREADS)
prep_neuron_kernel(NetDev net, BatchDev b, int* __restrict__ err_flag, int q_base) {
  const int q = q_base + blockIdx.y;
  const long long acdim = net.acdim, beta = b.beta;
  const long long j = (long long)blockIdx.x * PREP_THREADS + threadIdx.x;
  double part = 0.0;
  if (j < acdim) {
    const double* gsec = b.gsec + q * b.s_gsec;
    const double sm = b.smin[q * b.s_smin + j], sx = b.smax[q * b.s_smax + j];
    const double ymin = b.ymin[q * b.s_ymin + j], ymax = b.ymax[q * b.s_ymax + j];
    const double gb = b.gbnd[q * b.s_gbnd + j];
    // @assert's of QcActivBounded (activ_bounded.jl:8) and QcActivSector (activ_sector.jl:14-16)
    if (!(ymin <= ymax)) atomicOr(err_flag, 1);
    if (!(sm <= sx) || !(0.0 <= sm) || !(sx <= 1.0)) atomicOr(err_flag, 2);
    const double lam = gsec[j];
    const double eta = gsec[b.lamdim + j], nu = gsec[b.lamdim + acdim + j];
    const double* v = gsec + acdim;
    const double* bias = net.bias_all;
    const double bj = bias[j];
    const double p = sm * sx, qq = sm + sx;
    const double d11 = -2.0 * (p * lam);  // _Q11 diagonal, activ_sector.jl:42 (base_smin*base_smax = 0)
    double t0 = 0.0;   // T[j, j]
    double sb = 0.0;   // sum_{j' != j} b_j' T[j', j]
    const long long base_j = pair_base(j, acdim, beta);
    double* Bt = b.Bt + (long long)q * beta * acdim;
    for (long long t = 1; t <= beta; ++t) {
      double tv = 0.0;
      if (j + t < acdim) {
        const double vv = v[base_j + t - 1];
        t0 += vv;
        tv = -vv;
        sb = fma(bias[j + t], tv, sb);
      }
      Bt[(t - 1) * acdim + j] = tv;
    }
    for (long long t = 1; t <= beta; ++t) {
      if (j - t >= 0) {
        const double vv = v[pair_base(j - t, acdim, beta) + t - 1];
        t0 += vv;
        sb = fma(bias[j - t], -vv, sb);
      }
    }
    const double md = qq * lam + t0;  // _Q12 diagonal, activ_sector.jl:43 (base_smin+base_smax = 1)
    const double c13 = -sm * eta - sx * nu;  // activ_sector.jl:55
    const double c23 = eta + nu;             // activ_sector.jl:56
    const long long o = (long long)q * acdim + j;
    b.d11[o] = d11;
    b.T0[o] = t0;
    b.Md[o] = md;
    b.u[o] = fma(d11, bj, c13);
    // Z[eps_j, a] = gamma_bnd (ymin + ymax) + (eta + nu) + sum_j' b_j' M[j', j]
    b.aff[(long long)q * net.Zdim + net.n_in + j] = gb * (ymin + ymax) + c23 + fma(bj, md, sb);
    // contribution to Z[a, a]
    part = -2.0 * (gb * ymin * ymax) + d11 * bj * bj + 2.0 * (c13 * bj);
  }
  // deterministic CTA reduction (fixed tree), one partial per CTA
  __shared__ double red[PREP_THREADS];
  red[threadIdx.x] = part;
  __syncthreads();
  for (int s = PREP_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) b.part[(long long)q * b.npart + blockIdx.x] = red[0];
}

// One CTA per query: Zin/Zout pieces, Z[a,a], x_1 / x_K rows of the affine column, active sets.
__global__ void __launch_bounds__(PREP_THREADS)
prep_final_kernel(NetDev net, BatchDev b) {
  extern __shared__ double sh[];
  const int q = blockIdx.x, tid = threadIdx.x;
  const int K = net.K, n_in = net.n_in, n_out = net.n_out, sdim = b.sdim;
  const int nK = net.n[K - 1];
  double* S = sh;                 // sdim * sdim, column-major, full symmetric
  double* t2 = S + sdim * sdim;   // n_out : S22 b_K + S23

  // ---- S of the output QC (src/Qc/output.jl:52-98) ---------------------------------
  for (int i = tid; i < sdim * sdim; i += PREP_THREADS) S[i] = 0.0;
  __syncthreads();
  const int i2 = n_in, i3 = n_in + n_out;
  if (b.out_kind == NNSDP_OUT_SAFETY) {
    const double* Sq = b.outS + q * b.s_outS;
    for (int i = tid; i < sdim * sdim; i += PREP_THREADS) {
      const int r = i % sdim, c = i / sdim;  // Symmetric(S): the upper triangle is authoritative
      S[i] = (r <= c) ? Sq[r + c * sdim] : Sq[c + r * sdim];
    }
  } else {
    const double* vec = b.outvec + q * b.s_outvec;
    const double gout = b.gout[q * b.s_gout];
    if (b.out_kind == NNSDP_OUT_HPLANE) {
      for (int o = tid; o < n_out; o += PREP_THREADS) {
        S[(i2 + o) + i3 * sdim] = vec[o];  // _S23 = normal
        S[i3 + (i2 + o) * sdim] = vec[o];
      }
      if (tid == 0) S[i3 + i3 * sdim] = -2.0 * gout;  // _S33
    } else {
      const double* invP = b.outinvP + q * b.s_outinvP;  // n_out x n_out column-major
      const bool ell = (b.out_kind == NNSDP_OUT_ELLIPSOID);
      for (int i = tid; i < n_out * n_out; i += PREP_THREADS) {
        const int r = i % n_out, c = i / n_out;
        double s22 = (r == c) ? 1.0 : 0.0;  // circle: I
        if (ell) {                         // invP' * invP
          s22 = 0.0;
          for (int m = 0; m < n_out; ++m) s22 = fma(invP[m + r * n_out], invP[m + c * n_out], s22);
        }
        S[(i2 + r) + (i2 + c) * sdim] = s22;
      }
      for (int o = tid; o < n_out; o += PREP_THREADS) {
        double s23 = -vec[o];  // circle: -yc
        if (ell) {             // -invP' * yc
          s23 = 0.0;
          for (int m = 0; m < n_out; ++m) s23 = fma(invP[m + o * n_out], vec[m], s23);
          s23 = -s23;
        }
        S[(i2 + o) + i3 * sdim] = s23;
        S[i3 + (i2 + o) * sdim] = s23;
      }
      if (tid == 0) {
        double yy = 0.0;
        for (int m = 0; m < n_out; ++m) yy = fma(vec[m], vec[m], yy);
        S[i3 + i3 * sdim] = yy - gout;  // yc'yc - gamma_out (output.jl:84,93)
      }
    }
  }
  __syncthreads();

  const double* MK = net.M[K - 1];            // n_out x (nK + 1)
  const double* bK = MK + (long long)nK * n_out;
  for (int o = tid; o < n_out; o += PREP_THREADS) {
    double s = S[(i2 + o) + i3 * sdim];
    for (int m = 0; m < n_out; ++m) s = fma(S[(i2 + o) + (i2 + m) * sdim], bK[m], s);
    t2[o] = s;
  }
  __syncthreads();

  const double* gin = b.gin + q * b.s_gin;
  const double* x1min = b.x1min + q * b.s_x1min;
  const double* x1max = b.x1max + q * b.s_x1max;
  double* aff = b.aff + (long long)q * net.Zdim;

  // Z11 = S11 - 2 diag(gamma_in)   (input.jl:24, output.jl Eout'R'SR Eout restricted to x_1)
  double* Z11 = b.Z11 + (long long)q * n_in * n_in;
  for (int i = tid; i < n_in * n_in; i += PREP_THREADS) {
    const int r = i % n_in, c = i / n_in;
    Z11[i] = S[r + c * sdim] + (r == c ? -2.0 * gin[r] : 0.0);
  }
  // x_1 rows of the affine column: gamma_in (xmin + xmax) + S12 b_K + S13
  for (int r = tid; r < n_in; r += PREP_THREADS) {
    double s = S[r + i3 * sdim];
    for (int m = 0; m < n_out; ++m) s = fma(S[r + (i2 + m) * sdim], bK[m], s);
    aff[r] = gin[r] * (x1min[r] + x1max[r]) + s;
  }
  // x_K rows: += W_K' (S22 b_K + S23);  Z1K = S12 W_K;  U = S22 W_K
  double* Z1K = b.Z1K + (long long)q * n_in * nK;
  double* U = b.U + (long long)q * n_out * nK;
  for (int c = tid; c < nK; c += PREP_THREADS) {
    const double* wc = MK + (long long)c * n_out;
    double s = 0.0;
    for (int m = 0; m < n_out; ++m) s = fma(wc[m], t2[m], s);
    aff[net.off[K - 1] + c] += s;
    if (b.has_s12)
      for (int r = 0; r < n_in; ++r) {
        double z = 0.0;
        for (int m = 0; m < n_out; ++m) z = fma(S[r + (i2 + m) * sdim], wc[m], z);
        Z1K[r + (long long)c * n_in] = z;
      }
    if (b.has_s22)
      for (int o = 0; o < n_out; ++o) {
        double z = 0.0;
        for (int m = 0; m < n_out; ++m) z = fma(S[(i2 + o) + (i2 + m) * sdim], wc[m], z);
        U[o + (long long)c * n_out] = z;
      }
  }
  // Z[a, a]
  if (tid == 0) {
    double s = 0.0;
    const double* part = b.part + (long long)q * b.npart;
    for (int i = 0; i < b.npart; ++i) s += part[i];
    double zin = 0.0;
    for (int r = 0; r < n_in; ++r) zin = fma(gin[r] * x1min[r], x1max[r], zin);
    double so = S[i3 + i3 * sdim];
    for (int m = 0; m < n_out; ++m) so = fma(bK[m], t2[m] + S[(i2 + m) + i3 * sdim], so);
    // b_K' S22 b_K + 2 b_K' S23 + S33 = b_K' (t2 + S23) + S33
    aff[net.Zdim - 1] = s - 2.0 * zin + so;
  }

}

// Ordered compaction of the Gram-active neurons (d11 != 0) of one layer of one query: grid (K - 1, Q).
__global__ void __launch_bounds__(PREP_THREADS)
prep_compact_kernel(NetDev net, BatchDev b, int q_base) {
  __shared__ int wsum[PREP_THREADS / 32];
  __shared__ int base_sh;
  const int blk = blockIdx.x, q = q_base + blockIdx.y, tid = threadIdx.x;
  const int K = net.K, n_in = net.n_in;
  const double* d11 = b.d11 + (long long)q * net.acdim;
  int* act = b.act + (long long)q * net.acdim;
  const int lane = tid & 31, wid = tid >> 5;
  const int L0 = net.off[blk + 1] - n_in, nl = net.n[blk + 1];
  if (tid == 0) base_sh = 0;
  __syncthreads();
  for (int i0 = 0; i0 < nl; i0 += PREP_THREADS) {
    const int i = i0 + tid;
    const bool on = (i < nl) && (d11[L0 + i] != 0.0);
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (lane == 0) wsum[wid] = __popc(m);
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += wsum[w];
    const int base = base_sh;
    if (on) act[L0 + base + woff + __popc(m & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < PREP_THREADS / 32; ++w) tot += wsum[w];
      base_sh = base + tot;
    }
    __syncthreads();
  }
  if (tid == 0) b.cnt[(long long)q * K + blk] = base_sh;
}

}  // namespace

int launch_prep(const NetDev& net, const BatchDev& b, int* err_flag, cudaStream_t st) {
  // queries ride on grid.y (<= 65535): larger batches are launched in slices
  int launches = 0;
  for (int q_base = 0; q_base < b.Q; q_base += 65535) {
    const int nq = min(65535, b.Q - q_base);
    prep_neuron_kernel<<<dim3(b.npart, nq), PREP_THREADS, 0, st>>>(net, b, err_flag, q_base);
    ++launches;
  }
  const size_t shbytes = (size_t)(b.sdim * b.sdim + net.n_out) * sizeof(double);
  if (shbytes > 40000)  // beyond the default 48 KB of dynamic shared memory (the host checks sdim <= 160)
    cudaFuncSetAttribute(prep_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shbytes);
  prep_final_kernel<<<b.Q, PREP_THREADS, shbytes, st>>>(net, b);
  ++launches;
  if (net.K >= 2)
    for (int q_base = 0; q_base < b.Q; q_base += 65535) {
      prep_compact_kernel<<<dim3(net.K - 1, min(65535, b.Q - q_base)), PREP_THREADS, 0, st>>>(net, b, q_base);
      ++launches;
    }
  return launches;
}

}  // namespace nnsdp
