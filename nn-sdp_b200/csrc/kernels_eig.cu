// Certificate check: lambda_max of Z(gamma) without ever forming Z.
// The reference calls eigmax(Symmetric(Matrix(Z))) after every solve (src/Methods/Methods.jl:116-117) and
// accepts non-OPTIMAL ACAS runs on eigmax(Z) <= 1e-4 (experiments/acas.jl:71-79); at Zdim = 20,003 that is
// a 3.2 GB dense eigenproblem.  Here Z x is evaluated from the factored form  Z = R' Q R + Zin + Zout
// (src/Qc/activ.jl:30-41):  with t = A x~ (one GEMM per layer over the batch of queries), u = B x~,
//   s1 = d11 o t + M u,     s2 = M t + (-2 T - 2 D_bnd) u,     M = diag(q lambda) + T   (band, half-width beta)
//   (Z x)~ = A' s1 + B' s2 + (Zin + Zout) x~ + aff x_a,         (Z x)_a = aff . x
// and a batched Lanczos iteration with full re-orthogonalisation runs on top of it.
#include <algorithm>

#include "internal.h"

namespace nnsdp {

namespace {

constexpr int EIG_THREADS = 256;

__device__ __forceinline__ double block_reduce_sum(double v, double* sh) {
  // fixed tree: deterministic
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = EIG_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// start vector: fixed pseudo-random signs and magnitudes (splitmix64 of (query, index))
__global__ void eig_init_kernel(double* __restrict__ v, int Zdim, int q_first) {
  const int q = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Zdim; i += gridDim.x * blockDim.x) {
    unsigned long long z = (unsigned long long)(q_first + q) * 0x9E3779B97F4A7C15ull + (unsigned long long)i * 0xBF58476D1CE4E5B9ull + 0x94D049BB133111EBull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    v[(long long)q * Zdim + i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  }
}

// per neuron: s1 and the neuron rows of y.  x, y: Zdim per query; t, s1: acdim per query.
__global__ void __launch_bounds__(EIG_THREADS)
eig_mid_kernel(NetDev net, BatchDev b, int q_first, const double* __restrict__ x, const double* __restrict__ t,
               double* __restrict__ s1, double* __restrict__ y) {
  const int ql = blockIdx.y, q = q_first + ql;
  const long long acdim = net.acdim;
  const int j = blockIdx.x * EIG_THREADS + threadIdx.x;
  if (j >= acdim) return;
  const int beta = b.beta, n0 = net.n_in, a = net.Zdim - 1;
  const double* xq = x + (long long)ql * net.Zdim;
  const double* u = xq + n0;
  const double* tq = t + (long long)ql * acdim;
  const double* Md = b.Md + (long long)q * acdim;
  const double* T0 = b.T0 + (long long)q * acdim;
  const double* Bt = b.Bt + (long long)q * beta * acdim;
  const double* d11 = b.d11 + (long long)q * acdim;
  const double gb = b.gbnd[q * b.s_gbnd + j];
  double mu = Md[j] * u[j], mt = Md[j] * tq[j], tu = T0[j] * u[j];
  for (int k = 1; k <= beta; ++k) {
    if (j + k < acdim) {
      const double c = Bt[(long long)(k - 1) * acdim + j];  // T[j, j+k]
      mu = fma(c, u[j + k], mu);
      mt = fma(c, tq[j + k], mt);
      tu = fma(c, u[j + k], tu);
    }
    if (j - k >= 0) {
      const double c = Bt[(long long)(k - 1) * acdim + (j - k)];  // T[j-k, j]
      mu = fma(c, u[j - k], mu);
      mt = fma(c, tq[j - k], mt);
      tu = fma(c, u[j - k], tu);
    }
  }
  s1[(long long)ql * acdim + j] = fma(d11[j], tq[j], mu);
  const double* aff = b.aff + (long long)q * net.Zdim;
  y[(long long)ql * net.Zdim + n0 + j] = mt - 2.0 * tu - 2.0 * gb * u[j] + aff[n0 + j] * xq[a];
}

// per query: x_1 rows (=), x_K rows (+=) from Zin / Zout, and y_a = aff . x
__global__ void __launch_bounds__(EIG_THREADS)
eig_io_kernel(NetDev net, BatchDev b, int q_first, const double* __restrict__ x, double* __restrict__ y) {
  extern __shared__ double shm[];  // n_out (U x_K) + reduction scratch
  double* Ux = shm;
  double* red = shm + net.n_out;
  const int ql = blockIdx.x, q = q_first + ql, tid = threadIdx.x;
  const int K = net.K, n0 = net.n_in, n_out = net.n_out, nK = net.n[K - 1], oK = net.off[K - 1], a = net.Zdim - 1;
  const double* xq = x + (long long)ql * net.Zdim;
  double* yq = y + (long long)ql * net.Zdim;
  const double* aff = b.aff + (long long)q * net.Zdim;
  const double* Z11 = b.Z11 + (long long)q * n0 * n0;
  const double* Z1K = b.Z1K + (long long)q * n0 * nK;
  const double* U = b.U + (long long)q * n_out * nK;
  const double* WK = net.M[K - 1];
  const double xa = xq[a];
  // U x_K
  for (int o = 0; o < n_out; ++o) {
    double p = 0.0;
    if (b.has_s22)
      for (int c = tid; c < nK; c += EIG_THREADS) p = fma(U[o + (long long)c * n_out], xq[oK + c], p);
    const double tot = block_reduce_sum(p, red);
    if (tid == 0) Ux[o] = tot;
  }
  __syncthreads();
  // x_1 rows
  for (int r = 0; r < n0; ++r) {
    double p = 0.0;
    if (b.has_s12)
      for (int c = tid; c < nK; c += EIG_THREADS) p = fma(Z1K[r + (long long)c * n0], xq[oK + c], p);
    const double tot = block_reduce_sum(p, red);
    if (tid == 0) {
      double v = tot + aff[r] * xa;
      for (int c = 0; c < n0; ++c) v = fma(Z11[r + c * n0], xq[c], v);
      yq[r] = v;
    }
  }
  // x_K rows: += Z1K' x_1 + W_K' (U x_K)
  for (int c = tid; c < nK; c += EIG_THREADS) {
    double v = 0.0;
    if (b.has_s12)
      for (int r = 0; r < n0; ++r) v = fma(Z1K[r + (long long)c * n0], xq[r], v);
    if (b.has_s22)
      for (int o = 0; o < n_out; ++o) v = fma(WK[o + (long long)c * n_out], Ux[o], v);
    yq[oK + c] += v;
  }
  // y_a
  double p = 0.0;
  for (int i = tid; i <= a; i += EIG_THREADS) p = fma(aff[i], xq[i], p);
  const double tot = block_reduce_sum(p, red);
  if (tid == 0) yq[a] = tot;
}

// c[q][i] = V_i[q] . w[q], i = 0..nv-1.   V: [nv_cap][Qc][n]
__global__ void __launch_bounds__(EIG_THREADS)
eig_multidot_kernel(const double* __restrict__ V, const double* __restrict__ w, int n, int Qc,
                    double* __restrict__ c, int c_ld) {
  __shared__ double red[EIG_THREADS];
  const int i = blockIdx.x, q = blockIdx.y;
  const double* v = V + ((long long)i * Qc + q) * n;
  const double* wq = w + (long long)q * n;
  double p = 0.0;
  for (int k = threadIdx.x; k < n; k += EIG_THREADS) p = fma(v[k], wq[k], p);
  const double tot = block_reduce_sum(p, red);
  if (threadIdx.x == 0) c[(long long)q * c_ld + i] = tot;
}

// w[q] -= sum_i c[q][i] V_i[q]
__global__ void __launch_bounds__(EIG_THREADS)
eig_project_kernel(const double* __restrict__ V, double* __restrict__ w, int n, int Qc, int nv,
                   const double* __restrict__ c, int c_ld) {
  const int q = blockIdx.y;
  const int k = blockIdx.x * EIG_THREADS + threadIdx.x;
  if (k >= n) return;
  double acc = w[(long long)q * n + k];
  for (int i = 0; i < nv; ++i) acc = fma(-c[(long long)q * c_ld + i], V[((long long)i * Qc + q) * n + k], acc);
  w[(long long)q * n + k] = acc;
}

// nrm[q] = ||w[q]||;  dst[q] = w[q] / nrm[q]  (zero vector when the norm vanishes: invariant subspace)
__global__ void __launch_bounds__(EIG_THREADS)
eig_normalize_kernel(const double* __restrict__ w, double* __restrict__ dst, int n, double* __restrict__ nrm) {
  __shared__ double red[EIG_THREADS];
  const int q = blockIdx.x;
  const double* wq = w + (long long)q * n;
  double p = 0.0;
  for (int k = threadIdx.x; k < n; k += EIG_THREADS) p = fma(wq[k], wq[k], p);
  const double nn = sqrt(block_reduce_sum(p, red));
  if (threadIdx.x == 0) nrm[q] = nn;
  const double inv = nn > 0.0 ? 1.0 / nn : 0.0;
  for (int k = threadIdx.x; k < n; k += EIG_THREADS) dst[(long long)q * n + k] = wq[k] * inv;
}

}  // namespace

int launch_eig_init(double* v, int Zdim, int Qc, int q_first, cudaStream_t st) {
  eig_init_kernel<<<dim3(std::min(64, (Zdim + 255) / 256), Qc), 256, 0, st>>>(v, Zdim, q_first);
  return 1;
}

int launch_eig_mid(const NetDev& net, const BatchDev& b, int q_first, int Qc, const double* x,
                   const double* t, double* s1, double* y, cudaStream_t st) {
  eig_mid_kernel<<<dim3((net.acdim + EIG_THREADS - 1) / EIG_THREADS, Qc), EIG_THREADS, 0, st>>>(net, b, q_first, x, t, s1, y);
  return 1;
}

int launch_eig_io(const NetDev& net, const BatchDev& b, int q_first, int Qc, const double* x, double* y,
                  cudaStream_t st) {
  const size_t sh = (size_t)(net.n_out + EIG_THREADS) * sizeof(double);
  eig_io_kernel<<<Qc, EIG_THREADS, sh, st>>>(net, b, q_first, x, y);
  return 1;
}

int launch_eig_multidot(const double* V, const double* w, int n, int Qc, int nv, double* c, int c_ld,
                        cudaStream_t st) {
  if (nv <= 0) return 0;
  eig_multidot_kernel<<<dim3(nv, Qc), EIG_THREADS, 0, st>>>(V, w, n, Qc, c, c_ld);
  return 1;
}

int launch_eig_project(const double* V, double* w, int n, int Qc, int nv, const double* c, int c_ld,
                       cudaStream_t st) {
  if (nv <= 0) return 0;
  eig_project_kernel<<<dim3((n + EIG_THREADS - 1) / EIG_THREADS, Qc), EIG_THREADS, 0, st>>>(V, w, n, Qc, nv, c, c_ld);
  return 1;
}

int launch_eig_normalize(const double* w, double* dst, int n, int Qc, double* nrm, cudaStream_t st) {
  eig_normalize_kernel<<<Qc, EIG_THREADS, 0, st>>>(w, dst, n, nrm);
  return 1;
}

}  // namespace nnsdp
