// Affine-coefficient ("symbolic") mode: Z(gamma) = Z0 + sum_v gamma_v Z_v as COO triplets over the upper
// triangle of the clique cover, for the hand-off that replaces the reference's AffExpr algebra
//   Z = Zin + Zout + sum(Zacs);  @constraint(model, Z .== Zksum)      src/Methods/chordal_sdp.jl:96-153
// (SURVEY.md section 8f-1).  Variables are stacked in the reference's creation order
//   [gamma_in (n_in); gamma_out (reach kinds: 1); gamma_ac1 = bounded (acdim); gamma_ac2 = sector (secdim)]
// and the coefficient patterns follow the closed form of DESIGN.md section 2:
//   gamma_in_i : -2 e_i e_i' + (xmin_i + xmax_i)(e_i e_a' + e_a e_i') - 2 xmin_i xmax_i e_a e_a'          input.jl:22-26
//   gamma_bnd_j: -2 eps eps' + (ymin_j + ymax_j)(eps e_a' + e_a eps') - 2 ymin_j ymax_j e_a e_a'          activ_bounded.jl:19-21
//   lambda_j   : -2 p_j rho rho' + q_j (rho eps' + eps rho')                                               activ_sector.jl:42-43
//   v_ij       : (rho_i - rho_j)(eps_i - eps_j)' + transpose - 2 (eps_i - eps_j)(eps_i - eps_j)'           activ_sector.jl:29-35,43,45
//   eta_j, nu_j: -s_j (rho e_a' + e_a rho') + (eps e_a' + e_a eps'),  s = smin | smax                      activ_sector.jl:55-56
//   gamma_out  : -2 e_a e_a' (hyperplane) | -e_a e_a' (circle, ellipsoid)                                  output.jl:76,84,93
// with rho_j = row j of [W b] placed in the block that feeds neuron j (bias at the affine index a) and
// eps_j the neuron's own index.  One CTA per variable; duplicate (entry, variable) pairs are allowed
// (consumers sum them, like Julia's sparse()).  All indices are 0-based here; the host adds 1.
#include "internal.h"

namespace nnsdp {

namespace {

constexpr int AFF_THREADS = 256;

struct VarInfo {
  int kind;       // 0 gin, 1 gout, 2 gbnd, 3 lambda, 4 v pair, 5 eta, 6 nu
  long long j;    // neuron (or input index); for pairs the smaller neuron i
  int t;          // pair distance j2 - i
};

__device__ __forceinline__ long long pair_base_dev(long long i, long long acdim, long long beta) {
  long long m = acdim - beta;
  if (m < 0) m = 0;
  if (i <= m) return i * beta;
  return m * beta + (i - m) * (acdim - 1) - ((m + i - 1) * (i - m)) / 2;
}

// variable index -> what it is.  Sector layout: [lambda(acdim); v(pairs); eta(acdim); nu(acdim)].
__device__ VarInfo decode_var(const AffineDev& A, long long v) {
  VarInfo r{0, 0, 0};
  if (v < A.var_out) { r.kind = 0; r.j = v; return r; }
  if (v < A.var_bnd) { r.kind = 1; return r; }
  if (v < A.var_sec) { r.kind = 2; r.j = v - A.var_bnd; return r; }
  long long s = v - A.var_sec;
  if (s < A.acdim) { r.kind = 3; r.j = s; return r; }
  if (s < A.lamdim) {
    // pair index -> (i, t): pairs of i start at pair_base(i); binary search on i
    const long long pidx = s - A.acdim;
    long long lo = 0, hi = A.acdim - 1;
    while (lo < hi) {
      const long long mid = (lo + hi + 1) >> 1;
      if (pair_base_dev(mid, A.acdim, A.beta) <= pidx) lo = mid; else hi = mid - 1;
    }
    r.kind = 4; r.j = lo; r.t = (int)(pidx - pair_base_dev(lo, A.acdim, A.beta)) + 1;
    return r;
  }
  s -= A.lamdim;
  if (s < A.acdim) { r.kind = 5; r.j = s; return r; }
  r.kind = 6; r.j = s - A.acdim;
  return r;
}

// block (0-based, >= 1) that holds neuron j, i.e. eps_j = n_in + j lies in block blk_of[n_in + j]
__device__ __forceinline__ int nblk(const NetDev& net, long long j) { return net.blk_of[net.n_in + j]; }

__global__ void affine_count_kernel(NetDev net, AffineDev A, long long* __restrict__ counts) {
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= A.nvar) return;
  const VarInfo vi = decode_var(A, v);
  long long n = 0;
  switch (vi.kind) {
    case 0: n = 3; break;
    case 1: n = 1; break;
    case 2: n = 3; break;
    case 3: {
      const long long nk = net.n[nblk(net, vi.j) - 1];
      const double p = A.smin[vi.j] * A.smax[vi.j], q = A.smin[vi.j] + A.smax[vi.j];
      if (p != 0.0) n += nk * (nk + 1) / 2 + nk + 1;
      if (q != 0.0) n += nk + 1;
    } break;
    case 4: {
      const long long ni = net.n[nblk(net, vi.j) - 1], nj = net.n[nblk(net, vi.j + vi.t) - 1];
      n = 2 * (ni + 1) + 2 * (nj + 1) + 3;
    } break;
    default: n = net.n[nblk(net, vi.j) - 1] + 2; break;
  }
  counts[v] = n;
}

struct Writer {
  long long* ent;
  long long* var;
  double* val;
  const long long* col_ptr;
  const int* lo;
  long long v;
  __device__ __forceinline__ void put(long long at, int r, int c, double x) const {
    // upper triangle: r <= c; entry index inside column c of the cover
    ent[at] = col_ptr[c] + (r - lo[c]);
    var[at] = v;
    val[at] = x;
  }
  // (u e_pos' + e_pos u')[.,.] for one support index r of u with value x
  __device__ __forceinline__ void put_sym(long long at, int r, int pos, double x) const {
    if (r < pos) put(at, r, pos, x);
    else if (r > pos) put(at, pos, r, x);
    else put(at, pos, pos, 2.0 * x);
  }
};

// rho_j (e_pos)' + transpose, scaled: support = block rows (n_k entries) and the affine index
__device__ void emit_rho_sym(const NetDev& net, const Writer& w, long long at, long long j, int pos,
                             double scale) {
  const int B = nblk(net, j), kb = B - 1;        // rows of M[kb] produce block B; inputs are block kb
  const int nk = net.n[kb], nout = net.n[B];
  const int jl = (int)(net.n_in + j - net.off[B]);
  const double* M = net.M[kb];
  const int a = net.Zdim - 1;
  for (int r = threadIdx.x; r < nk; r += AFF_THREADS)
    w.put_sym(at + r, net.off[kb] + r, pos, scale * M[jl + (long long)r * nout]);
  if (threadIdx.x == 0) w.put_sym(at + nk, a, pos, scale * M[jl + (long long)nk * nout]);
}

__global__ void __launch_bounds__(AFF_THREADS)
affine_fill_kernel(NetDev net, AffineDev A, const long long* __restrict__ offs, long long* ent,
                   long long* var, double* val) {
  const long long v = blockIdx.x;
  const VarInfo vi = decode_var(A, v);
  const Writer w{ent, var, val, A.col_ptr, A.lo, v};
  long long at = offs[v];
  const int a = net.Zdim - 1, n0 = net.n_in, tid = threadIdx.x;
  switch (vi.kind) {
    case 0: {
      if (tid == 0) {
        const int i = (int)vi.j;
        w.put(at, i, i, -2.0);
        w.put(at + 1, i, a, A.x1min[i] + A.x1max[i]);
        w.put(at + 2, a, a, -2.0 * (A.x1min[i] * A.x1max[i]));
      }
    } break;
    case 1: {
      if (tid == 0) w.put(at, a, a, A.out_kind == NNSDP_OUT_HPLANE ? -2.0 : -1.0);
    } break;
    case 2: {
      if (tid == 0) {
        const int e = n0 + (int)vi.j;
        w.put(at, e, e, -2.0);
        w.put(at + 1, e, a, A.ymin[vi.j] + A.ymax[vi.j]);
        w.put(at + 2, a, a, -2.0 * (A.ymin[vi.j] * A.ymax[vi.j]));
      }
    } break;
    case 3: {
      const long long j = vi.j;
      const int B = nblk(net, j), kb = B - 1, nk = net.n[kb], nout = net.n[B];
      const int jl = (int)(n0 + j - net.off[B]), r0 = net.off[kb];
      const double* M = net.M[kb];
      const double p = A.smin[j] * A.smax[j], q = A.smin[j] + A.smax[j];
      const double bj = M[jl + (long long)nk * nout];
      if (p != 0.0) {
        const double m2p = -2.0 * p;
        for (int c = tid; c < nk; c += AFF_THREADS) {  // column c of the upper triangle: rows 0..c
          const double wc = M[jl + (long long)c * nout];
          const long long base = at + (long long)c * (c + 1) / 2;
          for (int r = 0; r <= c; ++r) w.put(base + r, r0 + r, r0 + c, m2p * (M[jl + (long long)r * nout] * wc));
        }
        at += (long long)nk * (nk + 1) / 2;
        for (int r = tid; r < nk; r += AFF_THREADS) w.put(at + r, r0 + r, a, m2p * (M[jl + (long long)r * nout] * bj));
        if (tid == 0) w.put(at + nk, a, a, m2p * (bj * bj));
        at += nk + 1;
      }
      if (q != 0.0) emit_rho_sym(net, w, at, j, n0 + (int)j, q);
    } break;
    case 4: {
      const long long i = vi.j, j = vi.j + vi.t;
      const int ei = n0 + (int)i, ej = n0 + (int)j;
      const long long ni = net.n[nblk(net, i) - 1], nj = net.n[nblk(net, j) - 1];
      emit_rho_sym(net, w, at, i, ei, 1.0);   at += ni + 1;
      emit_rho_sym(net, w, at, i, ej, -1.0);  at += ni + 1;
      emit_rho_sym(net, w, at, j, ei, -1.0);  at += nj + 1;
      emit_rho_sym(net, w, at, j, ej, 1.0);   at += nj + 1;
      if (tid == 0) {
        w.put(at, ei, ei, -2.0);
        w.put(at + 1, ei, ej, 2.0);
        w.put(at + 2, ej, ej, -2.0);
      }
    } break;
    default: {
      const long long j = vi.j;
      const double s = (vi.kind == 5) ? A.smin[j] : A.smax[j];
      emit_rho_sym(net, w, at, j, a, -s);  // -s (rho e_a' + e_a rho'): (r, a) -s w_r and (a, a) -2 s b_j
      const long long nk = net.n[nblk(net, j) - 1];
      if (tid == 0) w.put(at + nk + 1, n0 + (int)j, a, 1.0);
    } break;
  }
}

// constant part Z0 at the cover entries: Zout with gamma_out = 0 (and every other multiplier 0), taken
// from what prep_final_kernel left for a batch whose multipliers are all zero.
__global__ void affine_z0_kernel(NetDev net, BatchDev b, AffineDev A, double* __restrict__ z0) {
  const int c = blockIdx.x;  // column of Z
  const int K = net.K, n0 = net.n_in, a = net.Zdim - 1;
  const int lo = A.lo[c];
  const long long base = A.col_ptr[c];
  const int Bc = net.blk_of[c];
  const double* WK = net.M[K - 1];
  for (int r = lo + threadIdx.x; r <= c; r += blockDim.x) {
    double val = 0.0;
    if (c == a) {
      val = b.aff[r];
    } else {
      const int Br = net.blk_of[r];
      if (Br == 0 && Bc == 0) val = b.Z11[r + c * n0];
      else if (Br == 0 && Bc == K - 1 && b.has_s12) val = b.Z1K[r + (long long)(c - net.off[K - 1]) * n0];
      else if (Br == K - 1 && Bc == K - 1 && b.has_s22) {
        const int rl = r - net.off[K - 1], cl = c - net.off[K - 1];
        double s = 0.0;
        for (int m = 0; m < net.n_out; ++m)
          s = fma(WK[m + (long long)rl * net.n_out], b.U[m + (long long)cl * net.n_out], s);
        val = s;
      }
    }
    z0[base + (r - lo)] = val;
  }
}

}  // namespace

int launch_affine_count(const NetDev& net, const AffineDev& A, long long* counts, cudaStream_t st) {
  const long long blocks = (A.nvar + 255) / 256;
  affine_count_kernel<<<(int)blocks, 256, 0, st>>>(net, A, counts);
  return 1;
}

int launch_affine_fill(const NetDev& net, const AffineDev& A, const long long* offs, long long* ent,
                       long long* var, double* val, cudaStream_t st) {
  affine_fill_kernel<<<(int)A.nvar, AFF_THREADS, 0, st>>>(net, A, offs, ent, var, val);
  return 1;
}

int launch_affine_z0(const NetDev& net, const BatchDev& b, const AffineDev& A, double* z0,
                     cudaStream_t st) {
  affine_z0_kernel<<<net.Zdim, 128, 0, st>>>(net, b, A, z0);
  return 1;
}

}  // namespace nnsdp
