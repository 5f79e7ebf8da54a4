// K4 + K5: emission of the dense FP64 output matrices -- the clique blocks Z[C_k, C_k] of the
// chordal decomposition (reference: Ec(Ck) Z Ec(Ck)' for the cliques of
// src/Methods/chordal_cliques.jl:13-59, the blocks setupZksum! scatters at
// src/Methods/chordal_sdp.jl:60-93) or the whole Z (src/Methods/chordal_sdp.jl:114,145).
//
// Z is never materialised: every output entry is evaluated from the closed form
//   Z[r,c] = [same block] (G_b[r,c] | Z11 | W_K' S22 W_K)  +  [x_1 / x_K] S12 W_K
//          + F(r,c) + F(c,r)  +  [band] (-2 T[jr,jc] - 2 gamma_bnd [jr == jc])
//   F(r,c) = sum_{j in layer(blk(r)+1), |j - jc| <= beta} W[j, r] M[j, jc],   M = diag(q lambda) + T
//  and written once, column-major, with consecutive threads on consecutive rows.
// The kernel is HBM-write bound: 8 |C_k|^2 bytes per block.  Tiles carry host-computed flags
// saying which terms can be non-zero inside them, so the bulk tiles reduce to a zero fill, a
// Gram copy or a (2 beta + 1)-wide window sum over coalesced columns of W' / W.
#include "internal.h"

namespace nnsdp {

namespace {

constexpr int ETHREADS = 256;

struct QView {  // per-query pointers, resolved once per CTA
  const double* Md;
  const double* T0;
  const double* Bt;
  const double* gbnd;
  const double* aff;
  const double* Z11;
  const double* Z1K;
  const double* U;
  const int* cnt;
  const double* G;
};

// M[j, c] = delta_jc q_j lambda_j + T[j, c],  |j - c| <= beta
__device__ __forceinline__ double m_coef(const QView& v, long long acdim, int j, int c) {
  if (j == c) return v.Md[j];
  const int t = j > c ? j - c : c - j;
  return v.Bt[(long long)(t - 1) * acdim + (j < c ? j : c)];
}

__global__ void __launch_bounds__(ETHREADS)
emit_kernel(NetDev net, BatchDev b, GramDev g, PlanDev plan, int q0, double* __restrict__ out) {
  const TileDev t = plan.tiles[blockIdx.x];
  const int slot = blockIdx.y, q = q0 + slot;
  const MatDev mat = plan.mats[t.mat];
  const int K = net.K, n0 = net.n_in, a = net.Zdim - 1, beta = b.beta;
  const long long acdim = net.acdim;

  const int TR = plan.tile_rows;
  const int tr = threadIdx.x % TR, cg = threadIdx.x / TR, ncg = ETHREADS / TR;
  if (tr >= t.nrows) return;

  double* o = out + (long long)slot * plan.per_query + mat.out_off + (t.row0 + tr) +
              (long long)t.col0 * mat.ld;
  const uint32_t flags = t.flags;
  if ((flags & TF_ALL) == 0) {  // structurally zero tile
    for (int c = cg; c < t.ncols; c += ncg) o[(long long)c * mat.ld] = 0.0;
    return;
  }

  QView v;
  v.Md = b.Md + (long long)q * acdim;
  v.T0 = b.T0 + (long long)q * acdim;
  v.Bt = b.Bt + (long long)q * beta * acdim;
  v.gbnd = b.gbnd + q * b.s_gbnd;
  v.aff = b.aff + (long long)q * net.Zdim;
  v.Z11 = b.Z11 + (long long)q * n0 * n0;
  v.Z1K = b.Z1K + (long long)q * n0 * net.n[K - 1];
  v.U = b.U + (long long)q * net.n_out * net.n[K - 1];
  v.cnt = b.cnt + (long long)q * K;
  v.G = g.scratch + (long long)slot * g.per_query;

  // row-dependent quantities
  const int gr = t.grow0 + tr;
  const int Br = net.blk_of[gr];
  const int rl = gr - (Br < K ? net.off[Br] : a);
  const int jr = (gr >= n0 && gr < a) ? gr - n0 : -1;
  // F(r,c): rows of block Br feed the neurons [Lr0, Lr0 + nLr) through M[Br]
  const bool r_feeds = (Br <= K - 2);
  const int Lr0 = r_feeds ? net.off[Br + 1] - n0 : 0;
  const int nLr = r_feeds ? net.n[Br + 1] : 0;
  const double* WtR = r_feeds ? net.Wt[Br] + rl : nullptr;
  const int ldTR = r_feeds ? net.ldT[Br] : 0;
  const double* WK = net.M[K - 1];
  const int n_out = net.n_out;

  for (int c = cg; c < t.ncols; c += ncg) {
    const int gc = t.gcol0 + c;
    double val = 0.0;
    if (gr == a || gc == a) {
      val = v.aff[gr == a ? gc : gr];
    } else {
      const int Bc = net.blk_of[gc];
      const int cl = gc - net.off[Bc];
      const int jc = gc >= n0 ? gc - n0 : -1;
      if ((flags & TF_SAME) && Br == Bc) {
        if (Br <= K - 2 && v.cnt[Br] > 0) val += v.G[g.goff[Br] + rl + (long long)cl * g.ldG[Br]];
        if (Br == 0) val += v.Z11[rl + cl * n0];
        if (Br == K - 1 && b.has_s22) {
          // evaluated as (min, max) so that Z[r,c] and Z[c,r] are bit-identical
          const int lo = rl < cl ? rl : cl, hi = rl < cl ? cl : rl;
          double s = 0.0;
          for (int m = 0; m < n_out; ++m)
            s = fma(WK[m + (long long)lo * n_out], v.U[m + (long long)hi * n_out], s);
          val += s;
        }
      }
      if ((flags & TF_1K) && b.has_s12) {
        if (Br == 0 && Bc == K - 1) val += v.Z1K[rl + (long long)cl * n0];
        if (Bc == 0 && Br == K - 1) val += v.Z1K[cl + (long long)rl * n0];
      }
      double f1 = 0.0, f2 = 0.0;
      if ((flags & TF_RC) && r_feeds && jc >= 0) {
        const int jlo = max(Lr0, jc - beta), jhi = min(Lr0 + nLr - 1, jc + beta);
        for (int j = jlo; j <= jhi; ++j)
          f1 = fma(WtR[(long long)(j - Lr0) * ldTR], m_coef(v, acdim, j, jc), f1);
      }
      if ((flags & TF_CR) && Bc <= K - 2 && jr >= 0) {
        const int Lc0 = net.off[Bc + 1] - n0, nLc = net.n[Bc + 1];
        const double* Wc = net.M[Bc] + (long long)cl * nLc;  // column cl of W_Bc: neuron-contiguous
        const int jlo = max(Lc0, jr - beta), jhi = min(Lc0 + nLc - 1, jr + beta);
        for (int j = jlo; j <= jhi; ++j) f2 = fma(Wc[j - Lc0], m_coef(v, acdim, j, jr), f2);
      }
      val += f1 + f2;  // commutative: Z[r,c] and Z[c,r] come out bit-identical
      if ((flags & TF_BAND) && jr >= 0 && jc >= 0) {
        const int d = jr > jc ? jr - jc : jc - jr;
        if (d == 0)
          val += -2.0 * v.T0[jr] - 2.0 * v.gbnd[jr];
        else if (d <= beta)
          val += -2.0 * v.Bt[(long long)(d - 1) * acdim + (jr < jc ? jr : jc)];
      }
    }
    o[(long long)c * mat.ld] = val;
  }
}

}  // namespace

int launch_emit(const NetDev& net, const BatchDev& b, const GramDev& g, const PlanDev& plan,
                int q0, int nq, double* out, cudaStream_t st) {
  if (plan.ntiles <= 0 || nq <= 0) return 0;
  dim3 grid(plan.ntiles, nq);
  emit_kernel<<<grid, ETHREADS, 0, st>>>(net, b, g, plan, q0, out);
  return 1;
}

}  // namespace nnsdp
