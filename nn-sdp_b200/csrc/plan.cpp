// Host-side integer work: clique index sets and the emission plan.
//
// make_cliques_host restates makeCliques (reference: src/Methods/chordal_cliques.jl:13-59)
// with index arithmetic only; nothing is materialised as a selector matrix (the reference
// builds Ec(Ck, Zdim) at src/Methods/chordal_sdp.jl:54 and then discards it).
#include <algorithm>

#include "internal.h"

namespace nnsdp {

int64_t lambda_dim(int64_t acdim, int64_t beta) {
  // _lambda_dim = sum((acxdim - beta):acxdim), src/Qc/activ_sector.jl:18
  int64_t s = 0;
  for (int64_t t = acdim - beta; t <= acdim; ++t) s += t;
  return s;
}

static int64_t Sfun(const Shape& sh, int64_t k) {  // S(k) = sum(xdims[1:k]), k = 0..K+1
  int64_t s = 0;
  for (int64_t i = 0; i < k; ++i) s += sh.n[i];
  return s;
}

int32_t make_cliques_host(const Shape& sh, int64_t beta, CliqueInfoHost* out) {
  const int K = sh.K;
  out->ck.clear();
  out->d1.clear();
  out->d2.clear();
  int64_t p = 1;
  for (int64_t i = 1; i <= K; ++i) {  // chordal_cliques.jl:22-27
    if (Sfun(sh, i + 1) + beta >= Sfun(sh, K - 1)) {
      p = i;
      break;
    }
  }
  for (int64_t k = 1; k <= p - 1; ++k) {  // :31-52
    CliqueRanges c;
    c.nseg = 2;
    c.lo[0] = Sfun(sh, k - 1);             // 0-based of S(k-1)+1
    c.hi[0] = Sfun(sh, k + 1) + beta - 1;  // 0-based of S(k+1)+beta
    c.lo[1] = Sfun(sh, K - 1);
    c.hi[1] = Sfun(sh, K);                 // 0-based of S(K)+1 (the affine index)
    NN_CHECK(c.hi[0] + 1 <= c.lo[1] + 1, NNSDP_ERR_ASSERT,
             "makeCliques: Ck1[end] <= Ck2[1] violated (chordal_cliques.jl:35)");
    const int64_t ckdim = c.size();
    std::vector<int64_t> d1, d2;
    if (k == 1) {
      for (int64_t i = 1; i <= ckdim; ++i) d1.push_back(i);
    } else {
      const int64_t nk = sh.n[k - 1], nk1 = sh.n[k];  // zdims[k], zdims[k+1] (1-based)
      for (int64_t i = 1; i <= nk + nk1 + beta; ++i) d1.push_back(i);
      d1.push_back(ckdim);
      for (int64_t i = nk + nk1 + 1; i <= ckdim; ++i) d2.push_back(i);
    }
    out->ck.push_back(c);
    out->d1.push_back(d1);
    out->d2.push_back(d2);
  }
  CliqueRanges c;  // :55-57
  c.nseg = 1;
  c.lo[0] = Sfun(sh, p - 1);
  c.hi[0] = Sfun(sh, K);
  std::vector<int64_t> d1;
  for (int64_t i = 1; i <= c.size(); ++i) d1.push_back(i);
  out->ck.push_back(c);
  out->d1.push_back(d1);
  out->d2.push_back({});
  return NNSDP_OK;
}

namespace {

struct Piece {
  int64_t g0, g1;  // global range (inclusive)
  int blk;
  int64_t l0;      // local index of g0 inside the clique block
};

bool intersects(int64_t a0, int64_t a1, int64_t b0, int64_t b1) {
  return a0 <= a1 && b0 <= b1 && std::max(a0, b0) <= std::min(a1, b1);
}

uint8_t classify(const Shape& sh, int64_t beta, int Bi, int64_t r0, int64_t r1, int Bj, int64_t c0,
                 int64_t c1) {
  const int K = sh.K;
  if (Bj == K) return TILE_AFFCOL;
  const int64_t n0 = sh.n[0];
  const bool f_same = (Bi == Bj);
  bool f23 = false, f32 = false;
  if (Bi <= K - 2) {
    const int64_t lo = sh.off[Bi + 1] - beta, hi = sh.off[Bi + 2] - 1 + beta;
    f23 = intersects(std::max(c0, n0), c1, lo, hi);
  }
  if (Bj <= K - 2) {
    const int64_t lo = sh.off[Bj + 1] - beta, hi = sh.off[Bj + 2] - 1 + beta;
    f32 = intersects(std::max(r0, n0), r1, lo, hi);
  }
  const int64_t dist = std::max<int64_t>(0, std::max(c0 - r1, r0 - c1));
  const bool f_band = (r1 >= n0 && c1 >= n0 && dist <= beta);
  const bool f_1K = (Bi == 0 && Bj == K - 1) || (Bi == K - 1 && Bj == 0);
  if (!f_same && !f23 && !f32 && !f_band && !f_1K) return TILE_ZERO;
  if (f_same && !f23 && !f32 && !f_band && !f_1K) return TILE_DIAGPLAIN;
  if (f23 && !f_same && !f32 && !f_band && !f_1K && Bj == Bi + 1) return TILE_WT;
  if (f32 && !f_same && !f23 && !f_band && !f_1K && Bi == Bj + 1) return TILE_WTT;
  return TILE_GENERAL;
}

}  // namespace

int32_t build_plan(const Shape& sh, int64_t beta, const std::vector<CliqueRanges>& cliques,
                   PlanHost* plan) {
  const int K = sh.K;
  const int64_t a = sh.off[K];
  plan->strips.clear();
  plan->chunks.clear();
  plan->cliques.clear();
  plan->tiles.clear();
  int64_t out_off = 0;
  for (size_t ci = 0; ci < cliques.size(); ++ci) {
    const CliqueRanges& c = cliques[ci];
    const int64_t n = c.size();
    NN_CHECK(n < (int64_t(1) << 30), NNSDP_ERR_ARG, "clique too large");
    // cut the segments into per-block pieces
    std::vector<Piece> pieces;
    int64_t local = 0;
    bool has_a = false;
    for (int s = 0; s < c.nseg; ++s) {
      for (int b = 0; b <= K; ++b) {
        const int64_t b0 = sh.off[b], b1 = (b == K) ? a : sh.off[b + 1] - 1;
        const int64_t g0 = std::max(c.lo[s], b0), g1 = std::min(c.hi[s], b1);
        if (g0 > g1) continue;
        pieces.push_back({g0, g1, b, local + (g0 - c.lo[s])});
        if (b == K) has_a = true;
      }
      local += c.hi[s] - c.lo[s] + 1;
    }
    NN_CHECK(has_a, NNSDP_ERR_ARG, "clique without the affine index");
    CliqueDev cd{};
    cd.out_off = out_off;
    cd.n = (int32_t)n;
    cd.ld = (int32_t)n;
    cd.chunk0 = (int32_t)plan->chunks.size();
    cd.len1 = (int32_t)(c.hi[0] - c.lo[0] + 1);
    cd.g1 = (int32_t)c.lo[0];
    cd.g2 = (int32_t)(c.nseg > 1 ? c.lo[1] : 0);
    // chunks (columns), affine column included as a 1-wide chunk
    for (const Piece& pc : pieces) {
      for (int64_t g = pc.g0; g <= pc.g1; g += CHUNK_COLS) {
        ChunkDev ch{};
        ch.gcol0 = (int32_t)g;
        ch.ncols = (int32_t)std::min<int64_t>(CHUNK_COLS, pc.g1 - g + 1);
        ch.col0 = (int32_t)(pc.l0 + (g - pc.g0));
        ch.blk = pc.blk;
        plan->chunks.push_back(ch);
      }
    }
    cd.nchunks = (int32_t)plan->chunks.size() - cd.chunk0;
    // strips (rows), affine row excluded (written by the last strip's epilogue)
    const size_t first_strip = plan->strips.size();
    for (const Piece& pc : pieces) {
      if (pc.blk == K) continue;
      for (int64_t g = pc.g0; g <= pc.g1; g += STRIP_ROWS) {
        StripDev st{};
        st.clique = (int32_t)ci;
        st.grow0 = (int32_t)g;
        st.nrows = (int32_t)std::min<int64_t>(STRIP_ROWS, pc.g1 - g + 1);
        st.row0 = (int32_t)(pc.l0 + (g - pc.g0));
        st.blk = pc.blk;
        st.arow = 0;
        st.tile0 = (int32_t)plan->tiles.size();
        for (int j = 0; j < cd.nchunks; ++j) {
          const ChunkDev& ch = plan->chunks[cd.chunk0 + j];
          plan->tiles.push_back(classify(sh, beta, st.blk, st.grow0, st.grow0 + st.nrows - 1, ch.blk,
                                         ch.gcol0, ch.gcol0 + ch.ncols - 1));
        }
        plan->strips.push_back(st);
      }
    }
    NN_CHECK(plan->strips.size() > first_strip, NNSDP_ERR_ARG, "clique with only the affine index");
    plan->strips.back().arow = 1;
    plan->cliques.push_back(cd);
    out_off += n * n;
  }
  plan->per_query_doubles = out_off;
  return NNSDP_OK;
}

}  // namespace nnsdp
