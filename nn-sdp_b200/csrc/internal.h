// Internal declarations shared by the host runtime and the sm_100a kernels.
// Nothing in here is part of the C ABI (see include/nnsdp_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/nnsdp_b200.h"

namespace nnsdp {

// ---------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define NN_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) return ::nnsdp::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define NN_CHECK(cond, code, ...)       \
  do {                                  \
    if (!(cond)) {                      \
      ::nnsdp::set_error(__VA_ARGS__);  \
      return (code);                    \
    }                                   \
  } while (0)

#define NN_TRY(expr)               \
  do {                             \
    int32_t _s = (expr);           \
    if (_s != NNSDP_OK) return _s; \
  } while (0)

// ---------------------------------------------------------------------------------
// host-side shape bookkeeping (0-based internally)
//   block b = 0..K-1 holds x_{b+1} (size n[b]); "block" K is the affine index a.
//   neuron j = 0..acdim-1 is z-index n[0]+j; the neurons of block b (b>=1) are
//   [noff(b), noff(b)+n[b]) with noff(b) = off[b]-n[0]; they are the rows of W_{b-1}.
// ---------------------------------------------------------------------------------
struct Shape {
  int K = 0;
  std::vector<int64_t> n;     // K+1 : xdims
  std::vector<int64_t> off;   // K+1 : off[b] = sum n[0..b-1]; off[K] = a = Zdim-1
  std::vector<int64_t> xoff;  // K+2 : offsets into stacked x_intvs
  int64_t Zdim = 0, acdim = 0, xtot = 0;
  int64_t n_in() const { return n[0]; }
  int64_t n_out() const { return n[K]; }
  int64_t n_last_hidden() const { return n[K - 1]; }
};

// A clique = one or two contiguous global index ranges [lo, hi] (0-based, inclusive).
struct CliqueRanges {
  int nseg = 0;
  int64_t lo[2] = {0, 0}, hi[2] = {0, 0};
  int64_t size() const {
    int64_t s = 0;
    for (int i = 0; i < nseg; ++i) s += hi[i] - lo[i] + 1;
    return s;
  }
};

struct CliqueInfoHost {
  std::vector<CliqueRanges> ck;            // p cliques
  std::vector<std::vector<int64_t>> d1, d2;  // 1-based local indices (Dk1, Dk2)
};

int32_t make_cliques_host(const Shape& sh, int64_t beta, CliqueInfoHost* out);
int64_t lambda_dim(int64_t acdim, int64_t beta);

// ---------------------------------------------------------------------------------
// emission plan: every clique block is cut into strips (<= STRIP_ROWS rows inside one
// Z block) x chunks (<= CHUNK_COLS columns inside one Z block); each tile gets a class.
// ---------------------------------------------------------------------------------
constexpr int STRIP_ROWS = 128;
constexpr int CHUNK_COLS = 32;

enum TileClass : uint8_t {
  TILE_ZERO = 0,      // structurally zero
  TILE_AFFCOL = 1,    // the affine column (1 column)
  TILE_DIAGPLAIN = 2, // same block, away from band/slivers: Gram copy / output Gram / zero
  TILE_WT = 3,        // rows in block b, cols = neurons of block b+1: W_b' * M  (window sum)
  TILE_WTT = 4,       // transpose of the above
  TILE_GENERAL = 5,   // everything else: per-entry evaluation
};

struct StripDev {
  int32_t clique;     // clique index
  int32_t row0;       // first local row in the clique block
  int32_t nrows;      // <= STRIP_ROWS
  int32_t grow0;      // global z index of the first row
  int32_t blk;        // Z block of the rows
  int32_t arow;       // 1: this strip also writes the affine row of its clique
  int32_t tile0;      // index of this strip's first tile class in the tile table
  int32_t pad;
};

struct ChunkDev {
  int32_t col0;       // first local column in the clique block
  int32_t ncols;      // <= CHUNK_COLS
  int32_t gcol0;      // global z index of the first column
  int32_t blk;        // Z block of the columns (K = affine)
};

struct CliqueDev {
  int64_t out_off;    // offset (doubles) of this block inside one query's output
  int32_t n;          // |C_k|
  int32_t ld;         // leading dimension of the block
  int32_t chunk0;     // first chunk of this clique in the chunk table
  int32_t nchunks;
  int32_t len1;       // length of the first segment
  int32_t g1, g2;     // global start of segment 1 / 2
  int32_t pad;
};

struct PlanHost {
  std::vector<StripDev> strips;
  std::vector<ChunkDev> chunks;
  std::vector<CliqueDev> cliques;
  std::vector<uint8_t> tiles;
  int64_t per_query_doubles = 0;
};

int32_t build_plan(const Shape& sh, int64_t beta, const std::vector<CliqueRanges>& cliques,
                   PlanHost* plan);

// ---------------------------------------------------------------------------------
// device-side descriptors (passed by value to kernels)
// ---------------------------------------------------------------------------------
struct NetDev {
  int K, n_in, n_out, Zdim, acdim, xtot;
  const int* n;               // K+1
  const int* off;             // K+1
  const int* xoff;            // K+2
  const double* const* M;     // K   : [W_k b_k], ld = n[k+1]            (neuron-contiguous)
  const double* const* Wt;    // K   : W_k', (n[k] x n[k+1]), ld = ldT[k] (input-contiguous)
  const int* ldT;             // K
  const double* bias_all;     // acdim : b_1..b_{K-1} stacked
};

// Per-batch device arrays.  Strides are in doubles between queries (0 = shared).
struct BatchDev {
  int Q, beta;
  int out_kind, has_s22, has_s1x;
  // inputs
  const double* x1min;  long long s_x1min;
  const double* x1max;  long long s_x1max;
  const double* ymin;   long long s_ymin;
  const double* ymax;   long long s_ymax;
  const double* smin;   long long s_smin;
  const double* smax;   long long s_smax;
  const double* gin;    long long s_gin;
  const double* gbnd;   long long s_gbnd;
  const double* gsec;   long long s_gsec;
  const double* outS;   long long s_outS;
  const double* outvec; long long s_outvec;
  const double* outinvP; long long s_outinvP;
  const double* gout;   long long s_gout;
  long long lamdim;
  // prepared per-query vectors (stride = natural size)
  double* d11;    // acdim
  double* Mb;     // (beta+1) * acdim
  double* dg;     // acdim
  double* u;      // acdim
  double* aff;    // Zdim
  int* cnt;       // K   (active-neuron count feeding the Gram of block b)
  double* Z11;    // n_in * n_in
  double* Z1K;    // n_in * n[K-1]
  double* U;      // n_out * n[K-1]
};

struct GramDev {
  double* scratch;           // chunk * gram_per_query
  long long per_query;       // doubles
  const long long* goff;     // K-1 : offset of block b's Gram inside one query's scratch
};

struct PlanDev {
  const StripDev* strips;
  const ChunkDev* chunks;
  const CliqueDev* cliques;
  const uint8_t* tiles;
  int nstrips;
  long long per_query;       // doubles per query in the output
};

// ---------------------------------------------------------------------------------
// kernel launchers (each returns the number of kernels it launched, < 0 on error)
// ---------------------------------------------------------------------------------
// K1: one IBP layer for Q boxes: y = W+ x- + W- x+ + b (and the mirrored bound); optional ReLU.
int ibp_layer_launch(const double* Mk, int n_out_k, int n_in_k, const double* xin_min,
                     const double* xin_max, long long x_stride, double* xout_min, double* xout_max,
                     double* acx_min, double* acx_max, long long acx_stride, int Q, int relu,
                     int write_x, int* flag_bad, cudaStream_t st);
// affine column: aff[r] += sum_j Wt[r, j] u[j]  (batched over queries)
int affine_layer_launch(const double* Wt, int ldT, int n_rows, int n_neurons, const double* u,
                        long long u_stride, double* aff, long long aff_stride, int Q,
                        cudaStream_t st);
// K2: smin/smax from acx bounds.
int launch_sector_minmax(long long n, const double* acxmin, const double* acxmax, double* smin,
                         double* smax, cudaStream_t st);
// K2b: per-query QC diagonals, band multipliers, affine-column seeds.
int launch_prep(const NetDev& net, const BatchDev& b, int* err_flag, cudaStream_t st);
// K3: Gram contractions (DMMA) for queries [q0, q0+nq).
int launch_gram(const NetDev& net, const BatchDev& b, const GramDev& g, int q0, int nq,
                cudaStream_t st);
// K4+K5: emit all clique blocks of queries [q0, q0+nq) to out (slot s = q - q0).
int launch_emit(const NetDev& net, const BatchDev& b, const GramDev& g, const PlanDev& plan,
                int q0, int nq, double* out, cudaStream_t st);

}  // namespace nnsdp
