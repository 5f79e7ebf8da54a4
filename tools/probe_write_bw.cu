// Microbenchmarks behind the design of the block emitter: how fast can B200 take a
// column-major tiled write stream, and what does a leading dimension that is not a multiple of
// 4 doubles (32 B sectors) cost?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe tools/probe_write_bw.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// tile = 128 rows x 32 cols, 256 threads, thread = (row, column group), 8 B stores
__global__ void tile_fill8(double* out, int n, int ld, int tiles_r, double v) {
  const int tr = threadIdx.x & 127, cg = threadIdx.x >> 7;
  const int r0 = (blockIdx.x % tiles_r) * 128, c0 = (blockIdx.x / tiles_r) * 32;
  double* o = out + (size_t)blockIdx.y * (size_t)ld * n;
  const int r = r0 + tr;
  if (r >= n) return;
  for (int c = cg; c < 32; c += 2) if (c0 + c < n) o[r + (size_t)(c0 + c) * ld] = v;
}
// same tile, but rows are re-based per column so that every warp store covers whole 32 B sectors
__global__ void tile_fill8_aligned(double* out, int n, int ld, int tiles_r, double v) {
  const int tr = threadIdx.x & 127, cg = threadIdx.x >> 7;
  const int r0 = (blockIdx.x % tiles_r) * 128, c0 = (blockIdx.x / tiles_r) * 32;
  double* o = out + (size_t)blockIdx.y * (size_t)ld * n;
  const int rend = min(r0 + 128, n);
  for (int c = cg; c < 32; c += 2) {
    if (c0 + c >= n) break;
    double* col = o + (size_t)(c0 + c) * ld;
    const int mis = (int)(((size_t)(col + r0) >> 3) & 3);     // doubles past a sector boundary
    const int start = r0 - mis;                                 // sector-aligned virtual start
    for (int r = start + tr; r < rend; r += 128) if (r >= r0) col[r] = v;
    // rows [start+128, rend) are picked up by the second trip of the loop (at most 3 rows)
  }
}
// 16 B stores where the address allows it (ld even)
__global__ void tile_fill16(double* out, int n, int ld, int tiles_r, double v) {
  const int tr = threadIdx.x & 63, cg = threadIdx.x >> 6;
  const int r0 = (blockIdx.x % tiles_r) * 128, c0 = (blockIdx.x / tiles_r) * 32;
  double* o = out + (size_t)blockIdx.y * (size_t)ld * n;
  const int r = r0 + 2 * tr;
  if (r + 1 >= n) return;
  for (int c = cg; c < 32; c += 4) if (c0 + c < n) *reinterpret_cast<double2*>(o + r + (size_t)(c0 + c) * ld) = make_double2(v, v);
}
// one CTA writes the same tile of `nslots` matrices that lie `slot_stride` doubles apart
__global__ void tile_fill8_slots(double* out, int n, int ld, int tiles_r, size_t slot_stride, int nslots, double v) {
  const int tr = threadIdx.x & 127, cg = threadIdx.x >> 7;
  const int r0 = (blockIdx.x % tiles_r) * 128, c0 = (blockIdx.x / tiles_r) * 32;
  const int r = r0 + tr;
  if (r >= n) return;
  for (int s = 0; s < nslots; ++s) {
    double* o = out + ((size_t)blockIdx.y * nslots + s) * slot_stride;
    for (int c = cg; c < 32; c += 2) if (c0 + c < n) o[r + (size_t)(c0 + c) * ld] = v;
  }
}
// slot index fastest in the grid: blockIdx.x = slot, blockIdx.y = tile
__global__ void tile_fill8_slotfast(double* out, int n, int ld, int tiles_r, size_t slot_stride, double v) {
  const int tr = threadIdx.x & 127, cg = threadIdx.x >> 7;
  const int r0 = (blockIdx.y % tiles_r) * 128, c0 = (blockIdx.y / tiles_r) * 32;
  const int r = r0 + tr;
  if (r >= n) return;
  double* o = out + (size_t)blockIdx.x * slot_stride;
  for (int c = cg; c < 32; c += 2) if (c0 + c < n) o[r + (size_t)(c0 + c) * ld] = v;
}
// tall strips: a CTA owns `h` rows x 8 columns; per column the 256 threads walk 2 KB chunks that start on
// a 32 B sector boundary (rows before the strip are masked), so every warp store covers 8 whole sectors
__global__ void strip_fill_aligned(double* out, int n, int ld, int h, int strips_r, double v) {
  const int r_lo = (blockIdx.x % strips_r) * h, c0 = (blockIdx.x / strips_r) * 8;
  const int r_hi = min(r_lo + h, n);
  double* o = out + (size_t)blockIdx.y * (size_t)ld * n;
  for (int c = c0; c < min(c0 + 8, n); ++c) {
    double* col = o + (size_t)c * ld;
    const int mis = (int)(((size_t)(col + r_lo) >> 3) & 3);
    for (int r = r_lo - mis + threadIdx.x; r < r_hi; r += 256) if (r >= r_lo) col[r] = v;
  }
}
// aligned strips over a subset of the row strips only (mask bit i = row strip i is written): mimics the
// emitter's fill kernel, which leaves the window regions of every column to another kernel
__global__ void strip_fill_masked(double* out, int n, int ld, int h, int strips_r, unsigned mask, int nsel, double v) {
  int k = blockIdx.x % nsel, rs = 0;
  for (int i = 0; i < strips_r; ++i) if (mask >> i & 1) { if (k == 0) { rs = i; break; } --k; }
  const int r_lo = rs * h, c0 = (blockIdx.x / nsel) * 8;
  const int r_hi = min(r_lo + h, n);
  double* o = out + (size_t)blockIdx.y * (size_t)ld * n;
  for (int c = c0; c < min(c0 + 8, n); ++c) {
    double* col = o + (size_t)c * ld;
    const int mis = (int)(((size_t)(col + r_lo) >> 3) & 3);
    for (int r = r_lo - mis + threadIdx.x; r < r_hi; r += 256) if (r >= r_lo) col[r] = v;
  }
}
// ---- does interleaving the fill strips and the window tiles of one column panel pay? ----------------
// rows [0, 1002) and [2004, 3003) of every column are "fill" (aligned 501-row strips x 8 cols), rows
// [1002, 2004) are "window" (128 x 32 tiles, 8 B stores, like emit_window_kernel).  Variant `panel` runs
// both in ONE kernel with CTAs ordered by 32-column panel; the baseline runs two kernels back to back.
__device__ __forceinline__ void fill_strip_dev(double* o, int ld, int r_lo, int r_hi, int c0, int n, double v) {
  for (int c = c0; c < min(c0 + 8, n); ++c) {
    double* col = o + (size_t)c * ld;
    const int mis = (int)(((size_t)(col + r_lo) >> 3) & 3);
    for (int r = r_lo - mis + threadIdx.x; r < r_hi; r += 256) if (r >= r_lo) col[r] = v;
  }
}
__device__ __forceinline__ void window_tile_dev(double* o, int ld, int r0, int r_hi, int c0, int n, double v) {
  const int tr = threadIdx.x & 127, cg = threadIdx.x >> 7;
  const int r = r0 + tr;
  if (r >= r_hi) return;
  for (int c = cg; c < 32; c += 2) if (c0 + c < n) o[r + (size_t)(c0 + c) * ld] = v;
}
// items of one panel: 4 column strips x 4 row strips (0,1,4,5) = 16 fill items, then 8 window tiles
__global__ void panel_unified(double* out, int n, int ld, double v) {
  const int panel = blockIdx.x / 24, it = blockIdx.x % 24;
  double* o = out + (size_t)blockIdx.y * (size_t)ld * n;
  const int c0 = panel * 32;
  if (it < 16) {
    const int cs = it / 4, k = it % 4, rs = (k < 2) ? k : k + 2;
    fill_strip_dev(o, ld, rs * 501, min(rs * 501 + 501, n), c0 + cs * 8, n, v);
  } else {
    window_tile_dev(o, ld, 1002 + (it - 16) * 128, 2004, c0, n, v);
  }
}
__global__ void panel_fill_only(double* out, int n, int ld, double v) {
  const int panel = blockIdx.x / 16, it = blockIdx.x % 16;
  double* o = out + (size_t)blockIdx.y * (size_t)ld * n;
  const int cs = it / 4, k = it % 4, rs = (k < 2) ? k : k + 2;
  fill_strip_dev(o, ld, rs * 501, min(rs * 501 + 501, n), panel * 32 + cs * 8, n, v);
}
__global__ void panel_window_only(double* out, int n, int ld, double v) {
  const int panel = blockIdx.x / 8, it = blockIdx.x % 8;
  double* o = out + (size_t)blockIdx.y * (size_t)ld * n;
  window_tile_dev(o, ld, 1002 + it * 128, 2004, panel * 32, n, v);
}
// window part as tall strips too (what a unified kernel could do if the window program allowed it)
__global__ void panel_unified_strips(double* out, int n, int ld, double v) {
  const int panel = blockIdx.x / 24, it = blockIdx.x % 24;
  double* o = out + (size_t)blockIdx.y * (size_t)ld * n;
  const int cs = it / 6, rs = it % 6;
  fill_strip_dev(o, ld, rs * 501, min(rs * 501 + 501, n), panel * 32 + cs * 8, n, v);
}
// same strips, no alignment (rows start at r_lo)
__global__ void strip_fill_plain(double* out, int n, int ld, int h, int strips_r, double v) {
  const int r_lo = (blockIdx.x % strips_r) * h, c0 = (blockIdx.x / strips_r) * 8;
  const int r_hi = min(r_lo + h, n);
  double* o = out + (size_t)blockIdx.y * (size_t)ld * n;
  for (int c = c0; c < min(c0 + 8, n); ++c) {
    double* col = o + (size_t)c * ld;
    for (int r = r_lo + threadIdx.x; r < r_hi; r += 256) col[r] = v;
  }
}
__global__ void linear_fill(double* out, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}

int main() {
  const int nmat = 64;
  const int n = 3003;
  size_t cap = (size_t)nmat * 3008 * 3008 + 1024;
  double* buf; CK(cudaMalloc(&buf, cap * 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto fn, double bytes) {
    for (int i = 0; i < 2; ++i) fn();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; ++i) fn();
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms / reps, bytes / (ms / reps * 1e-3) / 1e9);
  };
  const int tiles_r = (n + 127) / 128, tiles_c = (n + 31) / 32;
  dim3 grid(tiles_r * tiles_c, nmat);
  double bytes = (double)nmat * n * n * 8;
  timeit("cudaMemsetAsync", [&] { cudaMemsetAsync(buf, 0, (size_t)nmat * n * n * 8); }, bytes);
  timeit("linear_fill 8B grid-stride", [&] { linear_fill<<<148 * 16, 256>>>(buf, (size_t)nmat * n * n, 1.0); }, bytes);
  timeit("tile_fill8 ld=3003", [&] { tile_fill8<<<grid, 256>>>(buf, n, 3003, tiles_r, 1.0); }, bytes);
  timeit("tile_fill8 ld=3004", [&] { tile_fill8<<<grid, 256>>>(buf, n, 3004, tiles_r, 1.0); }, bytes);
  timeit("tile_fill8 ld=3008", [&] { tile_fill8<<<grid, 256>>>(buf, n, 3008, tiles_r, 1.0); }, bytes);
  timeit("tile_fill8_aligned ld=3003", [&] { tile_fill8_aligned<<<grid, 256>>>(buf, n, 3003, tiles_r, 1.0); }, bytes);
  for (int h : {1001, 501, 256}) {
    const int strips_r = (n + h - 1) / h;
    dim3 g2(strips_r * ((n + 7) / 8), nmat);
    char nm[96];
    snprintf(nm, sizeof nm, "strip_fill_aligned ld=3003 h=%d x 8 cols", h);
    timeit(nm, [&] { strip_fill_aligned<<<g2, 256>>>(buf, n, 3003, h, strips_r, 1.0); }, bytes);
    snprintf(nm, sizeof nm, "strip_fill_plain   ld=3003 h=%d x 8 cols", h);
    timeit(nm, [&] { strip_fill_plain<<<g2, 256>>>(buf, n, 3003, h, strips_r, 1.0); }, bytes);
  }
  {
    const int h = 501, strips_r = 6;
    for (unsigned mask : {0x3Fu, 0x33u, 0x0Fu, 0x15u}) {
      int nsel = __builtin_popcount(mask);
      dim3 g2(nsel * ((n + 7) / 8), nmat);
      char nm[96];
      snprintf(nm, sizeof nm, "strip_fill_masked h=501 row-strip mask 0x%02x", mask);
      timeit(nm, [&] { strip_fill_masked<<<g2, 256>>>(buf, n, 3003, h, strips_r, mask, nsel, 1.0); }, bytes * nsel / 6.0);
    }
  }
  {
    const int panels = (n + 31) / 32;
    timeit("panel: fill kernel then window kernel (2 launches)", [&] {
      panel_fill_only<<<dim3(panels * 16, nmat), 256>>>(buf, n, 3003, 1.0);
      panel_window_only<<<dim3(panels * 8, nmat), 256>>>(buf, n, 3003, 1.0);
    }, bytes);
    timeit("panel: fill kernel alone", [&] { panel_fill_only<<<dim3(panels * 16, nmat), 256>>>(buf, n, 3003, 1.0); }, bytes * 2 / 3);
    timeit("panel: window kernel alone", [&] { panel_window_only<<<dim3(panels * 8, nmat), 256>>>(buf, n, 3003, 1.0); }, bytes / 3);
    timeit("panel: ONE kernel, items ordered by panel", [&] { panel_unified<<<dim3(panels * 24, nmat), 256>>>(buf, n, 3003, 1.0); }, bytes);
    timeit("panel: ONE kernel, all strips (upper bound)", [&] { panel_unified_strips<<<dim3(panels * 24, nmat), 256>>>(buf, n, 3003, 1.0); }, bytes);
  }
  timeit("tile_fill16 ld=3004", [&] { tile_fill16<<<grid, 256>>>(buf, n, 3004, tiles_r, 1.0); }, bytes);
  timeit("tile_fill16 ld=3008", [&] { tile_fill16<<<grid, 256>>>(buf, n, 3008, tiles_r, 1.0); }, bytes);
  {
    // 8 slots of 19 cliques each (1.33 GB per slot, like the stress workload): slot stride 166e6 doubles
    const size_t slot_stride = 166332179;  // sum |Ck|^2 at stress
    double* big; CK(cudaMalloc(&big, slot_stride * 8 * 8 + 1024));
    double b2 = 8.0 * n * (double)n * 8;
    timeit("slots-in-CTA x8, stride 1.33 GB (1 matrix)", [&] { tile_fill8_slots<<<dim3(tiles_r * tiles_c, 1), 256>>>(big, n, 3003, tiles_r, slot_stride, 8, 1.0); }, b2);
    timeit("slot-fastest grid x8, stride 1.33 GB", [&] { tile_fill8_slotfast<<<dim3(8, tiles_r * tiles_c), 256>>>(big, n, 3003, tiles_r, slot_stride, 1.0); }, b2);
    timeit("slot-slowest grid x8, stride 1.33 GB", [&] { tile_fill8<<<dim3(tiles_r * tiles_c, 1), 256>>>(big, n, 3003, tiles_r, 1.0); }, b2 / 8);
    cudaFree(big);
  }
  CK(cudaGetLastError());
  return 0;
}
