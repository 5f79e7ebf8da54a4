// Microbenchmarks behind the host-gather design ("end to end"): how fast can dense clique
// blocks reach caller-owned host memory?
//   (a) cudaMemcpyAsync D2H, contiguous                      -- the current ring drain
//   (b) cudaMemcpy2DAsync D2H of 1000 x 1000 sub-blocks, pitch 3003 doubles
//   (c) SM stores straight into mapped pinned host memory (zero copy), 128 x 32 tiles, ld = 3003
//   (d) host memset with 1..N threads                        -- zero-filling structural zeros on the CPU
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xcompiler -pthread -o tools/probe_host_xfer tools/probe_host_xfer.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void tile_fill8(double* out, int n, int ld, int tiles_r, double v) {
  const int tr = threadIdx.x & 127, cg = threadIdx.x >> 7;
  const int r0 = (blockIdx.x % tiles_r) * 128, c0 = (blockIdx.x / tiles_r) * 32;
  double* o = out + (size_t)blockIdx.y * (size_t)ld * n;
  const int r = r0 + tr;
  if (r >= n) return;
  for (int c = cg; c < 32; c += 2) if (c0 + c < n) o[r + (size_t)(c0 + c) * ld] = v;
}
// only a sub-rectangle [r_lo, r_hi) x [c_lo, c_hi) of every matrix (the non-zero part of a clique block)
__global__ void tile_fill8_rect(double* out, int n, int ld, int r_lo, int r_hi, int c_lo, int c_hi, double v) {
  const int tiles_r = (r_hi - r_lo + 127) / 128;
  const int tr = threadIdx.x & 127, cg = threadIdx.x >> 7;
  const int r0 = r_lo + (blockIdx.x % tiles_r) * 128, c0 = c_lo + (blockIdx.x / tiles_r) * 32;
  double* o = out + (size_t)blockIdx.y * (size_t)ld * n;
  const int r = r0 + tr;
  if (r >= r_hi) return;
  for (int c = cg; c < 32; c += 2) if (c0 + c < c_hi) o[r + (size_t)(c0 + c) * ld] = v;
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  const int n = 3003, nmat = 32;
  const size_t per = (size_t)n * n, total = per * nmat;  // 2.3 GB
  double *dev, *host;
  CK(cudaMalloc(&dev, total * 8));
  CK(cudaHostAlloc(&host, total * 8, cudaHostAllocPortable | cudaHostAllocMapped));
  CK(cudaMemset(dev, 0, total * 8));
  memset(host, 0, total * 8);
  double* hmap; CK(cudaHostGetDevicePointer(&hmap, host, 0));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  auto wall = [&](const char* name, auto fn, double bytes) {
    fn(); CK(cudaStreamSynchronize(st));
    const int reps = 3;
    const double t0 = now();
    for (int i = 0; i < reps; ++i) fn();
    CK(cudaStreamSynchronize(st));
    const double dt = (now() - t0) / reps;
    printf("%-58s %9.3f ms  %8.1f GB/s\n", name, dt * 1e3, bytes / dt / 1e9);
    fflush(stdout);
  };
  wall("(a) cudaMemcpyAsync D2H contiguous 2.3 GB", [&] { CK(cudaMemcpyAsync(host, dev, total * 8, cudaMemcpyDeviceToHost, st)); }, total * 8.0);
  wall("(b) cudaMemcpy2DAsync D2H 1000x1000 rects, pitch 3003 (x4 per matrix)", [&] {
    for (int m = 0; m < nmat; ++m)
      for (int k = 0; k < 4; ++k) {
        const size_t off = m * per + (size_t)(k & 1) * 1000 + (size_t)(k >> 1) * 1000 * n;
        CK(cudaMemcpy2DAsync(host + off, (size_t)n * 8, dev + off, (size_t)n * 8, 1000 * 8, 1000, cudaMemcpyDeviceToHost, st));
      }
  }, 4.0 * nmat * 1e6 * 8);
  wall("(b2) cudaMemcpy2DAsync D2H 128x32 tiles (x200 per matrix)", [&] {
    for (int m = 0; m < nmat; ++m)
      for (int k = 0; k < 200; ++k) {
        const size_t off = m * per + (size_t)(k % 20) * 128 + (size_t)(k / 20) * 32 * n;
        CK(cudaMemcpy2DAsync(host + off, (size_t)n * 8, dev + off, (size_t)n * 8, 128 * 8, 32, cudaMemcpyDeviceToHost, st));
      }
  }, 200.0 * nmat * 4096 * 8);
  const int tiles_r = (n + 127) / 128, tiles_c = (n + 31) / 32;
  wall("(c) SM stores to mapped host memory, full 3003^2, ld 3003", [&] { tile_fill8<<<dim3(tiles_r * tiles_c, nmat), 256, 0, st>>>(hmap, n, n, tiles_r, 1.0); }, total * 8.0);
  wall("(c2) SM stores to mapped host, 2002x2002 rect of each matrix", [&] {
    tile_fill8_rect<<<dim3(((2002 + 127) / 128) * ((2002 + 31) / 32), nmat), 256, 0, st>>>(hmap, n, n, 0, 2002, 0, 2002, 2.0);
  }, (double)nmat * 2002 * 2002 * 8);
  wall("(c3) SM stores to device memory, same rect (reference)", [&] {
    tile_fill8_rect<<<dim3(((2002 + 127) / 128) * ((2002 + 31) / 32), nmat), 256, 0, st>>>(dev, n, n, 0, 2002, 0, 2002, 2.0);
  }, (double)nmat * 2002 * 2002 * 8);
  CK(cudaGetLastError());
  // (d) host memset bandwidth
  const unsigned hw = std::thread::hardware_concurrency();
  for (unsigned nt : {1u, 2u, 4u, 8u, 16u, 32u}) {
    if (nt > hw) break;
    const size_t chunk = total * 8 / nt;
    const double t0 = now();
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t] { memset((char*)host + t * chunk, 0, chunk); });
    for (auto& x : th) x.join();
    const double dt = now() - t0;
    printf("(d) host memset %2u threads                                  %9.3f ms  %8.1f GB/s\n", nt, dt * 1e3, total * 8.0 / dt / 1e9);
  }
  // (e) memset on 8 threads while a D2H copy runs
  {
    CK(cudaMemcpyAsync(host, dev, total * 4, cudaMemcpyDeviceToHost, st));
    const double t0 = now();
    std::vector<std::thread> th;
    const size_t chunk = total * 4 / 8;
    for (unsigned t = 0; t < 8; ++t) th.emplace_back([&, t] { memset((char*)host + total * 4 + t * chunk, 0, chunk); });
    for (auto& x : th) x.join();
    const double dtm = now() - t0;
    CK(cudaStreamSynchronize(st));
    const double dtc = now() - t0;
    printf("(e) concurrently: memset 8 thr %.1f GB/s, D2H %.1f GB/s\n", total * 4.0 / dtm / 1e9, total * 4.0 / dtc / 1e9);
  }
  printf("hardware_concurrency %u\n", hw);
  return 0;
}
