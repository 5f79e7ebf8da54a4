// Measured FP64 GEMM throughput of this GPU (cuBLAS DGEMM, which runs on the DMMA tensor path): the
// denominator for the Gram kernel's executed flops.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_dgemm tools/probe_dgemm.cu -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
int main() {
  cublasHandle_t h; cublasCreate(&h);
  for (int n : {2048, 4096, 8192}) {
    double *A, *B, *C;
    CK(cudaMalloc(&A, (size_t)n * n * 8)); CK(cudaMalloc(&B, (size_t)n * n * 8)); CK(cudaMalloc(&C, (size_t)n * n * 8));
    CK(cudaMemset(A, 0, (size_t)n * n * 8)); CK(cudaMemset(B, 0, (size_t)n * n * 8));
    const double one = 1.0, zero = 0.0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("cublasDgemm n = %5d  %8.3f ms  %7.2f TFLOP/s\n", n, best, 2.0 * n * (double)n * n / (best * 1e-3) / 1e12);
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  // a rank-k update of the Gram kernel's shape: C (1000 x 1000) = A' A with k = 442 active neurons, 53 of them
  {
    const int n = 1000, k = 442, batch = 53;
    double *A, *C; CK(cudaMalloc(&A, (size_t)k * n * 8 * batch)); CK(cudaMalloc(&C, (size_t)n * n * 8 * batch));
    CK(cudaMemset(A, 0, (size_t)k * n * 8 * batch));
    const double one = 1.0, zero = 0.0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      cublasDgemmStridedBatched(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, k, &one, A, k, (long long)k * n, A, k, (long long)k * n, &zero, C, n, (long long)n * n, batch);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("cublasDgemmStridedBatched 53 x (1000 x 1000 x 442)  %8.3f ms  %7.2f TFLOP/s (full square; the Gram kernel computes the upper half)\n",
           best, 2.0 * n * (double)n * k * batch / (best * 1e-3) / 1e12);
  }
  return 0;
}
