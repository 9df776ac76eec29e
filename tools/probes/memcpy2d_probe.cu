// memcpy2d_probe.cu -- is a pitched (2D) pinned-host -> device copy slower than a contiguous one?
// Shapes of the NL host path: PCLV rows of 2 of 5 slabs, B_CML rows of 1 and 3 of 8 slabs (n2 = 128*137 doubles).
#include <cstdio>
#include <cuda_runtime.h>
int main() {
  const size_t n2 = 128 * 137 * sizeof(double), rows = 1280;
  char *h, *d;
  cudaMallocHost(&h, 8 * n2 * rows);
  cudaMalloc(&d, 8 * n2 * rows);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](const char *name, size_t bytes, auto fn) {
    fn(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) fn();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-46s %.1f GB/s\n", name, bytes * 5 / (ms * 1e6));
  };
  time("1D contiguous, 2*n2*rows", 2 * n2 * rows, [&] { cudaMemcpyAsync(d, h, 2 * n2 * rows, cudaMemcpyHostToDevice, 0); });
  time("2D width 2*n2 of pitch 5*n2 (PCLV)", 2 * n2 * rows, [&] { cudaMemcpy2DAsync(d, 2 * n2, h, 5 * n2, 2 * n2, rows, cudaMemcpyHostToDevice, 0); });
  time("2D width 1*n2 of pitch 8*n2 (B_CML T)", n2 * rows, [&] { cudaMemcpy2DAsync(d, 4 * n2, h, 8 * n2, n2, rows, cudaMemcpyHostToDevice, 0); });
  time("2D width 3*n2 of pitch 8*n2 (B_CML Q,QL,QI)", 3 * n2 * rows, [&] { cudaMemcpy2DAsync(d, 4 * n2, h, 8 * n2, 3 * n2, rows, cudaMemcpyHostToDevice, 0); });
  time("D2H 2D width 3*n2 into pitch 8*n2 (B_LOC)", 3 * n2 * rows, [&] { cudaMemcpy2DAsync(h, 8 * n2, d, 5 * n2, 3 * n2, rows, cudaMemcpyDeviceToHost, 0); });
  time("D2H 1D contiguous 3*n2*rows", 3 * n2 * rows, [&] { cudaMemcpyAsync(h, d, 3 * n2 * rows, cudaMemcpyDeviceToHost, 0); });
  return 0;
}
