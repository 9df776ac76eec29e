// issue_probe.cu -- does an FP64 instruction occupy the sub-partition's issue port for 1 or 2 cycles?
// Streams of independent DFMA mixed with K independent 32-bit integer (or FP32) instructions per
// DFMA.  If a DFMA blocks issue for 2 cycles the group costs 2+K cycles, otherwise max(2, 1+K).
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int MODE>
__global__ void k(double *out, int iters, double x, int seed) {
  double a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = x + i * 1e-3 + threadIdx.x * 1e-6;
  const double b = x * 1.0000001, c = x * 1e-9;
  unsigned u[8];
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { u[i] = seed + i + threadIdx.x; f[i] = (float)(seed + i); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = fma(a[i], b, c);
#pragma unroll
        for (int j = 0; j < K; ++j) {
          if (MODE == 0) u[(i * K + j) & 7] = u[(i * K + j) & 7] * 3u + 7u;          // IMAD
          if (MODE == 1) f[(i * K + j) & 7] = fmaf(f[(i * K + j) & 7], 1.0001f, 0.5f);  // FFMA
        }
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += u[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}

template <int K, int MODE>
void run(const char *name) {
  const int warps = 4, threads = warps * 4 * 32, iters = 1000;
  double *d;
  cudaMalloc(&d, sizeof(double) * 148 * threads);
  k<K, MODE><<<148, threads>>>(d, iters, 1.0, 3);
  cudaDeviceSynchronize();
  k<K, MODE><<<148, threads>>>(d, iters, 1.0, 3);
  cudaDeviceSynchronize();
  double cyc;
  cudaMemcpy(&cyc, d, sizeof(double), cudaMemcpyDeviceToHost);
  const double groups = (double)iters * 8 * 4 * warps;    // DFMA groups per SMSP
  printf("%-6s K=%d : %.2f cycles per (DFMA + %d other) group per SMSP\n", name, K, cyc / groups, K);
  cudaFree(d);
}

int main() {
  run<0, 0>("IMAD"); run<1, 0>("IMAD"); run<2, 0>("IMAD"); run<3, 0>("IMAD"); run<4, 0>("IMAD");
  run<1, 1>("FFMA"); run<2, 1>("FFMA"); run<3, 1>("FFMA");
  return 0;
}
