// mufu_probe.cu -- throughput of MUFU.RCP64H / MUFU.RSQ64H (the seeds of csc2_rcp / csc2_sqrt) and of
// MUFU.EX2 per sub-partition, alone and mixed with DFMA.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, int iters, double x) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = x + i * 1e-3 + threadIdx.x * 1e-6;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double y;
        if (MODE == 0) { asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[i])); a[i] = y; }
        if (MODE == 1) { asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[i])); a[i] = y; }
        if (MODE == 2) { asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[i])); a[i] = fma(a[i], 1.0000001, y); }
        if (MODE == 3) { asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[i]));
                         double e = fma(-a[i], y, 1.0); e = fma(e, e, e); a[i] = fma(y, e, y) + 1.0; }
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int MODE>
void run(const char *name, int warps, int per_group) {
  const int threads = warps * 4 * 32, iters = 500;
  double *d;
  cudaMalloc(&d, sizeof(double) * 148 * threads);
  k<MODE><<<148, threads>>>(d, iters, 1.5);
  cudaDeviceSynchronize();
  k<MODE><<<148, threads>>>(d, iters, 1.5);
  cudaDeviceSynchronize();
  double cyc;
  cudaMemcpy(&cyc, d, sizeof(double), cudaMemcpyDeviceToHost);
  const double groups = (double)iters * 4 * 8 * warps;
  printf("%-34s warps/SMSP %d : %.2f cycles per group per SMSP (%d FP64-pipe instr per group)\n", name, warps,
         cyc / groups, per_group);
  cudaFree(d);
}
int main() {
  for (int w : {1, 3, 4}) {
    run<0>("MUFU.RCP64H", w, 0);
    run<1>("MUFU.RSQ64H", w, 0);
    run<2>("MUFU.RCP64H + 1 DFMA", w, 1);
    run<3>("csc2_rcp + 1 DADD (4 FP64)", w, 4);
  }
  return 0;
}
