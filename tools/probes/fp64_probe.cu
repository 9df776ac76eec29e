// fp64_probe.cu -- what FP64 issue rate can an sm_100a sub-partition actually sustain?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu && ./fp64_probe
// Measures warp-instructions per cycle per sub-partition for streams of independent DFMA with
// (a) three register operands, (b) two register operands + one constant-bank operand,
// (c) DMUL/DADD with two register operands, at 1..8 warps per sub-partition and ILP 1..8.
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double kc[4] = {1.0000001, 0.9999999, 1e-9, -1e-9};

template <int ILP, int MODE>
__global__ void k(double *out, int iters, double x, double y) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = x + i * 1e-3 + threadIdx.x * 1e-6;
  double b = x * 1.0000001, c = y * 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (MODE == 0) a[i] = fma(a[i], b, c);            // 3 register operands
        if (MODE == 1) a[i] = fma(a[i], kc[0], c);        // 2 registers + constant bank
        if (MODE == 2) a[i] = fma(a[i], kc[0], kc[2]);    // 1 register + 2 constant... (one const allowed) 
        if (MODE == 3) a[i] = a[i] * b;                   // DMUL 2 registers
        if (MODE == 4) a[i] = a[i] + c;                   // DADD 2 registers
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}

template <int ILP, int MODE>
void run(const char *name, int warps_per_smsp) {
  const int threads = warps_per_smsp * 4 * 32;     // one CTA per SM, warps spread over the 4 SMSPs
  const int iters = 2000;
  double *d;
  cudaMalloc(&d, sizeof(double) * 148 * threads);
  k<ILP, MODE><<<148, threads>>>(d, iters, 1.0, 1.0);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<ILP, MODE><<<148, threads>>>(d, iters, 1.0, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  double cyc;
  cudaMemcpy(&cyc, d, sizeof(double), cudaMemcpyDeviceToHost);
  const double inst = (double)iters * 8 * ILP * warps_per_smsp;   // warp instructions per SMSP
  printf("%-28s warps/SMSP %d ILP %d : %.3f warp-inst/cycle/SMSP (%.0f cycles)\n", name,
         warps_per_smsp, ILP, inst / cyc, cyc);
  cudaFree(d);
}

int main() {
  printf("--- DFMA r,r,r : warps per SMSP x ILP (independent chains per warp)\n");
  for (int w : {1, 2, 3, 4, 6, 8, 12, 16}) {
    run<1, 0>("DFMA r,r,r", w); run<2, 0>("DFMA r,r,r", w); run<4, 0>("DFMA r,r,r", w);
  }
  printf("--- DFMA r,c[],r\n");
  for (int w : {2, 3, 4, 8}) { run<1, 1>("DFMA r,c[],r", w); run<2, 1>("DFMA r,c[],r", w); }
  printf("--- DMUL / DADD r,r\n");
  for (int w : {3, 4}) { run<2, 3>("DMUL r,r", w); run<2, 4>("DADD r,r", w); }
  return 0;
}
