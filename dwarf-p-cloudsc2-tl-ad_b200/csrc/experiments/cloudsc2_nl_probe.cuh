// cloudsc2_nl_probe.cuh -- NOT PART OF THE PRODUCT LIBRARY.
// Measured-and-rejected variants of the nonlinear kernel and the instrumented (PROBE) builds of the
// default one, kept for the record of DESIGN.md 3.1 / 3.3.  Compiled only with -DCSC2_EXPERIMENTS
// (tools/probes/Makefile builds libcloudsc2_b200_experiments.so); __graft_entry__.build() does not
// define it, so the shipped libcloudsc2_b200.so holds one NL shape and none of this.
// Included by cloudsc2_nl_kernel.cu inside its anonymous namespace.
#pragma once

// PROBE: extra dummy instructions per level to measure what the kernel is sensitive to --
// 1: 64 integer-ALU ops, 2: 32 independent FP64 FMAs with a constant operand, 3: 64 FP32 FMAs.
template <int PROBE>
struct NlProbe {
  unsigned probe_i;
  double probe_d[4] = {1.0, 2.0, 3.0, 4.0};
  float probe_f[4] = {1.f, 2.f, 3.f, 4.f};
  __device__ __forceinline__ NlProbe() : probe_i(threadIdx.x) {}
  __device__ __forceinline__ void level(int jk) {
    if (PROBE == 1) {
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(probe_i) : "r"(jk), "r"(i));
    } else if (PROBE == 2) {
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(probe_d[i & 3]) : "d"(1.0000001));
    } else if (PROBE == 3) {
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(probe_f[i & 3]) : "f"(1.0000001f));
    }
  }
  // never true: keeps the dummy chains observable
  __device__ __forceinline__ bool fired() const {
    return PROBE != 0 && (probe_i == 0xdeadbeefu || probe_d[0] + probe_d[1] + probe_d[2] + probe_d[3] == 0.5 ||
                          probe_f[0] + probe_f[1] + probe_f[2] + probe_f[3] == 0.5f);
  }
};
