// tools/fp64_probe.cu — what bounds an FP64 instruction stream on B200 besides the 2-cycle
// FP64 pipe: register-operand bandwidth.  Each mode issues the same number of FP64
// instructions per thread, differing in how many DISTINCT 64-bit register operands each
// instruction reads (and whether consecutive instructions share one in the same slot).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(256) probe(int iters, double s, double *sink) {
  double a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = s + i + threadIdx.x * 1e-9;
    b[i] = 1.0 + (s + i) * 1e-9;
    c[i] = 1e-9 * (i + 1) + s * 1e-12;
  }
  const double kb = 1.0000001 + s * 1e-12, kc = 1e-9 + s * 1e-15;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) a[i] = fma(a[i], kb, kc);             // 1 distinct + 2 loop constants
        if (MODE == 1) a[i] = fma(b[i], c[i], a[i]);         // 3 distinct, nothing shared
        if (MODE == 2) a[i] = fma(b[i], kc, a[i]);           // 2 distinct + 1 shared
        if (MODE == 3) a[i] = a[i] + b[i];                   // DADD, 2 distinct
        if (MODE == 4) a[i] = a[i] * b[i];                   // DMUL, 2 distinct
        if (MODE == 5) a[i] = fma(b[i], b[i], a[i]);         // 2 distinct (one twice)
        if (MODE == 6) a[i] = fma(b[i], c[(i + u) & 7], a[i]);  // 3 distinct, rotating
        if (MODE == 7) a[i] = fma(kb, b[i], a[i]);           // shared in slot A
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i] + b[i] + c[i];
  if (r == 12345.678) sink[0] = r;
}

template <int MODE> void run(int sms, double *sink) {
  const int iters = 20000, blocks = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<MODE><<<blocks, 256>>>(64, 1.0, sink);
  cudaEventRecord(e0);
  probe<MODE><<<blocks, 256>>>(iters, 1.0, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double inst = 32.0 * iters * blocks * 256.0;  // FP64 thread-instructions
  double per_sm_clk = inst / (ms * 1e-3) / sms / 1.965e9;
  printf("mode %d: %8.3f ms  %6.2f T inst/s  %5.1f lanes/clk/SM @1965MHz (64 = pipe peak)\n", MODE,
         ms, inst / (ms * 1e-3) / 1e12, per_sm_clk);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double *sink;
  cudaMalloc(&sink, 8);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<0>(p.multiProcessorCount, sink);
  run<1>(p.multiProcessorCount, sink);
  run<2>(p.multiProcessorCount, sink);
  run<3>(p.multiProcessorCount, sink);
  run<4>(p.multiProcessorCount, sink);
  run<5>(p.multiProcessorCount, sink);
  run<6>(p.multiProcessorCount, sink);
  run<7>(p.multiProcessorCount, sink);
  return 0;
}
