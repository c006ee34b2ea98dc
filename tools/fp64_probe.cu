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
  float fa[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) fa[i] = 0.f;
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
        // modes 8-10: mode 0's DFMA stream plus a double->float conversion of the result
        if (MODE == 8) {  // one F2F.F32.F64 (+ FADD) per DFMA
          a[i] = fma(a[i], kb, kc);
          fa[i] += __double2float_rn(a[i]);
        }
        if (MODE == 9) {  // one F2F per 4 DFMA
          a[i] = fma(a[i], kb, kc);
          if ((i & 3) == 0) fa[i] += __double2float_rn(a[i]);
        }
        if (MODE == 10) {  // conversion by integer ops on the high word (truncating, normal range)
          a[i] = fma(a[i], kb, kc);
          int hi = __double2hiint(a[i]);
          int fb = (hi & 0x80000000) | (((hi & 0x7fffffff) - (896 << 20)) << 3);
          fa[i] += __int_as_float(fb);
        }
        if (MODE == 11) {  // F2F only (a[] is refreshed by an integer xor): conversion throughput
          a[i] = __hiloint2double(__double2hiint(a[i]) ^ (it & 1), __double2loint(a[i]));
          fa[i] += __double2float_rn(a[i]);
        }
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i] + b[i] + c[i] + fa[i];
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
  run<8>(p.multiProcessorCount, sink);
  run<9>(p.multiProcessorCount, sink);
  run<10>(p.multiProcessorCount, sink);
  run<11>(p.multiProcessorCount, sink);
  return 0;
}
