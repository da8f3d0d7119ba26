// Developer micro-benchmark: dependent-issue latency and per-SMSP throughput of the FP64 pipe and
// the MUFU operations the kernel's chains are built from (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void rsq_chain(double* out, long long* cyc, int iters) {
  double x = 1.0 + threadIdx.x * 1e-3;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double y;
      asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
      x = y;
    }
  }
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void lg2ex2_chain(float* out, long long* cyc, int iters) {
  float x = 1.5f + threadIdx.x * 1e-3f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) x = exp2f(__log2f(x) * 0.999f);
  }
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void cvt_chain(double* out, long long* cyc, int iters) {
  double x = 1.5 + threadIdx.x * 1e-3;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) x = (double)__double2float_rn(x) + 1e-9;
  }
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 4096;
#define RUN(name, ...)                                                      \
  __VA_ARGS__; cudaDeviceSynchronize(); __VA_ARGS__; cudaDeviceSynchronize(); \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // one warp on one SM: latency (ILP 1) and issue-limited throughput (ILP 8)
  RUN("", dfma_chain<1><<<1, 32>>>(out, cyc, iters, 0.999, 1e-3)); printf("DFMA dependent, 1 warp:          %.2f cycles/DFMA\n", (double)h / (iters * 16));
  RUN("", dfma_chain<2><<<1, 32>>>(out, cyc, iters, 0.999, 1e-3)); printf("DFMA ILP2, 1 warp:               %.2f cycles/DFMA\n", (double)h / (iters * 32));
  RUN("", dfma_chain<4><<<1, 32>>>(out, cyc, iters, 0.999, 1e-3)); printf("DFMA ILP4, 1 warp:               %.2f cycles/DFMA\n", (double)h / (iters * 64));
  RUN("", dfma_chain<8><<<1, 32>>>(out, cyc, iters, 0.999, 1e-3)); printf("DFMA ILP8, 1 warp:               %.2f cycles/DFMA\n", (double)h / (iters * 128));
  // 4 warps on one SMSP each (128 threads = 4 warps, one per SMSP), then 16 warps (4 per SMSP)
  RUN("", dfma_chain<1><<<1, 512>>>(out, cyc, iters, 0.999, 1e-3)); printf("DFMA dependent, 4 warps/SMSP:    %.2f cycles/DFMA/warp  (pipe: %.2f cycles per warp-instr per SMSP)\n", (double)h / (iters * 16), (double)h / (iters * 16) / 4);
  RUN("", dfma_chain<2><<<1, 512>>>(out, cyc, iters, 0.999, 1e-3)); printf("DFMA ILP2, 4 warps/SMSP:         %.2f cycles per warp-instr per SMSP\n", (double)h / (iters * 32) / 4);
  RUN("", dfma_chain<1><<<1, 1024>>>(out, cyc, iters, 0.999, 1e-3)); printf("DFMA dependent, 8 warps/SMSP:    %.2f cycles per warp-instr per SMSP\n", (double)h / (iters * 16) / 8);
  RUN("", rsq_chain<<<1, 32>>>(out, cyc, iters)); printf("MUFU.RSQ64H dependent:           %.2f cycles\n", (double)h / (iters * 16));
  RUN("", lg2ex2_chain<<<1, 32>>>((float*)out, cyc, iters)); printf("LG2+FMUL+EX2 dependent:          %.2f cycles\n", (double)h / (iters * 16));
  RUN("", cvt_chain<<<1, 32>>>(out, cyc, iters)); printf("F2F.F32.F64+F2F.F64.F32+DADD:    %.2f cycles\n", (double)h / (iters * 16));
  return 0;
}
