// Throughput probe: scalar FFMA / FADD / FMUL vs the sm_100 packed FFMA2 / FADD2 / FMUL2 (fma.rn.f32x2 ...), one CTA per SM,
// W warps per CTA, 8 independent accumulator chains per thread.  Prints warp-instructions per cycle per SM and FP32
// lane-operations per cycle per SM.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/f32x2_probe tools/f32x2_probe.cu && /tmp/f32x2_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float* out, long long* cycles, int iters) {
  float a[8], b = 1.0001f + threadIdx.x * 1e-7f, c = 0.5f;
  uint64_t A[8], B2, C2;
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(A[i]) : "f"(a[i]), "f"(a[i] + 1.f));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B2) : "f"(b), "f"(b));
  asm("mov.b64 %0, {%1, %2};" : "=l"(C2) : "f"(c), "f"(c));
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fmaf(a[i], b, c);
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(B2), "l"(C2));
      if (MODE == 2) a[i] = a[i] + b;
      if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B2));
      if (MODE == 4) a[i] = a[i] * b;
      if (MODE == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B2));
      if (MODE == 6) { a[i] = fmaf(a[i], b, c); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(i + 4) & 7])); }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(A[i])); s += a[i] + lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps, int lanes_per_instr) {
  int sms = 148, iters = 4096;
  float* out; long long* cyc;
  cudaMalloc(&out, sms * warps * 32 * sizeof(float));
  cudaMalloc(&cyc, sms * sizeof(long long));
  probe<MODE><<<sms, warps * 32>>>(out, cyc, iters);
  probe<MODE><<<sms, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
  double winstr = (double)iters * 8 * warps * (MODE == 6 ? 2 : 1);
  printf("%-28s warps/SM %2d: %.3f warp-instr/cycle/SM, %.1f fp32 lane-ops/cycle/SM (%.0f cycles)\n", name, warps, winstr / c,
         winstr * 32 * lanes_per_instr / c / (MODE == 6 ? 2 : 1), c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("FFMA  (scalar, 3-reg)", w, 1);
    run<1>("FFMA2 (fma.rn.f32x2)", w, 2);
    run<2>("FADD  (scalar)", w, 1);
    run<3>("FADD2 (add.rn.f32x2)", w, 2);
    run<4>("FMUL  (scalar)", w, 1);
    run<5>("FMUL2 (mul.rn.f32x2)", w, 2);
    run<6>("FFMA + MUFU.EX2 interleaved", w, 1);
  }
  return 0;
}
