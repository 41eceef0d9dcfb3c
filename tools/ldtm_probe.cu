// tcgen05.ld throughput probe: one CTA per SM, W warps (warp w reads TMEM lane quadrant w % 4), every iteration issues
// DEPTH loads of SHAPE columns (32x32b.x32 / .x16 / .x8) before one tcgen05.wait::ld.  Prints bytes per cycle per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ldtm_probe tools/ldtm_probe.cu && /tmp/ldtm_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int COLS>
__device__ __forceinline__ uint32_t ld_cols(uint32_t taddr) {
  uint32_t acc = 0;
  if (COLS == 32) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r[i];
  } else if (COLS == 16) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[i];
  } else {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= r[i];
  }
  return acc;
}

template <int COLS, int DEPTH>
__global__ void probe(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) acc ^= ld_cols<COLS>(base + (uint32_t)(((warp >> 2) * DEPTH + d) * COLS) % (512 - COLS));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_slot) : "memory");
}

template <int COLS, int DEPTH>
void run(int warps) {
  const int sms = 148, iters = 2000;
  long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, sms * sizeof(long long));
  cudaMalloc(&sink, sms * warps * 32 * sizeof(uint32_t));
  probe<COLS, DEPTH><<<sms, warps * 32>>>(iters, cyc, sink);
  probe<COLS, DEPTH><<<sms, warps * 32>>>(iters, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < sms; ++i) c += h[i];
  c /= sms;
  const double bytes = (double)iters * DEPTH * COLS * 32 * 4 * warps;
  printf("x%-2d depth %d warps %2d: %7.1f B/cycle/SM, %7.0f cycles per iteration (%s)\n", COLS, DEPTH, warps, bytes / c, c / iters,
         cudaGetErrorString(e));
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  for (int w : {1, 4, 8, 16}) {
    run<32, 1>(w);
    run<32, 2>(w);
    run<32, 4>(w);
    run<16, 2>(w);
    run<16, 4>(w);
    run<8, 4>(w);
  }
  return 0;
}
