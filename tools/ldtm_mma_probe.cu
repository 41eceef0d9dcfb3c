// How do tcgen05.ld (epilogue pulls) and TS-mode tcgen05.mma (A operand in tensor memory) share tensor memory?
// One CTA per SM: W reader warps loop on tcgen05.ld.32x32b.x32 + wait of accumulator columns; one more warp issues the
// production MMA pattern (three bf16 products per k-step, M = 128, N = 128 or 256) back to back.  Prints the readers'
// bytes per cycle per SM and the cycles per MMA, with and without the other side running.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I bgflow_b200/csrc -I include -o /tmp/ldtm_mma_probe tools/ldtm_mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "bgx_tc.cuh"

using namespace bgx::tc;

__global__ void probe(int readers, int run_readers, int run_mma, int n_cols, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  __shared__ volatile int done;
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;   // small bf16 values
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    done = 0;
  }
  if (warp == readers) tmem_alloc<512>(&tmem_slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == readers) {
    // MMA issuer: accumulator at column 0 (n_cols wide), A operand (two bf16 terms) at columns 384 / 448
    const uint32_t idesc = idesc_bf16(128, n_cols);
    const long long t0 = clock64();
    if (run_mma) {
      for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t d1 = smem_desc_sw128(smem_u32(base)) + 2 * ks, d2 = smem_desc_sw128(smem_u32(base) + 32768) + 2 * ks;
          mma3_bf16x3_elect(tmem, tmem + 384 + ks * 8, tmem + 448 + ks * 8, d1, d2, idesc, 1u);
        }
      }
      mma_commit_elect(&bar);
      mbar_wait(&bar, 0, nullptr);
    } else {
      while (clock64() - t0 < 400000) {}
    }
    if (lane == 0) {
      out[blockIdx.x * 4 + 0] = clock64() - t0;
      done = 1;
    }
    __syncwarp();
  } else if (run_readers) {
    const uint32_t lb = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t sink = 0;
    long long n = 0;
    const long long t0 = clock64();
    while (!done) {
      uint32_t r[32];
      tmem_ld32(tmem + lb + ((warp >> 2) * 32) % 96, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) sink ^= r[i];      // (static indices: the registers must not spill to local memory)
      ++n;
    }
    const long long t1 = clock64();
    if (sink == 0x12345u) out[3] = 1;
    if (lane == 0) {
      atomicAdd((unsigned long long*)&out[blockIdx.x * 4 + 1], (unsigned long long)n);
      if (warp == 0) out[blockIdx.x * 4 + 2] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == readers) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

int main() {
  const int sms = 148, reps = 512;
  long long* out;
  cudaMalloc(&out, sms * 4 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  for (int n_cols : {128, 256})
    for (int readers : {4, 8, 16})
      for (int mode = 0; mode < 3; ++mode) {     // 0: MMAs alone, 1: readers alone, 2: both
        const int run_mma = mode != 1, run_rd = mode != 0;
        if (mode == 1 && n_cols == 256) continue;
        for (int pass = 0; pass < 2; ++pass) {
          cudaMemset(out, 0, sms * 4 * sizeof(long long));
          probe<<<sms, (readers + 1) * 32, 70 * 1024>>>(readers, run_rd, run_mma, n_cols, reps, out);
        }
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148 * 4];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double mc = 0, nl = 0, rc = 0;
        for (int i = 0; i < sms; ++i) { mc += h[4 * i]; nl += h[4 * i + 1]; rc += h[4 * i + 2]; }
        mc /= sms; nl /= sms; rc /= sms;
        printf("N=%3d readers %2d %-13s: ", n_cols, readers, mode == 0 ? "MMAs alone" : mode == 1 ? "readers alone" : "both");
        if (run_mma) printf("%6.1f cycles/MMA  ", mc / (reps * 12.0));
        if (run_rd) printf("tcgen05.ld %7.1f B/cycle/SM (%5.0f cycles per x32 load + wait per warp)", nl * 4096.0 / rc, rc * readers / nl);
        printf("  (%s)\n", cudaGetErrorString(e));
      }
  return 0;
}
