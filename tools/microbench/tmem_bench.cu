// TMEM load/store throughput on sm_100a: N warps of one CTA per SM each issue back-to-back tcgen05.ld / tcgen05.st
// of 32 lanes x 32 columns (4 KB per warp instruction); reports bytes per clock per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/tmem_bench tools/microbench/tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}

// mode 0: ld with a wait after every load; 1: 4 loads in flight then wait; 2: stores
__global__ void k(uint32_t* out, long long* cycles, int mode) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t r[4][32];
  for (int j = 0; j < 4; ++j) for (int i = 0; i < 32; ++i) r[j][i] = threadIdx.x + i + j;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {
    for (int it = 0; it < ITERS; ++it) {
      ld32(base + (it & 3) * 32, r[0]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0][it & 31];
    }
  } else if (mode == 1) {
    for (int it = 0; it < ITERS; it += 4) {
      ld32(base, r[0]); ld32(base + 32, r[1]); ld32(base + 64, r[2]); ld32(base + 96, r[3]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0][it & 31] + r[1][it & 31] + r[2][it & 31] + r[3][it & 31];
    }
  } else {
    for (int it = 0; it < ITERS; it += 4) {
      st32(base, r[0]); st32(base + 32, r[1]); st32(base + 64, r[2]); st32(base + 96, r[3]);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* out; long long* cyc; cudaMalloc(&out, sms * 512 * 4); cudaMalloc(&cyc, sms * 8);
  const char* names[3] = {"ld x32, wait each", "ld x32, 4 in flight", "st x32, 4 in flight"};
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 2, 4, 8, 16}) {
      k<<<sms, warps * 32>>>(out, cyc, mode); cudaDeviceSynchronize();
      k<<<sms, warps * 32>>>(out, cyc, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)warps * ITERS * 4096.0;
      printf("{\"bench\":\"tmem %s\",\"warps\":%d,\"cycles\":%lld,\"bytes_per_clk_per_sm\":%.1f,\"cycles_per_warp_instr\":%.1f}\n",
             names[mode], warps, c, bytes / c, (double)c / ITERS);
    }
  return 0;
}
