// Throughput of the exp2 paths available to the attention softmax on sm_100a (per SM, per clock):
//   ex2.approx.ftz.f32 | ex2.approx.ftz.bf16x2 | ex2.approx.ftz.f16x2 | degree-3 polynomial exp2 on the FMA pipe
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/mufu_bench tools/microbench/mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

constexpr int ITERS = 4096;
constexpr int UNROLL = 8;

__global__ void k_ex2_f32(float* out, float seed) {
  float x[UNROLL];
  for (int i = 0; i < UNROLL; ++i) x[i] = seed - i * 0.01f - threadIdx.x * 1e-4f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  float s = 0; for (int i = 0; i < UNROLL; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ex2_bf16x2(uint32_t* out, uint32_t seed) {
  uint32_t x[UNROLL];
  for (int i = 0; i < UNROLL; ++i) x[i] = seed + i * 3 + threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
  }
  uint32_t s = 0; for (int i = 0; i < UNROLL; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ex2_f16x2(uint32_t* out, uint32_t seed) {
  uint32_t x[UNROLL];
  for (int i = 0; i < UNROLL; ++i) x[i] = seed + i * 3 + threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
  }
  uint32_t s = 0; for (int i = 0; i < UNROLL; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// Cody-Waite: 2^x = 2^floor(x) * p(frac), degree-3 p, exponent inserted with integer add (FA4-style emulation)
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.0f);
  const float fl = floorf(x);
  const float f = x - fl;
  float p = 0.0555054f;            // minimax-ish coefficients for 2^f on [0,1)
  p = fmaf(p, f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + ((int)fl << 23));
}
__global__ void k_poly(float* out, float seed) {
  float x[UNROLL];
  for (int i = 0; i < UNROLL; ++i) x[i] = seed - i * 0.37f - threadIdx.x * 1e-3f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) x[i] = poly_exp2(x[i]) - 1.5f;
  }
  float s = 0; for (int i = 0; i < UNROLL; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: half MUFU, half polynomial (what the softmax would issue)
__global__ void k_mixed(float* out, float seed) {
  float x[UNROLL];
  for (int i = 0; i < UNROLL; ++i) x[i] = seed - i * 0.37f - threadIdx.x * 1e-3f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; i += 2) {
      x[i] = poly_exp2(x[i]) - 1.5f;
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i + 1]));
      x[i + 1] -= 1.5f;
    }
  }
  float s = 0; for (int i = 0; i < UNROLL; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char* name, F launch, double ops_per_thread_iter) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  const double threads = (double)sms * 4 * 512;
  const double total = threads * ITERS * UNROLL * ops_per_thread_iter;
  printf("{\"bench\":\"%s\",\"ms\":%.3f,\"results_per_ns_per_sm\":%.2f,\"results_per_clk_per_sm_at_max_clock\":%.2f}\n", name, ms,
         total / (ms * 1e6) / sms, total / (ms * 1e-3) / sms / (khz * 1e3));
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* buf; cudaMalloc(&buf, (size_t)sms * 4 * 512 * 4);
  dim3 grid(sms * 4), block(512);
  run("ex2.f32", [&] { k_ex2_f32<<<grid, block>>>((float*)buf, -0.5f); }, 1);
  run("ex2.bf16x2 (2 results/op)", [&] { k_ex2_bf16x2<<<grid, block>>>((uint32_t*)buf, 0xbf00bf00u); }, 2);
  run("ex2.f16x2 (2 results/op)", [&] { k_ex2_f16x2<<<grid, block>>>((uint32_t*)buf, 0xb800b800u); }, 2);
  run("poly3 exp2 on FMA/ALU pipes", [&] { k_poly<<<grid, block>>>((float*)buf, -0.5f); }, 1);
  run("mixed 50% MUFU / 50% poly", [&] { k_mixed<<<grid, block>>>((float*)buf, -0.5f); }, 1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
