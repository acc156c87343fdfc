// How MUFU.EX2, packed fp32x2 FMA, scalar FMA, ALU (FMNMX) and F2FP instructions share one SM sub-partition on sm_100a
// when only W warps per scheduler are resident (the attention softmax has 2): cycles per loop iteration for a given mix of
// INDEPENDENT instructions, measured with clock64 inside the kernel.  One CTA of 128*W threads per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/issue_mix_bench tools/microbench/issue_mix_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

template <int NM, int NF2, int NF1, int NALU, int NCVT>
__global__ void mix_kernel(float* out, long long* cycles, float seed) {
  float m[NM > 0 ? NM : 1];
  float2 f2[NF2 > 0 ? NF2 : 1];
  float f1[NF1 > 0 ? NF1 : 1];
  float a[NALU > 0 ? NALU : 1];
  float c[NCVT > 0 ? 2 * NCVT : 1];
  uint32_t pk[NCVT > 0 ? NCVT : 1];
  for (int i = 0; i < NM; ++i) m[i] = seed - i * 0.01f - threadIdx.x * 1e-4f;
  for (int i = 0; i < NF2; ++i) f2[i] = make_float2(seed + i, seed - i);
  for (int i = 0; i < NF1; ++i) f1[i] = seed + 0.5f * i;
  for (int i = 0; i < NALU; ++i) a[i] = seed - 0.25f * i;
  for (int i = 0; i < 2 * NCVT; ++i) c[i] = seed + 0.125f * i;
  for (int i = 0; i < NCVT; ++i) pk[i] = 0;
  const float2 k2 = make_float2(0.999f, 1.001f), b2 = make_float2(1e-3f, -1e-3f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NM; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
#pragma unroll
    for (int i = 0; i < NF2; ++i) {
      unsigned long long v = *reinterpret_cast<unsigned long long*>(&f2[i]);
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(*reinterpret_cast<const unsigned long long*>(&k2)),
                   "l"(*reinterpret_cast<const unsigned long long*>(&b2)));
      *reinterpret_cast<unsigned long long*>(&f2[i]) = v;
    }
#pragma unroll
    for (int i = 0; i < NF1; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f1[i]) : "f"(0.999f), "f"(1e-3f));
#pragma unroll
    for (int i = 0; i < NALU; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed * (float)it));
#pragma unroll
    for (int i = 0; i < NCVT; ++i) {
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(c[2 * i]), "f"(c[2 * i + 1]));
      c[2 * i] = __uint_as_float(pk[i] & 0xffff0000u);  // keep a dependency so the conversions are not hoisted
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < NM; ++i) s += m[i];
  for (int i = 0; i < NF2; ++i) s += f2[i].x + f2[i].y;
  for (int i = 0; i < NF1; ++i) s += f1[i];
  for (int i = 0; i < NALU; ++i) s += a[i];
  for (int i = 0; i < NCVT; ++i) s += __uint_as_float(pk[i]) + c[2 * i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NM, int NF2, int NF1, int NALU, int NCVT>
static void run(const char* name, int warps_per_smsp) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int threads = 128 * warps_per_smsp;
  float* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * threads);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  auto k = mix_kernel<NM, NF2, NF1, NALU, NCVT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<sms, threads, 200 * 1024>>>(out, cyc, 0.5f);  // 200 KB of dynamic shared memory: one CTA per SM
  k<<<sms, threads, 200 * 1024>>>(out, cyc, 0.5f);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += (double)h[i];
  avg /= sms;
  const double per_iter = avg / ITERS;
  printf("{\"mix\": \"%s\", \"warps_per_smsp\": %d, \"mufu\": %d, \"ffma2\": %d, \"ffma\": %d, \"fmnmx\": %d, \"f2fp\": %d, "
         "\"cycles_per_iter\": %.2f, \"cycles_per_iter_per_warp_instr\": %.3f}\n",
         name, warps_per_smsp, NM, NF2, NF1, NALU, NCVT, per_iter,
         per_iter / (warps_per_smsp * (NM + NF2 + NF1 + NALU + NCVT)));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {1, 2, 4}) {
    run<8, 0, 0, 0, 0>("mufu8", w);
    run<0, 16, 0, 0, 0>("ffma2x16", w);
    run<0, 0, 16, 0, 0>("ffma16", w);
    run<0, 0, 0, 16, 0>("fmnmx16", w);
    run<0, 0, 0, 0, 8>("f2fp8", w);
    run<8, 16, 0, 0, 0>("mufu8+ffma2x16", w);
    run<8, 0, 16, 0, 0>("mufu8+ffma16", w);
    run<8, 0, 0, 16, 0>("mufu8+fmnmx16", w);
    run<8, 0, 0, 0, 8>("mufu8+f2fp8", w);
    run<0, 16, 0, 16, 0>("ffma2x16+fmnmx16", w);
    run<8, 8, 0, 0, 4>("softmax_all_mufu(8 ex2, 4 ffma2, 4 fadd2, 4 f2fp)", w);
    run<6, 14, 4, 3, 4>("softmax_poly1(6 ex2, 14 packed, 2 imad~ffma, 3 alu, 4 f2fp)", w);
    run<8, 8, 0, 8, 4>("all_mufu+8 alu", w);
    run<8, 16, 0, 8, 4>("all_mufu+8 packed+8 alu", w);
  }
  return 0;
}
