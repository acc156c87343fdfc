// The softmax inner loop of the attention kernels in isolation (no TMEM, no MMA, no barriers): 112 scores per thread in
// registers, POLY8 of every 8 column pairs through the FMA-pipe polynomial and the rest through MUFU.EX2, packed by
// truncation (PRMT) or F2FP, with or without the FADD2 row sum.  W warps per sub-partition, one CTA per SM, cycles per
// row block from clock64.  Answers: how many cycles does a softmax warp need per key block if nothing else is in the way?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/softmax_mix_bench tools/microbench/softmax_mix_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 512;
constexpr int kKT = 112;
constexpr int kPairs = kKT / 2;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool CLAMP>
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  if (CLAMP) {
    x.x = fmaxf(x.x, -126.0f);
    x.y = fmaxf(x.y, -126.0f);
  }
  const float2 y = __fadd2_rn(x, make_float2(12582912.0f, 12582912.0f));
  const float2 n = __fadd2_rn(y, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __ffma2_rn(n, make_float2(-1.0f, -1.0f), x);
  float2 q = __ffma2_rn(f, make_float2(0.05500892f, 0.05500892f), make_float2(0.24221097f, 0.24221097f));
  q = __ffma2_rn(q, f, make_float2(0.69328290f, 0.69328290f));
  q = __ffma2_rn(q, f, make_float2(1.0f, 1.0f));
  float2 e;
  e.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(y.x) << 23));
  e.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(y.y) << 23));
  return e;
}

// MODE bit0: F2FP (round to nearest) instead of PRMT truncation; bit1: FADD2 row sum; bit2: no clamp / no guard in the
// polynomial path (lower bound of its cost)
template <int POLY8, int MODE>
__global__ void softmax_kernel(const float* in, uint32_t* out, long long* cycles, float scale) {
  uint32_t s[kKT];
  for (int i = 0; i < kKT; ++i) s[i] = __float_as_uint(in[(threadIdx.x * kKT + i) % 4096]);
  float m_ref = 3.0f;
  uint32_t keep = 0;
  float2 acc = make_float2(0.f, 0.f);
  float guard = -1e30f;
  const float2 c2 = make_float2(scale, scale);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    const float2 nm2 = make_float2(-m_ref, -m_ref);
    uint32_t orw = 0;
#pragma unroll
    for (int q = 0; q < kPairs; ++q) {
      const float s0 = __uint_as_float(s[2 * q]), s1 = __uint_as_float(s[2 * q + 1]);
      const float2 x = __ffma2_rn(make_float2(s0, s1), c2, nm2);
      float2 e;
      if ((q & 7) < POLY8) {
        if (!(MODE & 4)) guard = fmaxf(guard, fmaxf(s0, s1));
        e = exp2_poly2<!(MODE & 4)>(x);
      } else {
        e.x = ex2(x.x);
        e.y = ex2(x.y);
      }
      if (MODE & 2) acc = __fadd2_rn(acc, e);
      uint32_t pk;
      if (MODE & 1) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(e.y), "f"(e.x));
      else pk = __byte_perm(__float_as_uint(e.x), __float_as_uint(e.y), 0x7632);
      orw |= pk;
    }
    keep ^= orw;
    m_ref += 1e-3f + __uint_as_float(orw & 1u);  // every iteration depends on the previous one's result
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = keep ^ __float_as_uint(acc.x + acc.y + guard);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int POLY8, int MODE>
static void run(int warps_per_smsp, const float* in) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int threads = 128 * warps_per_smsp;
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(uint32_t) * sms * threads);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  auto k = softmax_kernel<POLY8, MODE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<sms, threads, 200 * 1024>>>(in, out, cyc, 0.18f);
  k<<<sms, threads, 200 * 1024>>>(in, out, cyc, 0.18f);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += (double)h[i];
  avg /= sms;
  const double per_block = avg / ITERS;
  printf("{\"poly8\": %d, \"pack\": \"%s\", \"row_sum_fadd2\": %d, \"poly_clamp_guard\": %d, \"warps_per_smsp\": %d, "
         "\"cycles_per_112key_block\": %.1f, \"xu_cycles_needed\": %d, \"err\": \"%s\"}\n",
         POLY8, (MODE & 1) ? "f2fp" : "prmt", (MODE >> 1) & 1, !(MODE & 4), warps_per_smsp, per_block,
         warps_per_smsp * (kPairs - 7 * POLY8) * 2 * 8, cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  float* in;
  cudaMalloc(&in, 4096 * sizeof(float));
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = (float)((i * 2654435761u) % 1000) / 100.0f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int w : {2, 4}) {
    run<0, 0>(w, in);
    run<1, 0>(w, in);
    run<2, 0>(w, in);
    run<3, 0>(w, in);
    run<4, 0>(w, in);
    run<2, 4>(w, in);
    run<4, 4>(w, in);
    run<0, 3>(w, in);  // the sixth generation's mix: F2FP + FADD2
    run<2, 3>(w, in);
    run<0, 1>(w, in);
    run<0, 2>(w, in);
  }
  return 0;
}
