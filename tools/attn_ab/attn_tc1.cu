// Joint text+video self-attention forward on tcgen05 (head_dim 64, non-causal, no mask):
//     O = softmax(Q K^T * scale) V         q,k,v: [B,H,S,64] bf16
// replacing F.scaled_dot_product_attention at easyanimate/models/processor.py:287-289 and the
// transpose/reshape/split that follows it (:291-303): the output is written token-major, split into the text rows
// [B,S_text,H*64] and the video rows [B,S-S_text,H*64] that the two out-projections consume.
//
// One CTA = one 128-row query tile of one (batch, head).  Roles:
//   warp 0 (1 thread) : TMA producer   - Q once, then K_j / V_j 128-key tiles through two 3-stage rings
//   warp 1 (1 thread) : MMA issuer     - S_j = Q K_j^T into a double-buffered TMEM accumulator (so QK_{j+1} overlaps
//                                        softmax_j), then O += P_j V_j
//   warp 2            : TMEM allocator
//   warps 4-7         : softmax        - one query row per thread: tcgen05.ld S_j, online softmax in fp32 with a lazy
//                                        running-max update (O rescaled in TMEM only when the max grows by > 2^8),
//                                        P_j written as bf16 either to SWIZZLE_128B shared memory or to TMEM;
//                                        final O / l, bf16 store.
// Template switches exist because two operand paths were brought up side by side on hardware:
//   P_TMEM : P_j is the A operand from TMEM (tcgen05.mma [d],[a],b) instead of from shared memory
//   V_TRANS: V is supplied pre-transposed ([B,H,64,S_pad], K-major B operand) instead of as an MN-major operand
#include "../../easyanimate_b200/csrc/common.cuh"
#include "../../easyanimate_b200/csrc/host.h"
#include "../../include/ea_b200.h"

namespace ea {

extern void count_launch();

constexpr int kAttnThreads = 256;
constexpr int kQT = 128;   // query rows per CTA
constexpr int kKT = 128;   // keys per block
constexpr int kHD = 64;
constexpr int kKVStages = 3;

struct AttnDevArgs {
  bf16* out_text;
  bf16* out_video;
  int B, H, S, S_text;
  float scale_log2;  // softmax scale * log2(e)
};

struct AttnSmem {
  static constexpr int kQBytes = kQT * kHD * 2;                 // 16 KB
  static constexpr int kKBytes = kKT * kHD * 2;                 // 16 KB
  static constexpr int kVBytes = kKT * kHD * 2;                 // 16 KB
  static constexpr int kPBytes = kQT * kKT * 2;                 // 32 KB (two 64-key halves of 16 KB)
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = kOffQ + kQBytes;
  static constexpr int kOffV = kOffK + kKVStages * kKBytes;
  static constexpr int kOffP = kOffV + kKVStages * kVBytes;
  static constexpr int kOffBar = kOffP + 2 * kPBytes;
  static constexpr int kTotal = kOffBar + 256 + 1024;
};

template <bool P_TMEM, bool V_TRANS>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnDevArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + AttnSmem::kOffQ;
  uint8_t* sK = smem + AttnSmem::kOffK;
  uint8_t* sV = smem + AttnSmem::kOffV;
  uint8_t* sP = smem + AttnSmem::kOffP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnSmem::kOffBar);
  uint64_t* q_full = bars;                       // 1
  uint64_t* k_full = bars + 1;                   // kKVStages
  uint64_t* k_empty = k_full + kKVStages;
  uint64_t* v_full = k_empty + kKVStages;
  uint64_t* v_empty = v_full + kKVStages;
  uint64_t* s_full = v_empty + kKVStages;        // 2
  uint64_t* p_ready = s_full + 2;                // 2
  uint64_t* o_done = p_ready + 2;                // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQT;
  const int bh = blockIdx.y;
  const int nblk = (p.S + kKT - 1) / kKT;

  // TMEM columns: S0 [0,128) S1 [128,256) O [256,320) P0 [320,384) P1 [384,448)
  constexpr uint32_t kColS = 0, kColO = 256, kColP = 320;
  constexpr uint32_t kTmemCols = 512;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKVStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
    }
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      mbar_arrive_expect_tx(q_full, AttnSmem::kQBytes);
      tma_load_3d(sQ, &tmap_q, q_full, 0, q0, bh);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], AttnSmem::kKBytes);
        tma_load_3d(sK + st * AttnSmem::kKBytes, &tmap_k, &k_full[st], 0, j * kKT, bh);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], AttnSmem::kVBytes);
        if constexpr (V_TRANS) {
          // V^T [64 hd rows][S_pad keys]: two 64-key chunks, each a K-major [64 x 64] tile
          tma_load_3d(sV + st * AttnSmem::kVBytes, &tmap_v, &v_full[st], j * kKT, 0, bh);
          tma_load_3d(sV + st * AttnSmem::kVBytes + 8192, &tmap_v, &v_full[st], j * kKT + 64, 0, bh);
        } else {
          tma_load_3d(sV + st * AttnSmem::kVBytes, &tmap_v, &v_full[st], 0, j * kKT, bh);
        }
        if (++st == kKVStages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc_qk = umma_idesc_bf16(kQT, kKT, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kQT, kHD, 0, V_TRANS ? 0 : 1);
      mbar_wait(q_full, 0);
      tc_fence_after();
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      auto issue_qk = [&](int j) {
        const int st = j % kKVStages;
        mbar_wait(&k_full[st], (j / kKVStages) & 1);
        tc_fence_after();
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + st * AttnSmem::kKBytes));
        const uint32_t d = tmem_base + kColS + (j & 1) * kKT;
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[j & 1]);
      };
      issue_qk(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_qk(j + 1);
        mbar_wait(&p_ready[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const int st = j % kKVStages;
        mbar_wait(&v_full[st], (j / kKVStages) & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + kColO;
        const uint32_t vaddr = smem_u32(sV + st * AttnSmem::kVBytes);
#pragma unroll
        for (int k = 0; k < kKT / 16; ++k) {
          uint64_t bdesc;
          if constexpr (V_TRANS) {
            bdesc = umma_desc_sw128(vaddr + (k >> 2) * 8192) + 2 * (k & 3);
          } else {
            // MN-major [128 keys][64 hd]: 16 keys per MMA = 2048 B
            bdesc = umma_desc_sw128_mn(vaddr + k * 2048, 16384, 1024);
          }
          if constexpr (P_TMEM) {
            umma_ts(d, tmem_base + kColP + (j & 1) * 64 + k * 8, bdesc, idesc_pv, (j | k) != 0);
          } else {
            const uint64_t adesc = umma_desc_sw128(smem_u32(sP + (j & 1) * AttnSmem::kPBytes + (k >> 2) * 16384)) + 2 * (k & 3);
            umma_ss(d, adesc, bdesc, idesc_pv, (j | k) != 0);
          }
        }
        umma_commit(&v_empty[st]);
        umma_commit(o_done);
      }
    }
  } else if (warp >= 4) {
    // ===== softmax / correction / epilogue: one query row per thread =====
    const int ew = warp - 4;
    const int r = ew * 32 + lane;  // row in tile == TMEM lane
    const uint32_t lane_off = uint32_t(ew * 32) << 16;
    float m_ref = -INFINITY;  // running reference max, in scaled log2 units
    float l = 0.f;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      float s[kKT];
      {
        const uint32_t ta = tmem_base + lane_off + kColS + (j & 1) * kKT;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t t[32];
          tmem_ld32(ta + c * 32, t);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(t[i]) * p.scale_log2;
        }
      }
      if (j == nblk - 1) {
        const int valid = p.S - j * kKT;  // keys beyond S were zero-filled by TMA: mask them out
#pragma unroll
        for (int i = 0; i < kKT; ++i)
          if (i >= valid) s[i] = -INFINITY;
      }
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < kKT; ++i) mx = fmaxf(mx, s[i]);
      if (j == 0) {
        m_ref = mx;
      } else {
        const bool grow = mx > m_ref + 8.0f;
        if (__any_sync(0xffffffffu, grow)) {
          // rescale the running output in TMEM; needs PV_{j-1} complete
          mbar_wait(o_done, (j - 1) & 1);
          tc_fence_after();
          const float m_new = grow ? mx : m_ref;
          const float f = exp2f(m_ref - m_new);
          const uint32_t to = tmem_base + lane_off + kColO;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t t[32];
            tmem_ld32(to + c * 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * f);
            tmem_st32(to + c * 32, t);
          }
          tmem_st_wait();
          l *= f;
          m_ref = m_new;
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < kKT; ++i) {
        s[i] = exp2f(s[i] - m_ref);
        sum += s[i];
      }
      l += sum;
      if constexpr (P_TMEM) {
        const uint32_t tp = tmem_base + lane_off + kColP + (j & 1) * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t t[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) t[i] = pack_bf16x2(s[c * 64 + 2 * i], s[c * 64 + 2 * i + 1]);
          tmem_st32(tp + c * 32, t);
        }
        tmem_st_wait();
        tc_fence_before();
      } else {
        // K-major SWIZZLE_128B: row r, 16-byte chunk c of each 64-key half lands at chunk (c ^ (r & 7))
        uint8_t* prow = sP + (j & 1) * AttnSmem::kPBytes + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 w;
            const int b0 = half * 64 + c * 8;
            w.x = pack_bf16x2(s[b0 + 0], s[b0 + 1]);
            w.y = pack_bf16x2(s[b0 + 2], s[b0 + 3]);
            w.z = pack_bf16x2(s[b0 + 4], s[b0 + 5]);
            w.w = pack_bf16x2(s[b0 + 6], s[b0 + 7]);
            *reinterpret_cast<uint4*>(prow + half * 16384 + ((c ^ (r & 7)) << 4)) = w;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
      }
      mbar_arrive(&p_ready[j & 1]);
    }
    // final: wait for the last PV, normalise, store
    mbar_wait(o_done, (nblk - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int srow = q0 + r;
    bf16* dst = nullptr;
    if (srow < p.S) {
      const int b = bh / p.H, h = bh % p.H;
      const int64_t d = (int64_t)p.H * kHD;
      if (srow < p.S_text)
        dst = p.out_text + ((int64_t)b * p.S_text + srow) * d + h * kHD;
      else
        dst = p.out_video + ((int64_t)b * (p.S - p.S_text) + (srow - p.S_text)) * d + h * kHD;
    }
    const uint32_t to = tmem_base + lane_off + kColO;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t t[32];
      tmem_ld32(to + c * 32, t);
      tmem_ld_wait();
      if (dst != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(t[i * 8 + 0]) * inv_l, __uint_as_float(t[i * 8 + 1]) * inv_l);
          w.y = pack_bf16x2(__uint_as_float(t[i * 8 + 2]) * inv_l, __uint_as_float(t[i * 8 + 3]) * inv_l);
          w.z = pack_bf16x2(__uint_as_float(t[i * 8 + 4]) * inv_l, __uint_as_float(t[i * 8 + 5]) * inv_l);
          w.w = pack_bf16x2(__uint_as_float(t[i * 8 + 6]) * inv_l, __uint_as_float(t[i * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c * 32 + i * 8) = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <bool P_TMEM, bool V_TRANS>
static int launch_attn(const ea_attn_args* g, cudaStream_t stream) {
  const int64_t BH = g->B * g->H;
  CUtensorMap tq, tk, tv;
  {
    uint64_t dims[3] = {(uint64_t)kHD, (uint64_t)g->S, (uint64_t)BH};
    uint64_t strides[2] = {(uint64_t)kHD * 2, (uint64_t)g->S * kHD * 2};
    uint32_t box[3] = {kHD, kQT, 1};
    int rc = make_tmap_bf16(&tq, g->q, 3, dims, strides, box, true);
    if (rc) return rc;
    rc = make_tmap_bf16(&tk, g->k, 3, dims, strides, box, true);
    if (rc) return rc;
    if (!V_TRANS) {
      rc = make_tmap_bf16(&tv, g->v, 3, dims, strides, box, true);
      if (rc) return rc;
    }
  }
  if (V_TRANS) {
    uint64_t dims[3] = {(uint64_t)g->S_pad, (uint64_t)kHD, (uint64_t)BH};
    uint64_t strides[2] = {(uint64_t)g->S_pad * 2, (uint64_t)g->S_pad * kHD * 2};
    uint32_t box[3] = {64, kHD, 1};
    int rc = make_tmap_bf16(&tv, g->v, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  AttnDevArgs p{};
  p.out_text = reinterpret_cast<bf16*>(g->out_text);
  p.out_video = reinterpret_cast<bf16*>(g->out_video);
  p.B = (int)g->B; p.H = (int)g->H; p.S = (int)g->S; p.S_text = (int)g->S_text;
  p.scale_log2 = g->scale * 1.4426950408889634f;
  auto kern = attn_fwd_kernel<P_TMEM, V_TRANS>;
  static ::ea::PerDeviceFlag attr_flag;
  const int attr_dev = ::ea::current_device();
  if (!attr_flag.get(attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal);
    if (e != cudaSuccess) return fail(EA_ERR_CUDA, std::string("cudaFuncSetAttribute(attn): ") + cudaGetErrorString(e));
    attr_flag.set(attr_dev);
  }
  dim3 grid((unsigned)((g->S + kQT - 1) / kQT), (unsigned)BH);
  kern<<<grid, kAttnThreads, AttnSmem::kTotal, stream>>>(tq, tk, tv, p);
  count_launch();
  return check_launch("attn_fwd_kernel");
}

}  // namespace ea

namespace ea {
// first-generation kernel behind ea_attn_fwd's variants 0-3 (A/B builds only, see easyanimate_b200/csrc/attn_api.cu)
int launch_attn1(const ea_attn_args* g, cudaStream_t stream) {
  EA_REQUIRE((g->variant & ~3) == 0, "ea_attn_fwd: unknown kernel variant");
  const bool vt = (g->variant & 2) != 0, pt = (g->variant & 1) != 0;
  if (vt) EA_REQUIRE(g->S_pad >= g->S && g->S_pad % 8 == 0, "ea_attn_fwd: S_pad must be >= S and a multiple of 8");
  if (!pt && !vt) return launch_attn<false, false>(g, stream);
  if (pt && !vt) return launch_attn<true, false>(g, stream);
  if (!pt && vt) return launch_attn<false, true>(g, stream);
  return launch_attn<true, true>(g, stream);
}
}  // namespace ea

namespace ea {
__global__ void transpose_v_kernel(const bf16* __restrict__ v, bf16* __restrict__ vt, int S, int S_pad) {
  __shared__ bf16 tile[64][66];
  const int bh = blockIdx.y;
  const int s0 = blockIdx.x * 64;
  const bf16* src = v + ((int64_t)bh * S + s0) * 64;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int r = i >> 6, c = i & 63;
    tile[r][c] = (s0 + r < S) ? src[(int64_t)r * 64 + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  bf16* dst = vt + (int64_t)bh * 64 * S_pad + s0;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int c = i >> 6, r = i & 63;  // c: head-dim row of vt, r: key
    if (s0 + r < S_pad) dst[(int64_t)c * S_pad + r] = tile[r][c];
  }
}
}  // namespace ea

extern "C" int ea_transpose_v(const void* v, void* vt, int64_t BH, int64_t S, int64_t S_pad, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  EA_REQUIRE(v && vt && BH > 0 && S > 0 && S_pad >= S && S_pad % 8 == 0 && BH <= 65535, "ea_transpose_v: bad arguments");
  dim3 grid((unsigned)((S_pad + 63) / 64), (unsigned)BH);
  ea::transpose_v_kernel<<<grid, 256, 0, stream>>>((const ea::bf16*)v, (ea::bf16*)vt, (int)S, (int)S_pad);
  ea::count_launch();
  return ea::check_launch("transpose_v_kernel");
}
