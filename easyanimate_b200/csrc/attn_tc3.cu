// Joint text+video attention forward, third generation (the default): two 128-row query tiles per CTA, 64-key
// blocks, and the score accumulator of EVERY tile double-buffered in TMEM.
//
// ncu on the second generation (attn_tc2.cu, one S buffer per tile) showed the softmax warps spending most of their
// cycles on `mbarrier.try_wait(s_full)`: each tile's loop was a strict chain softmax -> wake MMA -> PV -> QK -> wake
// softmax, leaving the exp unit (the real limiter at head_dim 64: 16 ex2/clk/SM) ~48 % busy.  Here QK_{j+1} of a tile
// is issued BEFORE the MMA warp waits for P_j, into the tile's other S buffer, so the next block's scores are already
// in TMEM when the softmax warps finish the current one.
//
//   TMEM (512 columns): S[t][b] at (2t+b)*64 (P_j aliases the first 32 columns of S[t][j&1]); O_t at 256 + 64 t.
//   warps 0-3 / 4-7 : softmax of tile A / B (one query row per thread, two 32-column chunks per pass)
//   warp 8          : TMA producer (Q once, K_j / V_j 64-key tiles through 6-stage rings)
//   warp 9          : MMA issuer;   warp 10: TMEM allocator
#include "common.cuh"
#include "host.h"
#include "../../include/ea_b200.h"

namespace ea {

extern void count_launch();

namespace a3 {

constexpr int kThreads = 384;
constexpr int kQT = 128;
constexpr int kKT = 64;
constexpr int kHD = 64;
constexpr int kStages = 6;

struct Args {
  bf16* out_text;
  bf16* out_video;
  int B, H, S, S_text;
  float scale_log2;
};

struct Smem {
  static constexpr int kQBytes = 2 * kQT * kHD * 2;  // 32 KB
  static constexpr int kKBytes = kKT * kHD * 2;      // 8 KB
  static constexpr int kVBytes = kKT * kHD * 2;
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = kOffQ + kQBytes;
  static constexpr int kOffV = kOffK + kStages * kKBytes;
  static constexpr int kOffBar = kOffV + kStages * kVBytes;
  static constexpr int kTotal = kOffBar + 512 + 1024;
};

EA_DEVICE void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
EA_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x (x <= 0) on the FMA/ALU pipes: x = n + f (round to nearest), degree-3 minimax 2^f, exponent add.
EA_DEVICE float exp2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float y = x + 12582912.0f;
  const float n = y - 12582912.0f;
  const float f = x - n;
  float p = 0.05500892f;
  p = fmaf(p, f, 0.24221097f);
  p = fmaf(p, f, 0.69328290f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(y) << 23));
}

// one 64-key block of one row: returns the (scaled) block max in pass 1, writes P and returns the row sum in pass 2
template <bool TAIL>
EA_DEVICE float row_max(uint32_t tS, int valid) {
  uint32_t va[32], vb[32];
  tmem_ld32(tS, va);
  tmem_ld32(tS + 32, vb);
  tmem_ld_wait();
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (!TAIL || i < valid) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(va[i]));
    if (!TAIL || 32 + i < valid) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(vb[i]));
  }
  return fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
}

template <int POLY, bool TAIL>
EA_DEVICE float row_exp_store(uint32_t tS, int valid, float scale_log2, float neg_m) {
  uint32_t va[32], vb[32];
  tmem_ld32(tS, va);
  tmem_ld32(tS + 32, vb);
  tmem_ld_wait();
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t(&cur)[32] = c ? vb : va;
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float x0 = fmaf(__uint_as_float(cur[i]), scale_log2, neg_m);
      const float x1 = fmaf(__uint_as_float(cur[i + 1]), scale_log2, neg_m);
      float e0 = ((i & 7) < POLY) ? exp2_poly(x0) : ex2(x0);
      float e1 = (((i + 1) & 7) < POLY) ? exp2_poly(x1) : ex2(x1);
      if (TAIL) {
        if (c * 32 + i >= valid) e0 = 0.f;
        if (c * 32 + i + 1 >= valid) e1 = 0.f;
      }
      s4[(i >> 1) & 3] += e0 + e1;
      pk[i >> 1] = pack_bf16x2(e0, e1);
    }
    tmem_st16(tS + c * 16, pk);  // both S chunks are already in registers: P may overwrite columns [0,32)
  }
  return (s4[0] + s4[1]) + (s4[2] + s4[3]);
}

template <int POLY>
__global__ void __launch_bounds__(kThreads, 1)
attn3_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
             const __grid_constant__ CUtensorMap tmap_v, const Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + Smem::kOffQ;
  uint8_t* sK = smem + Smem::kOffK;
  uint8_t* sV = smem + Smem::kOffV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::kOffBar);
  uint64_t* q_full = bars;                  // 1
  uint64_t* k_full = bars + 1;              // kStages
  uint64_t* k_empty = k_full + kStages;
  uint64_t* v_full = k_empty + kStages;
  uint64_t* v_empty = v_full + kStages;
  uint64_t* s_full = v_empty + kStages;     // [tile*2 + buf]
  uint64_t* p_ready = s_full + 4;           // [tile*2 + buf]
  uint64_t* o_done = p_ready + 4;           // [tile]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * kQT);
  const int bh = blockIdx.y;
  const int nblk = (p.S + kKT - 1) / kKT;

  constexpr uint32_t kColO = 256;
  constexpr uint32_t kTmemCols = 512;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
    }
    mbar_init(&o_done[0], 1);
    mbar_init(&o_done[1], 1);
    fence_mbar_init();
  }
  if (warp == 10) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      // ===== TMA producer =====
      mbar_arrive_expect_tx(q_full, Smem::kQBytes);
      tma_load_3d(sQ, &tmap_q, q_full, 0, q0, bh);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], Smem::kKBytes);
        tma_load_3d(sK + st * Smem::kKBytes, &tmap_k, &k_full[st], 0, j * kKT, bh);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], Smem::kVBytes);
        tma_load_3d(sV + st * Smem::kVBytes, &tmap_v, &v_full[st], 0, j * kKT, bh);
        if (++st == kStages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc_qk = umma_idesc_bf16(kQT, kKT, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kQT, kHD, 0, 1);  // V: MN-major B operand
      mbar_wait(q_full, 0);
      tc_fence_after();
      auto issue_qk = [&](int t, int j) {
        const int st = j % kStages;
        const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ + t * (kQT * kHD * 2)));
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + st * Smem::kKBytes));
        const uint32_t d = tmem_base + (2 * t + (j & 1)) * kKT;
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[2 * t + (j & 1)]);
      };
      auto issue_pv = [&](int t, int j) {
        const int st = j % kStages;
        const uint32_t vaddr = smem_u32(sV + st * Smem::kVBytes);
        const uint32_t d = tmem_base + kColO + t * kHD;
        const uint32_t pa = tmem_base + (2 * t + (j & 1)) * kKT;  // packed bf16 P over the first 32 columns
#pragma unroll
        for (int k = 0; k < kKT / 16; ++k)
          umma_ts(d, pa + k * 8, umma_desc_sw128_mn(vaddr + k * 2048, 8192, 1024), idesc_pv, (j | k) != 0);
        umma_commit(&o_done[t]);
      };
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      issue_qk(1, 0);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % kStages;
        const uint32_t par = (j >> 1) & 1;
        const bool more = j + 1 < nblk;
        if (more) {
          mbar_wait(&k_full[(j + 1) % kStages], ((j + 1) / kStages) & 1);
          tc_fence_after();
          issue_qk(0, j + 1);  // next block's scores of tile A, computed while its softmax works on block j
        }
        mbar_wait(&p_ready[0 + (j & 1)], par);
        mbar_wait(&v_full[st], (j / kStages) & 1);
        tc_fence_after();
        issue_pv(0, j);
        if (more) {
          issue_qk(1, j + 1);
          umma_commit(&k_empty[(j + 1) % kStages]);
        }
        mbar_wait(&p_ready[2 + (j & 1)], par);
        tc_fence_after();
        issue_pv(1, j);
        umma_commit(&v_empty[st]);
      }
    }
  } else if (warp < 8) {
    // ===== softmax / correction / epilogue: tile t, one query row per thread =====
    const int t = warp >> 2;
    const int ew = warp & 3;
    const int r = ew * 32 + lane;
    const uint32_t lane_off = uint32_t(ew * 32) << 16;
    const uint32_t tO = tmem_base + lane_off + kColO + t * kHD;
    float m_ref = -INFINITY;
    float l = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int b = j & 1;
      const uint32_t tS = tmem_base + lane_off + (2 * t + b) * kKT;
      mbar_wait(&s_full[2 * t + b], (j >> 1) & 1);
      tc_fence_after();
      const int valid = p.S - j * kKT;
      const bool tail = valid < kKT;  // only the last block can be ragged (TMA zero-filled the missing keys)
      const float mx = (tail ? row_max<true>(tS, valid) : row_max<false>(tS, valid)) * p.scale_log2;
      if (j == 0) {
        m_ref = mx;
      } else {
        const bool grow = mx > m_ref + 8.0f;
        if (__any_sync(0xffffffffu, grow)) {
          mbar_wait(&o_done[t], (j - 1) & 1);
          tc_fence_after();
          const float m_new = grow ? mx : m_ref;
          const float f = ex2(m_ref - m_new);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(tO + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st32(tO + c * 32, v);
          }
          tmem_st_wait();
          l *= f;
          m_ref = m_new;
        }
      }
      l += tail ? row_exp_store<POLY, true>(tS, valid, p.scale_log2, -m_ref)
                : row_exp_store<POLY, false>(tS, valid, p.scale_log2, -m_ref);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[2 * t + b]);
    }
    mbar_wait(&o_done[t], (nblk - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int srow = q0 + t * kQT + r;
    bf16* dst = nullptr;
    if (srow < p.S) {
      const int bb = bh / p.H, h = bh % p.H;
      const int64_t d = (int64_t)p.H * kHD;
      if (srow < p.S_text)
        dst = p.out_text + ((int64_t)bb * p.S_text + srow) * d + h * kHD;
      else
        dst = p.out_video + ((int64_t)bb * (p.S - p.S_text) + (srow - p.S_text)) * d + h * kHD;
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      __syncwarp();
      tmem_ld32(tO + c * 32, v);
      tmem_ld_wait();
      if (dst != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(v[i * 8 + 0]) * inv_l, __uint_as_float(v[i * 8 + 1]) * inv_l);
          w.y = pack_bf16x2(__uint_as_float(v[i * 8 + 2]) * inv_l, __uint_as_float(v[i * 8 + 3]) * inv_l);
          w.z = pack_bf16x2(__uint_as_float(v[i * 8 + 4]) * inv_l, __uint_as_float(v[i * 8 + 5]) * inv_l);
          w.w = pack_bf16x2(__uint_as_float(v[i * 8 + 6]) * inv_l, __uint_as_float(v[i * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c * 32 + i * 8) = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int POLY>
static int launch(const ea_attn_args* g, cudaStream_t stream) {
  const int64_t BH = g->B * g->H;
  CUtensorMap tq, tk, tv;
  uint64_t dims[3] = {(uint64_t)kHD, (uint64_t)g->S, (uint64_t)BH};
  uint64_t strides[2] = {(uint64_t)kHD * 2, (uint64_t)g->S * kHD * 2};
  uint32_t box_q[3] = {kHD, 2 * kQT, 1};
  uint32_t box_kv[3] = {kHD, kKT, 1};
  int rc = make_tmap_bf16(&tq, g->q, 3, dims, strides, box_q, true);
  if (rc) return rc;
  rc = make_tmap_bf16(&tk, g->k, 3, dims, strides, box_kv, true);
  if (rc) return rc;
  rc = make_tmap_bf16(&tv, g->v, 3, dims, strides, box_kv, true);
  if (rc) return rc;
  Args p{};
  p.out_text = reinterpret_cast<bf16*>(g->out_text);
  p.out_video = reinterpret_cast<bf16*>(g->out_video);
  p.B = (int)g->B; p.H = (int)g->H; p.S = (int)g->S; p.S_text = (int)g->S_text;
  p.scale_log2 = g->scale * 1.4426950408889634f;
  auto kern = attn3_kernel<POLY>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::kTotal);
    if (e != cudaSuccess) return fail(EA_ERR_CUDA, std::string("cudaFuncSetAttribute(attn3): ") + cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)((g->S + 2 * kQT - 1) / (2 * kQT)), (unsigned)BH);
  kern<<<grid, kThreads, Smem::kTotal, stream>>>(tq, tk, tv, p);
  count_launch();
  return check_launch("attn3_kernel");
}

}  // namespace a3

int launch_attn3(const ea_attn_args* g, int poly, cudaStream_t stream) {
  switch (poly) {
    case 0: return a3::launch<0>(g, stream);
    case 2: return a3::launch<2>(g, stream);
    case 3: return a3::launch<3>(g, stream);
    case 4: return a3::launch<4>(g, stream);
    default: return fail(EA_ERR_INVALID, "ea_attn_fwd: unsupported polynomial fraction (0,2,3,4 of every 8)");
  }
}

}  // namespace ea
