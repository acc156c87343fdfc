// Joint text+video attention forward, fourth generation (the default): two 128-row query tiles per CTA, ONE pass over
// the score tile.
//
// What the measurements on B200 said (profiles/r01_attn_*, r01_ncu_attn_*, r01_mufu_microbench, r01_tmem_microbench):
//  * head_dim 64 is exp-bound, not tensor-bound: a 128x128 score tile is 512 tensor cycles but 16 384 exponentials at
//    16 MUFU.EX2/clk/SM = 1 024 cycles; ex2 on packed bf16x2/f16x2 compiles to two MUFU ops, so it gains nothing.
//  * the softmax is also ISSUE-bound: the earlier generations executed 9-13 warp instructions per score element
//    (scalar FFMA/FADD, exp2f range fix-ups, mbarrier polling) on the 2 softmax warps each SM sub-partition has.
//  * one MMA issuer walking both tiles in a fixed order locks the tiles in phase (head-of-line blocking on p_ready).
// So here: each softmax thread reads its 128 scores ONCE into registers (setmaxnreg: 216 registers), one tcgen05.ld
// in flight with the row max computed under the next load; the S buffer is released right after the read so QK_{j+1}
// runs under this block's exponentials; scale-subtract, row sum and the exp2 polynomial use packed fp32x2
// instructions (FFMA2/FADD2); a fraction of the exponentials is moved from MUFU to the FMA pipe; each tile has its own
// MMA issuer warp and the tiles start half a period apart.
//
//   TMEM (512 columns): S_t at 128 t | P_t (packed bf16) at 256 + 64 t | O_t at 384 + 64 t
//   warps 0-3 / 4-7 : softmax of tile A / B, one query row per thread
//   warp 8          : TMA producer (Q once, K_j / V_j 128-key tiles through 4-stage rings)
//   warp 9          : TMEM allocator, then MMA issuer;  warps 10-11: idle (complete the third warpgroup)
#include "../../easyanimate_b200/csrc/common.cuh"
#include "../../easyanimate_b200/csrc/host.h"
#include "../../include/ea_b200.h"

namespace ea {

extern void count_launch();

namespace a4 {

constexpr int kThreads = 384;
constexpr int kQT = 128;
constexpr int kKT = 128;
constexpr int kHD = 64;
constexpr int kStages = 4;

struct Args {
  bf16* out_text;
  bf16* out_video;
  int B, H, S, S_text;
  float scale_log2;
};

struct Smem {
  static constexpr int kQBytes = 2 * kQT * kHD * 2;
  static constexpr int kKBytes = kKT * kHD * 2;
  static constexpr int kVBytes = kKT * kHD * 2;
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = kOffQ + kQBytes;
  static constexpr int kOffV = kOffK + kStages * kKBytes;
  static constexpr int kOffBar = kOffV + kStages * kVBytes;
  static constexpr int kTotal = kOffBar + 512 + 1024;
};

EA_DEVICE void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
EA_DEVICE void tmem_ld32p(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
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

// Register budget: each SM sub-partition owns 16 384 registers and hosts 3 of the 12 warps: the two softmax warps
// grow to 216 registers with setmaxnreg (they hold a 128-column score row), the producer/MMA warpgroup shrinks to 72
// (2 x 216 + 72 = 3 x 168, the launch-time allocation).
// the same on a pair of columns with packed fp32x2 instructions (3 FADD2/FFMA2 for the split, 3 FFMA2 for the polynomial)
EA_DEVICE float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
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

template <int POLY>
__global__ void __launch_bounds__(kThreads, 1)
attn4_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
             const __grid_constant__ CUtensorMap tmap_v, const Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + Smem::kOffQ;
  uint8_t* sK = smem + Smem::kOffK;
  uint8_t* sV = smem + Smem::kOffV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::kOffBar);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + kStages;
  uint64_t* v_full = k_empty + kStages;
  uint64_t* v_empty = v_full + kStages;
  uint64_t* s_full = v_empty + kStages;   // [tile]
  uint64_t* s_free = s_full + 2;          // [tile]
  uint64_t* p_ready = s_free + 2;         // [tile]
  uint64_t* o_done = p_ready + 2;         // [tile]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * kQT);
  const int bh = blockIdx.y;
  const int nblk = (p.S + kKT - 1) / kKT;

  constexpr uint32_t kColP = 256, kColO = 384;
  constexpr uint32_t kTmemCols = 512;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);  // one tcgen05.commit per tile issuer
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
  // ---- producer / MMA warpgroup: give registers back
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
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
  } else if (warp <= 10) {
    if (lane == 0) {
      // ===== MMA issuers: warp 9 drives tile A, warp 10 drives tile B, independently =====
      // (one issuer walking both tiles in a fixed order head-of-line blocks: it sits in wait(p_ready[A]) while
      //  s_free[B] has long fired, the tiles lock IN phase and the TMEM-read and exp phases never overlap.)
      const int t = warp - 9;
      constexpr uint32_t idesc_qk = umma_idesc_bf16(kQT, kKT, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kQT, kHD, 0, 1);  // V: MN-major B operand
      mbar_wait(q_full, 0);
      tc_fence_after();
      auto issue_qk = [&](int t, int j) {
        const int st = j % kStages;
        const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ + t * (kQT * kHD * 2)));
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + st * Smem::kKBytes));
        const uint32_t d = tmem_base + t * kKT;
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int j) {
        const int st = j % kStages;
        const uint32_t vaddr = smem_u32(sV + st * Smem::kVBytes);
        const uint32_t d = tmem_base + kColO + t * kHD;
        const uint32_t pa = tmem_base + kColP + t * 64;
#pragma unroll
        for (int k = 0; k < kKT / 16; ++k)
          umma_ts(d, pa + k * 8, umma_desc_sw128_mn(vaddr + k * 2048, 16384, 1024), idesc_pv, (j | k) != 0);
        umma_commit(&o_done[t]);
      };
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(t, 0);
      umma_commit(&k_empty[0]);  // K/V stages are released when BOTH issuers have committed (barrier count 2)
      for (int j = 0; j < nblk; ++j) {
        const int st = j % kStages;
        const uint32_t par = j & 1;
        if (j + 1 < nblk) {
          mbar_wait(&s_free[t], par);  // this tile's softmax has pulled S(j) into registers
          mbar_wait(&k_full[(j + 1) % kStages], ((j + 1) / kStages) & 1);
          tc_fence_after();
          issue_qk(t, j + 1);
          umma_commit(&k_empty[(j + 1) % kStages]);
        }
        mbar_wait(&p_ready[t], par);
        mbar_wait(&v_full[st], (j / kStages) & 1);
        tc_fence_after();
        issue_pv(t, j);
        umma_commit(&v_empty[st]);
      }
    }
  }
  } else {
    // ---- softmax warpgroups: take them
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ===== softmax / correction / epilogue: tile t, one query row per thread =====
    const int t = warp >> 2;
    const int ew = warp & 3;
    const int r = ew * 32 + lane;
    const uint32_t lane_off = uint32_t(ew * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + t * kKT;
    const uint32_t tP = tmem_base + lane_off + kColP + t * 64;
    const uint32_t tO = tmem_base + lane_off + kColO + t * kHD;
    float m_ref = -INFINITY;
    float l = 0.f;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[t], j & 1);
      // start the two tiles half a period apart: tile B reads its first scores when tile A starts exponentiating,
      // so that from then on one tile's TMEM-read phase runs under the other tile's MUFU phase
      if (t == 1 && j == 0) mbar_wait(&s_free[0], 0);
      tc_fence_after();
      // One tcgen05.ld in flight at a time (a 32x32b.x32 load completes in ~27 cycles, profiles/r01_tmem_microbench2.log;
      // queueing four and waiting once measured slower end to end); the row max of chunk c runs under load c+1.
      uint32_t s[kKT];
      const int valid = p.S - j * kKT;  // < 128 only in the last block (TMA zero-filled the missing keys)
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      tmem_ld32p(tS, s);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < 3) tmem_ld32p(tS + (c + 1) * 32, s + (c + 1) * 32);
        if (valid < kKT) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) s[c * 32 + i] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(s[c * 32 + i]));
        if (c < 3) tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(&s_free[t]);  // S_t may be overwritten by QK_{j+1}
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
      if (j > 0) {
        // PV_{j-1} must have finished reading P_t before it is overwritten, and O_t must be complete for a rescale
        mbar_wait(&o_done[t], (j - 1) & 1);
        tc_fence_after();
      }
      if (j == 0) {
        m_ref = mx;
      } else {
        const bool grow = mx > m_ref + 8.0f;
        if (__any_sync(0xffffffffu, grow)) {
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
      // Packed fp32x2 arithmetic (Blackwell FFMA2/FADD2: two lanes per issue slot) for the scale-subtract, the row sum
      // and the polynomial; POLY of every 4 column PAIRS take the FMA-pipe exp2, the rest MUFU.EX2.
      const float2 c2 = make_float2(p.scale_log2, p.scale_log2);
      const float2 nm2 = make_float2(-m_ref, -m_ref);
      float2 acc2[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[c * 32 + i]), __uint_as_float(s[c * 32 + i + 1])), c2, nm2);
          float2 e;
          if (((i >> 1) & 3) < POLY) {
            e = exp2_poly2(x);
          } else {
            e.x = ex2(x.x);
            e.y = ex2(x.y);
          }
          acc2[(i >> 1) & 3] = __fadd2_rn(acc2[(i >> 1) & 3], e);
          pk[i >> 1] = pack_bf16x2(e.x, e.y);
        }
        tmem_st16(tP + c * 16, pk);
      }
      l += ((acc2[0].x + acc2[0].y) + (acc2[1].x + acc2[1].y)) + ((acc2[2].x + acc2[2].y) + (acc2[3].x + acc2[3].y));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[t]);
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
  if (warp == 9) {
    __syncwarp();
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
  auto kern = attn4_kernel<POLY>;
  static ::ea::PerDeviceFlag attr_flag;
  const int attr_dev = ::ea::current_device();
  if (!attr_flag.get(attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::kTotal);
    if (e != cudaSuccess) return fail(EA_ERR_CUDA, std::string("cudaFuncSetAttribute(attn4): ") + cudaGetErrorString(e));
    attr_flag.set(attr_dev);
  }
  dim3 grid((unsigned)((g->S + 2 * kQT - 1) / (2 * kQT)), (unsigned)BH);
  kern<<<grid, kThreads, Smem::kTotal, stream>>>(tq, tk, tv, p);
  count_launch();
  return check_launch("attn4_kernel");
}

}  // namespace a4

int launch_attn4(const ea_attn_args* g, int poly, cudaStream_t stream) {
  switch (poly) {
    case 0: return a4::launch<0>(g, stream);
    case 1: return a4::launch<1>(g, stream);
    case 2: return a4::launch<2>(g, stream);
    case 3: return a4::launch<3>(g, stream);
    case 4: return a4::launch<4>(g, stream);
    default: return fail(EA_ERR_INVALID, "ea_attn_fwd: unsupported polynomial fraction (0..4 of every 8)");
  }
}

}  // namespace ea
