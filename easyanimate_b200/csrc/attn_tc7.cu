// Joint text+video attention forward, seventh generation: the sixth generation's optimistic softmax (no per-block row
// maximum on the hot path, attn_tc6.cu) on the fifth generation's layout (every score tile's columns split over two
// warps, FOUR softmax warps per SM sub-partition, attn_tc5.cu).
//
// ncu's warp-state samples on the sixth generation: 40 % of a softmax warp's stall samples sit on MUFU.EX2 instructions
// (the warp is blocked at the MUFU queue while its FMA/ALU work waits behind it, in order); with two warps per scheduler
// nobody fills those slots.  Without the row maximum the two halves of a row no longer exchange anything on the hot
// path: they share m_ref by construction and only OR their end-of-block verdicts (one bar.red.or.pred on a 64-thread
// named barrier, which also replaces the warp vote).
//
//   TMEM (512 columns): S_t at 128 t | P_t (packed bf16) at 256 + 64 t | O_t at 384 + 64 t
//   warps 0-15 : softmax; warp w: tile t = w >> 3, column half h = (w >> 2) & 1, TMEM lane group g = w & 3
//   warp 16    : TMA producer;  warps 17/18: MMA issuer of tile A / B (17 also owns the TMEM allocation);  warp 19: idle
#include "common.cuh"
#include "host.h"
#include "../../include/ea_b200.h"

namespace ea {

extern void count_launch();

namespace a7 {

constexpr int kThreads = 640;
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
  static constexpr int kOffX = kOffV + kStages * kVBytes;       // float xch[2 parities][2 tiles][128 rows][2 halves]
  static constexpr int kOffBar = kOffX + 2 * 2 * 128 * 2 * 4;
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
EA_DEVICE void tmem_st32p(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
EA_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x (x <= 0) for a column pair on the FMA/ALU pipes: x = n + f (round to nearest), degree-3 minimax 2^f, exponent add
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

// Hot-loop mbarrier operations on precomputed 32-bit shared addresses.  ncu's source view of the fourth generation
// showed ~45 instructions per key block re-deriving the barrier addresses (S2R CgaCtaId / SWINHI, align, add) because
// the compiler rematerialises instead of keeping them; `opaque` hides the derivation so that they stay in registers.
EA_DEVICE uint32_t opaque(uint32_t x) {
  asm volatile("" : "+r"(x));
  return x;
}
EA_DEVICE void bar_wait(uint32_t addr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(addr),
      "r"(parity)
      : "memory");
}
EA_DEVICE void bar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}

// tcgen05.wait::ld that also "touches" the 32 destination registers of the load it completes, so that the compiler
// cannot schedule their first use above the wait (the load's asm statement already names them as outputs).
EA_DEVICE void tmem_ld_fence(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// Column pairs [P0, P1) of one 32-column chunk: e = 2^(s * c - m) with packed fp32x2 arithmetic, POLY of every 4 pairs on
// the FMA pipe (their raw scores also feed `guard`), the rest on MUFU; row-sum partials in acc2, packed bf16 pairs in pk.
template <int POLY, int P0, int P1>
EA_DEVICE void exp_pairs(const uint32_t* s, float2 c2, float2 nm2, float2* acc2, float& guard, uint32_t* pk) {
#pragma unroll
  for (int q = P0; q < P1; ++q) {
    const float s0 = __uint_as_float(s[2 * q]), s1 = __uint_as_float(s[2 * q + 1]);
    const float2 x = __ffma2_rn(make_float2(s0, s1), c2, nm2);
    float2 e;
    if ((q & 3) < POLY) {
      guard = fmaxf(guard, fmaxf(s0, s1));
      e = exp2_poly2(x);
    } else {
      e.x = ex2(x.x);
      e.y = ex2(x.y);
    }
    acc2[q & 3] = __fadd2_rn(acc2[q & 3], e);
    pk[q] = pack_bf16x2(e.x, e.y);
  }
}

// OR of a predicate over the two warps that own the same 32 rows (64-thread named barrier; also synchronises them)
EA_DEVICE bool pair_or(uint32_t bar_id, bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %1, 0;\n\t"
      "bar.red.or.pred p, %2, 64, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(out)
      : "r"((uint32_t)pred), "r"(bar_id)
      : "memory");
  return out != 0;
}

template <int POLY>
__global__ void __launch_bounds__(kThreads, 1)
attn7_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
             const __grid_constant__ CUtensorMap tmap_v, const Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + Smem::kOffQ;
  uint8_t* sK = smem + Smem::kOffK;
  uint8_t* sV = smem + Smem::kOffV;
  float* xch = reinterpret_cast<float*>(smem + Smem::kOffX);
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
      mbar_init(&k_empty[i], 2);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 256);
      mbar_init(&p_ready[i], 256);
      mbar_init(&o_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 17) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 16) {
    // ---- producer / MMA warpgroup: give registers back.  setmaxnreg.inc draws from the CTA's own pool (what its
    // warps released), not from the SM's free registers: 512 x 104 + 128 x 40 <= 640 x 96 (the launch allocation)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 16) {
      if (lane == 0) {
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
    } else if (warp <= 18) {
      if (lane == 0) {
        // ===== MMA issuers: warp 17 drives tile A, warp 18 tile B, independently (see attn_tc4.cu) =====
        const int t = warp - 17;
        constexpr uint32_t idesc_qk = umma_idesc_bf16(kQT, kKT, 0, 0);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(kQT, kHD, 0, 1);
        mbar_wait(q_full, 0);
        tc_fence_after();
        const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ + t * (kQT * kHD * 2)));
        auto issue_qk = [&](int j) {
          const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + (j % kStages) * Smem::kKBytes));
          const uint32_t d = tmem_base + t * kKT;
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
          umma_commit(&s_full[t]);
        };
        auto issue_pv = [&](int j) {
          const uint32_t vaddr = smem_u32(sV + (j % kStages) * Smem::kVBytes);
          const uint32_t d = tmem_base + kColO + t * kHD;
          const uint32_t pa = tmem_base + kColP + t * 64;
#pragma unroll
          for (int k = 0; k < kKT / 16; ++k)
            umma_ts(d, pa + k * 8, umma_desc_sw128_mn(vaddr + k * 2048, 16384, 1024), idesc_pv, (j | k) != 0);
          umma_commit(&o_done[t]);
        };
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        issue_qk(0);
        umma_commit(&k_empty[0]);
        for (int j = 0; j < nblk; ++j) {
          const int st = j % kStages;
          const uint32_t par = j & 1;
          if (j + 1 < nblk) {
            mbar_wait(&s_free[t], par);
            mbar_wait(&k_full[(j + 1) % kStages], ((j + 1) / kStages) & 1);
            tc_fence_after();
            issue_qk(j + 1);
            umma_commit(&k_empty[(j + 1) % kStages]);
          }
          mbar_wait(&p_ready[t], par);
          mbar_wait(&v_full[st], (j / kStages) & 1);
          tc_fence_after();
          issue_pv(j);
          umma_commit(&v_empty[st]);
        }
      }
    }
  } else {
    // ---- softmax warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int t = warp >> 3;        // tile
    const int h = (warp >> 2) & 1;  // column half
    const int g = warp & 3;         // TMEM lane group
    const int r = g * 32 + lane;    // row in tile
    const uint32_t lane_off = uint32_t(g * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + t * kKT + h * 64;
    const uint32_t tP = tmem_base + lane_off + kColP + t * 64 + h * 32;
    const uint32_t tO = tmem_base + lane_off + kColO + t * kHD + h * 32;
    const uint32_t pair_bar = 1 + t * 4 + g;  // named barrier shared by the two warps that own the same 32 rows
    const uint32_t bar_t = opaque(smem_u32(&s_full[t]));  // s_full[t]; s_free[t] +16, p_ready[t] +32, o_done[t] +48
    constexpr uint32_t kSFree = 16, kPReady = 32, kODone = 48;
    const float2 c2 = make_float2(p.scale_log2, p.scale_log2);
    float m_ref = 0.f;
    float l = 0.f;  // this half's partial row sum
    for (int j = 0; j < nblk; ++j) {
      bar_wait(bar_t, j & 1);
      if (t == 1 && j == 0) mbar_wait(&s_free[0], 0);  // half-period start offset between the tiles
      tc_fence_after();
      uint32_t s[64];
      uint32_t pk[16];
      const int valid = p.S - j * kKT - h * 64;  // valid columns of this half (>= 64 except in the last block)
      float2 acc2[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
      float guard = -INFINITY;
      bool redo = (j == 0) || (p.S - j * kKT < kKT);  // (the same for both halves of a row)
      tmem_ld32p(tS, s);
      tmem_ld32p(tS + 32, s + 32);
      tmem_ld_fence(s);
      tmem_ld_fence(s + 32);
      tc_fence_before();
      bar_arrive(bar_t + kSFree);  // S_t may be overwritten by QK_{j+1}
      if (!redo) {
        // ---- fast pass: exponentiate against the reference kept from earlier blocks (see attn_tc6.cu)
        const float2 nm2 = make_float2(-m_ref, -m_ref);
        exp_pairs<POLY, 0, 16>(s, c2, nm2, acc2, guard, pk);
        bar_wait(bar_t + kODone, (j - 1) & 1);  // PV_{j-1} has read P_t
        tc_fence_after();
        tmem_st16(tP, pk);
        exp_pairs<POLY, 0, 16>(s + 32, c2, nm2, acc2, guard, pk);
        tmem_st16(tP + 16, pk);
        const float l_blk = ((acc2[0].x + acc2[0].y) + (acc2[1].x + acc2[1].y)) + ((acc2[2].x + acc2[2].y) + (acc2[3].x + acc2[3].y));
        const bool ok = (l_blk <= 1073741824.0f) && (POLY == 0 || fmaf(guard, p.scale_log2, -m_ref) <= 64.0f);
        redo = pair_or(pair_bar, !ok);
        if (!redo) l += l_blk;
      } else if (valid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= valid) s[i] = 0xff800000u;  // -inf
      }
      if (redo) {
        // ---- classic pass from the registers: true row maximum (halves exchanged through shared memory), rescale this
        // half's 32 columns of O_t and its partial l, exponentiate again
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 64; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(s[i]));
        const float mh = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        float* xc = xch + (((j & 1) * 2 + t) * 128 + r) * 2;
        xc[h] = mh;
        named_bar_sync(pair_bar, 64);
        const float mx = fmaxf(mh, xc[h ^ 1]) * p.scale_log2;
        if (j == 0) {
          m_ref = mx;
        } else {
          bar_wait(bar_t + kODone, (j - 1) & 1);
          tc_fence_after();
          const float m_new = fmaxf(m_ref, mx);
          const float f = ex2(m_ref - m_new);
          uint32_t v[32];
          tmem_ld32p(tO, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
          tmem_st32p(tO, v);
          tmem_st_wait();
          l *= f;
          m_ref = m_new;
        }
        const float2 nm2 = make_float2(-m_ref, -m_ref);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc2[i] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          exp_pairs<POLY, 0, 16>(s, c2, nm2, acc2, guard, pk);
          tmem_st16(tP + c * 16, pk);
#pragma unroll
          for (int i = 0; i < 32; ++i) s[i] = s[i + 32];
        }
        l += ((acc2[0].x + acc2[0].y) + (acc2[1].x + acc2[1].y)) + ((acc2[2].x + acc2[2].y) + (acc2[3].x + acc2[3].y));
      }
      tmem_st_wait();
      tc_fence_before();
      bar_arrive(bar_t + kPReady);
    }
    // ---- epilogue: combine the two partial row sums, each half stores its 32 columns of O / l
    bar_wait(bar_t + kODone, (nblk - 1) & 1);
    tc_fence_after();
    float* xc = xch + (((nblk & 1) * 2 + t) * 128 + r) * 2;
    xc[h] = l;
    named_bar_sync(pair_bar, 64);
    const float inv_l = 1.0f / (l + xc[h ^ 1]);
    const int srow = q0 + t * kQT + r;
    uint32_t v[32];
    tmem_ld32p(tO, v);
    tmem_ld_wait();
    if (srow < p.S) {
      const int bb = bh / p.H, hh = bh % p.H;
      const int64_t d = (int64_t)p.H * kHD;
      bf16* dst = (srow < p.S_text) ? p.out_text + ((int64_t)bb * p.S_text + srow) * d
                                    : p.out_video + ((int64_t)bb * (p.S - p.S_text) + (srow - p.S_text)) * d;
      dst += hh * kHD + h * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(v[i * 8 + 0]) * inv_l, __uint_as_float(v[i * 8 + 1]) * inv_l);
        w.y = pack_bf16x2(__uint_as_float(v[i * 8 + 2]) * inv_l, __uint_as_float(v[i * 8 + 3]) * inv_l);
        w.z = pack_bf16x2(__uint_as_float(v[i * 8 + 4]) * inv_l, __uint_as_float(v[i * 8 + 5]) * inv_l);
        w.w = pack_bf16x2(__uint_as_float(v[i * 8 + 6]) * inv_l, __uint_as_float(v[i * 8 + 7]) * inv_l);
        *reinterpret_cast<uint4*>(dst + i * 8) = w;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
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
  auto kern = attn7_kernel<POLY>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::kTotal);
    if (e != cudaSuccess) return fail(EA_ERR_CUDA, std::string("cudaFuncSetAttribute(attn7): ") + cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)((g->S + 2 * kQT - 1) / (2 * kQT)), (unsigned)BH);
  kern<<<grid, kThreads, Smem::kTotal, stream>>>(tq, tk, tv, p);
  count_launch();
  return check_launch("attn7_kernel");
}

}  // namespace a7

int launch_attn7(const ea_attn_args* g, int poly, cudaStream_t stream) {
  switch (poly) {
    case 0: return a7::launch<0>(g, stream);
    case 1: return a7::launch<1>(g, stream);
    case 2: return a7::launch<2>(g, stream);
    case 3: return a7::launch<3>(g, stream);
    default: return fail(EA_ERR_INVALID, "ea_attn_fwd: unsupported polynomial fraction (0..3 of every 4 pairs)");
  }
}

}  // namespace ea
