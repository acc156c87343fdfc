// Causal 3x3x3 convolution (unit stride) as an implicit GEMM on tcgen05, second generation: ONE halo tile per (kt, 64-channel
// slice) in shared memory, the nine spatial taps read out of it through shared-memory descriptor offsets.
//
// conv3d_tc.cu fetches a separate 128-pixel x 64-channel TMA box for each of the 27 taps: 16 KB of A operand per k-block
// next to 8-16 KB of weights, and ncu (profiles/r01_ncu_conv3d_v1_summary.txt) shows the kernel bound by the L2 -> SM path
// (14.2 TB/s) with the tensor pipe 43.5 % active and DRAM reads 2.7-3.9x the algorithmic input.  Here an output sub-tile is
// 16 rows x 8 pixels (GEMM row r = pixel (r / 8, r % 8)), so an 8-row core-matrix group of the A operand is ONE tile row: in a
// halo tile of 18 x P pixels (P = 8 MT + 2, each pixel one 128-byte row of 64 channels, TMA SWIZZLE_128B) the rows of tap
// (kh, kw), sub-tile m start at pixel (kh, kw + 8 m) and consecutive groups are P pixels apart - a K-major SWIZZLE_128B
// descriptor with start = halo + (kh P + kw + 8 m) 128 B and stride-byte-offset = 128 P.  The swizzle is a function of the
// shared-memory address bits, so rows that do not start on a 1024-byte atom boundary read what TMA wrote there.
// A traffic per k-block drops from 16 KB to 18 P 128 / 9 B (2.5 KB at MT = 1, 4.6 KB at MT = 2); the k loop becomes
//   for kt: for 64-channel slice: [halo tile] for (kh, kw): [weight tile] MMA
// with separate rings for halo tiles (2-3 stages) and weight tiles (5-8 stages).  Tiles are walked frame-fastest (after the
// output-channel tiles) so that the three frames a temporal tap triple reads are shared by concurrently running CTAs in L2.
// MT sub-tiles sit side by side (CTA tile 16 x 8 MT pixels); CTA pairs (cta_group::2) take horizontally adjacent CTA tiles
// against one weight tile, half of it per CTA.  Epilogue (bias, residual, frame duplication, planar store) as conv3d_tc.cu.
//
// Replaces CausalConv3d.forward (easyanimate/vae/ldm/modules/vaemodules/common.py:84-141) for the decoder's and the
// encoder's stride-1 layers; the strided down-samplers stay on conv3d_tc.cu (TMA element strides).
#include "common.cuh"
#include "host.h"
#include "../../include/ea_b200.h"

namespace ea {

extern void count_launch();

namespace halo {

constexpr int kThreads = 256;
constexpr int kCK = 64;
constexpr int kTH = 16, kTW = 8;   // sub-tile: 16 rows x 8 pixels
constexpr int kHaloH = kTH + 2;

struct DevArgs {
  int T, H, W, Cin, Cout;
  int row0, Hout;                  // output row window [row0, row0 + Hout), stored at row h - row0 (ea_conv3d_args.out_row0)
  int tiles_h, tiles_w, tiles_n;   // PAIR: tiles_w counts PAIRS of CTA tiles
  int P;                           // halo pitch in pixels: 8 MT + 2, or that rounded up to a multiple of 8 (launch())
  const bf16* bias;
  const bf16* residual;
  bf16* out;
  int dup_frames, out_planar, T_out;
};

template <int BN, int MT, bool PAIR>
struct Cfg {
  static constexpr int kPmax = ((8 * MT + 2) + 7) / 8 * 8;
  static constexpr int kABytes = kHaloH * kPmax * 128;  // multiple of 1024
  static constexpr int kAStages = 2;
  static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * kCK * 2;
  static constexpr int kBudget = 216 * 1024;
  static constexpr int kBStagesRaw = (kBudget - kAStages * kABytes) / kBBytes;
  static constexpr int kBStages = kBStagesRaw > 9 ? 9 : kBStagesRaw;
  static constexpr int kAccCols = MT * BN;
  static constexpr int kTmemCols = 2 * kAccCols < 32 ? 32 : 2 * kAccCols;
  static constexpr int kSmemBytes = kAStages * kABytes + kBStages * kBBytes + 1024 + 512;
  static_assert(kTmemCols <= 512, "accumulators exceed TMEM");
  static_assert(kBStages >= 3, "weight ring too short");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory");
};

// K-major SWIZZLE_128B descriptor whose 8-row groups are `sbo_bytes` apart and whose start is any 128-byte row of the halo
// tile.  The base-offset field (bits 49-51) stays 0: the swizzle is applied to the absolute shared-memory address bits,
// filling the field with (start >> 7) & 7 reads the wrong rows (profiles/r02_conv_halo_bringup.log, variants 0x24 / 0x34).
EA_DEVICE uint64_t desc_halo(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int BN, int MT, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const DevArgs p) {
  using C = Cfg<BN, MT, PAIR>;
  constexpr int kAS = C::kAStages, kBS = C::kBStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kAS * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kBS * C::kBBytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kAS;
  uint64_t* b_full = a_empty + kAS;
  uint64_t* b_empty = b_full + kBS;
  uint64_t* tfull = b_empty + kBS;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.T * p.tiles_h * p.tiles_w * p.tiles_n;
  const int crank = PAIR ? (int)cluster_ctarank() : 0;
  const int first_tile = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int cchunks = p.Cin / kCK;
  const uint32_t a_tx = (uint32_t)(kHaloH * p.P * 128);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < kAS; ++i) {
      mbar_init(&a_full[i], PAIR ? 2 : 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kBS; ++i) {
      mbar_init(&b_full[i], PAIR ? 2 : 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], PAIR ? 256 : 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_2sm(tmem_slot, C::kTmemCols);
    else tmem_alloc(tmem_slot, C::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile order: output-channel tile fastest, then the frame, then the column and row of the pixel tile
  auto decode_tile = [&](int tile, int& t, int& h0, int& w0, int& n0) {
    n0 = (tile % p.tiles_n) * BN;
    int r = tile / p.tiles_n;
    t = r % p.T;
    r /= p.T;
    w0 = ((r % p.tiles_w) * (PAIR ? 2 : 1) + crank) * (MT * kTW);
    h0 = p.row0 + (r / p.tiles_w) * kTH;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        int t, h0, w0, n0;
        decode_tile(tile, t, h0, w0, n0);
        for (int kt = 0; kt < 3; ++kt) {
          int tin = t + kt - 2;  // causal: taps reach back in time; the left edge replicates frame 0
          tin = tin < 0 ? 0 : tin;
          for (int cc = 0; cc < cchunks; ++cc) {
            const int c0 = cc * kCK;
            mbar_wait(&a_empty[as], aph ^ 1);
            if (PAIR) {
              const uint32_t lfull = mapa_shared(smem_u32(&a_full[as]), 0);
              mbar_arrive_expect_tx_cluster(lfull, a_tx);
              tma_load_4d_2sm(smem_a + as * C::kABytes, &tmap_x, lfull, c0, w0 - 1, h0 - 1, tin);
            } else {
              mbar_arrive_expect_tx(&a_full[as], a_tx);
              tma_load_4d(smem_a + as * C::kABytes, &tmap_x, &a_full[as], c0, w0 - 1, h0 - 1, tin);
            }
            if (++as == kAS) { as = 0; aph ^= 1; }
            for (int tap = 0; tap < 9; ++tap) {
              const int kidx = (kt * 9 + tap) * p.Cin + c0;
              mbar_wait(&b_empty[bs], bph ^ 1);
              if (PAIR) {
                const uint32_t lfull = mapa_shared(smem_u32(&b_full[bs]), 0);
                mbar_arrive_expect_tx_cluster(lfull, C::kBBytes);
                tma_load_2d_2sm(smem_b + bs * C::kBBytes, &tmap_w, lfull, kidx, n0 + crank * (BN / 2));
              } else {
                mbar_arrive_expect_tx(&b_full[bs], C::kBBytes);
                tma_load_2d(smem_b + bs * C::kBBytes, &tmap_w, &b_full[bs], kidx, n0);
              }
              if (++bs == kBS) { bs = 0; bph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      // ===== MMA issuer (PAIR: the leader CTA issues for both) =====
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, BN);
      const uint32_t sbo = (uint32_t)p.P * 128u;
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t accph = (it >> 1) & 1;
        if (PAIR) mbar_wait_cluster(&tempty[acc], accph ^ 1);
        else mbar_wait(&tempty[acc], accph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * C::kAccCols;
        uint32_t first = 0;
        for (int kc = 0; kc < 3 * cchunks; ++kc) {
          if (PAIR) mbar_wait_cluster(&a_full[as], aph);
          else mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem_a + as * C::kABytes);
#pragma unroll 1
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              if (PAIR) mbar_wait_cluster(&b_full[bs], bph);
              else mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + bs * C::kBBytes));
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                const uint64_t adesc = desc_halo(a_base + (uint32_t)((kh * p.P + kw + 8 * m) * 128), sbo);
#pragma unroll
                for (int k = 0; k < kCK / 16; ++k) {
                  if (PAIR) umma_ss_2sm(tmem_d + m * BN, adesc + 2 * k, bdesc + 2 * k, idesc, first | (uint32_t)k);
                  else umma_ss(tmem_d + m * BN, adesc + 2 * k, bdesc + 2 * k, idesc, first | (uint32_t)k);
                }
              }
              first = 1;
              if (PAIR) umma_commit_2sm(&b_empty[bs], 0x3);
              else umma_commit(&b_empty[bs]);
              if (++bs == kBS) { bs = 0; bph ^= 1; }
            }
          }
          if (PAIR) umma_commit_2sm(&a_empty[as], 0x3);
          else umma_commit(&a_empty[as]);
          if (++as == kAS) { as = 0; aph ^= 1; }
        }
        if (PAIR) umma_commit_2sm(&tfull[acc], 0x3);
        else umma_commit(&tfull[acc]);
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: one output pixel per thread =====
    const int ew = warp - 4;
    const int r = ew * 32 + lane;
    const int ph = r >> 3, pw = r & 7;
    int it = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t accph = (it >> 1) & 1;
      int t, h0, w0, n0;
      decode_tile(tile, t, h0, w0, n0);
      mbar_wait(&tfull[acc], accph);
      tc_fence_after();
      const int t_out = p.dup_frames ? (t == 0 ? 0 : 2 * t - 1) : t;
      const int ncopies = (p.dup_frames && t > 0) ? 2 : 1;
#pragma unroll 1
      for (int cc = 0; cc < MT * (BN / 32); ++cc) {
        const int m = cc / (BN / 32), c = cc % (BN / 32);
        const int h = h0 + ph, w = w0 + m * kTW + pw;
        const bool pix_ok = h < p.row0 + p.Hout && w < p.W;
        const int ho = h - p.row0;
        const uint32_t trow = tmem_base + (uint32_t(ew * 32) << 16) + acc * C::kAccCols + m * BN;
        uint32_t av[32];
        __syncwarp();  // threads may have diverged on pix_ok / channel bounds in the previous chunk
        tmem_ld32(trow + c * 32, av);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (!pix_ok || col0 >= p.Cout) continue;
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(av[j]);
        if (p.out_planar) {
          // conv_out: few real channels, planar [Cout, T_out, H, W] store
          for (int j = 0; j < 32 && col0 + j < p.Cout; ++j) {
            const float v = x[j] + __bfloat162float(p.bias[col0 + j]);
            for (int cpy = 0; cpy < ncopies; ++cpy)
              p.out[(((int64_t)(col0 + j) * p.T_out + t_out + cpy) * p.Hout + ho) * p.W + w] = __float2bfloat16_rn(v);
          }
          continue;
        }
        {
          const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 b = __ldg(bp + j);
            float2 f0 = unpack_bf16x2(b.x), f1 = unpack_bf16x2(b.y), f2 = unpack_bf16x2(b.z), f3 = unpack_bf16x2(b.w);
            x[j * 8 + 0] += f0.x; x[j * 8 + 1] += f0.y; x[j * 8 + 2] += f1.x; x[j * 8 + 3] += f1.y;
            x[j * 8 + 4] += f2.x; x[j * 8 + 5] += f2.y; x[j * 8 + 6] += f3.x; x[j * 8 + 7] += f3.y;
          }
        }
        if (p.residual != nullptr) {
          // (conv2(x) + shortcut): the conv output is a bf16 tensor in the reference before the add (common.py:323)
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (((int64_t)t * p.Hout + ho) * p.W + w) * p.Cout + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 rr = __ldg(rp + j);
            uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float2 rf = unpack_bf16x2(rw[q]);
              x[j * 8 + 2 * q] = bf16_round(x[j * 8 + 2 * q]) + rf.x;
              x[j * 8 + 2 * q + 1] = bf16_round(x[j * 8 + 2 * q + 1]) + rf.y;
            }
          }
        }
        uint4 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j].x = pack_bf16x2(x[j * 8 + 0], x[j * 8 + 1]);
          o[j].y = pack_bf16x2(x[j * 8 + 2], x[j * 8 + 3]);
          o[j].z = pack_bf16x2(x[j * 8 + 4], x[j * 8 + 5]);
          o[j].w = pack_bf16x2(x[j * 8 + 6], x[j * 8 + 7]);
        }
        for (int cpy = 0; cpy < ncopies; ++cpy) {
          uint4* op = reinterpret_cast<uint4*>(p.out + ((((int64_t)(t_out + cpy)) * p.Hout + ho) * p.W + w) * p.Cout + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) op[j] = o[j];
        }
      }
      tc_fence_before();
      if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[acc]), 0));
      else mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync();  // no CTA leaves while its peer may still signal it or read its shared memory
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, C::kTmemCols);
    else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

template <int BN, int MT, bool PAIR>
static int launch(const ea_conv3d_args* g, cudaStream_t stream) {
  using C = Cfg<BN, MT, PAIR>;
  DevArgs p{};
  p.T = (int)g->T; p.H = (int)g->H; p.W = (int)g->W;
  p.Cin = (int)g->Cin; p.Cout = (int)g->Cout;
  p.row0 = g->out_rows > 0 ? (int)g->out_row0 : 0;
  p.Hout = g->out_rows > 0 ? (int)g->out_rows : p.H;
  const int cta_w = MT * kTW * (PAIR ? 2 : 1);
  p.tiles_h = (p.Hout + kTH - 1) / kTH;
  p.tiles_w = (p.W + cta_w - 1) / cta_w;
  p.tiles_n = (int)((g->Cout_pad + BN - 1) / BN);
  // halo pitch: the natural 8 MT + 2 pixels, or rounded up to a multiple of 8 (the box then fetches up to 6 unused columns).
  // Measured (profiles/r02_conv_halo_bringup.log): no difference at MT = 1 (10 vs 16), the 24-pixel pitch is 7 % faster than
  // 18 at MT = 2 (128 -> 128 channels at 720 x 1280: 1 368 vs 1 278 TFLOP/s).  variant bit4 takes the other choice (A/B).
  const bool padded = (MT == 2) != ((g->variant & 0x10) != 0);
  p.P = padded ? C::kPmax : 8 * MT + 2;
  p.bias = reinterpret_cast<const bf16*>(g->bias);
  p.residual = reinterpret_cast<const bf16*>(g->residual);
  p.out = reinterpret_cast<bf16*>(g->out);
  p.dup_frames = g->dup_frames;
  p.out_planar = g->out_planar;
  p.T_out = g->dup_frames ? (int)(2 * g->T - 1) : p.T;

  CUtensorMap tx, tw;
  {
    uint64_t dims[4] = {(uint64_t)g->Cin, (uint64_t)g->W, (uint64_t)g->H, (uint64_t)g->T};
    uint64_t strides[3] = {(uint64_t)g->Cin * 2, (uint64_t)g->W * g->Cin * 2, (uint64_t)g->H * g->W * g->Cin * 2};
    uint32_t box[4] = {kCK, (uint32_t)p.P, (uint32_t)kHaloH, 1};
    int rc = make_tmap_bf16(&tx, g->x, 4, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    const uint64_t K = (uint64_t)27 * g->Cin;
    uint64_t dims[2] = {K, (uint64_t)g->Cout_pad};
    uint64_t strides[1] = {K * 2};
    uint32_t box[2] = {kCK, (uint32_t)(PAIR ? BN / 2 : BN)};
    int rc = make_tmap_bf16(&tw, g->w, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  auto kern = conv3d_halo_kernel<BN, MT, PAIR>;
  static ::ea::PerDeviceFlag attr_flag;
  const int attr_dev = ::ea::current_device();
  if (!attr_flag.get(attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return fail(EA_ERR_CUDA, std::string("cudaFuncSetAttribute(conv halo): ") + cudaGetErrorString(e));
    attr_flag.set(attr_dev);
  }
  const int64_t num_tiles = (int64_t)p.T * p.tiles_h * p.tiles_w * p.tiles_n;
  if (num_tiles >= (1ll << 31)) return fail(EA_ERR_INVALID, "ea_conv3d: too many tiles");
  if (PAIR) {
    const int clusters = (int)(num_tiles < sm_count() / 2 ? num_tiles : sm_count() / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tx, tw, p);
    if (e != cudaSuccess) return fail(EA_ERR_CUDA, std::string("cudaLaunchKernelEx(conv halo, CTA pairs): ") + cudaGetErrorString(e));
    count_launch();
    return check_launch("conv3d_halo_kernel (CTA pairs)");
  }
  const int grid = (int)(num_tiles < sm_count() ? num_tiles : sm_count());
  kern<<<grid, kThreads, C::kSmemBytes, stream>>>(tx, tw, p);
  count_launch();
  return check_launch("conv3d_halo_kernel");
}

}  // namespace halo

// Dispatch for unit-stride calls (ea_conv3d_causal has validated the arguments).  Same tiling rules as conv3d_tc.cu:
// 256-pixel CTA tiles once the frame fills the SMs twice over, CTA pairs from four waves of pair tiles.
int launch_conv3d_halo(const ea_conv3d_args* g, cudaStream_t stream) {
  const int64_t Hw = g->out_rows > 0 ? g->out_rows : g->H;  // rows actually computed
  const int64_t tiles256 = g->T * ((Hw + 15) / 16) * ((g->W + 15) / 16) * ((g->Cout_pad + 127) / 128);
  const bool big = tiles256 >= 2 * sm_count() && !(g->variant & 1);
  const bool pairs = big && !(g->variant & 2) && tiles256 >= 8 * sm_count();
  if (pairs && g->Cout_pad % 256 == 0) return halo::launch<256, 1, true>(g, stream);
  if (pairs && g->Cout_pad % 128 == 0) return halo::launch<128, 2, true>(g, stream);
  if (g->Cout_pad % 128 == 0) return big ? halo::launch<128, 2, false>(g, stream) : halo::launch<128, 1, false>(g, stream);
  if (g->Cout_pad % 64 == 0) return big ? halo::launch<64, 2, false>(g, stream) : halo::launch<64, 1, false>(g, stream);
  return big ? halo::launch<32, 2, false>(g, stream) : halo::launch<32, 1, false>(g, stream);
}

}  // namespace ea
