// 3x3 stride-1 pad-1 dense convolution with a shared-memory HALO tile: the input is staged ONCE per tile and the nine
// filter taps are nine shifted views of the same shared-memory rows.
//
// Why (round-1 ncu + per-layer model, profiles/): with TMA-im2col every 128-pixel A tile is re-fetched once per tap,
// so a 3x3 layer pulls 9x its input through the L2 -> SMEM fabric; all 3x3 layers of ResNet-50 sat at 9-10 TB/s of
// smem fill (the measured ~6300 B/clk LTS cap), 2-4x above their tensor / HBM bounds.  Here the fill traffic is
// (R+2)/R x the input, the weights never move, and the layer becomes tensor / HBM bound.
//
// Layout trick: a tile is R full output rows of one image.  Its input box - (R+2) rows x (W+2) pixels x 64 channels,
// fetched by ONE 4-D tiled TMA load starting at (w = -1, h = h0-1) so the zero padding is materialised by the TMA
// unit's out-of-bounds fill - lands in shared memory as a dense list of 128-byte pixel rows (SWIZZLE_128B), i.e.
// exactly the canonical K-major operand layout with "row" = padded pixel index.  Output pixel q = r*(W+2) + c of the
// tile needs, for tap (fr, fs), input row q + fr*(W+2) + fs: the tap is a constant row offset, so the A operand of
// tap (fr, fs) and M-block j is the SAME buffer with the descriptor start address advanced by
// (j*128 + fr*(W+2) + fs) * 128 bytes (the 128B swizzle is a function of the absolute smem address, which is what
// makes the K-advance of +32 B work as well).  Columns c >= W of each row are computed and dropped (2/(W+2) waste).
//
// CTA pair (cta_group::2, one 256 x BN MMA per K=16 step): each CTA owns its own tile (128 A rows per M-block from
// each CTA) and HALF of the output channels' weights, which stay RESIDENT in shared memory for the whole kernel
// (9 * Cin/64 blocks of [BN/2 x 64]).  Roles per CTA: warp 0 producer (weights once, then one A box per 64-channel
// block into a ring), warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator, warps 4-7 epilogue
// (tcgen05.ld -> +bias -> clamp -> bf16 -> direct 64-byte-per-thread global stores; garbage columns masked).
//
// Output path (measured with the PCV_IGEMM3_DBG knobs, profiles/README.md): with BN = 64 the MMA is bound by the
// shared-memory operand read rate (A 4 KB + B 1 KB per 32-cycle MMA), and a direct epilogue - one pixel per thread, so
// every STG.128 touches 32 different 128-byte lines - costs the same L1/shared-memory port ~128 cycles per warp
// instruction: MMA-only 0.063 ms, epilogue-only 0.060 ms, together 0.089 ms.  STAGED = true stages the tile in shared
// memory as R image rows of round_up(W, 8) swizzled pixel rows and writes each image row with one 4-D TMA store
// (garbage columns never staged), double-buffered; used for BN = 64 whenever the staging buffers fit.
#include "igemm_common.cuh"

namespace pcv {
namespace PCV_TIER {

struct Igemm3Params {
  const float* bias;
  e16* out;
  int out_pitch;
  int N, H, W, Cout;
  int R, PW, NMB;            // output rows per tile, padded width W+2, 128-row M-blocks per tile
  int cblocks;               // Cin / 64 (dense) or 1 (grouped: every 64-channel block is its own 64 -> 64 conv)
  int G;                     // 64-channel group blocks (1 for dense); pair tiles iterate (spatial pair, g), g fastest
  int tiles_per_img, num_tiles;
  int NA;                    // A ring depth
  int a_buf_bytes;           // smem stride of one A buffer (multiple of 1024, includes the over-read slack)
  int a_tx_bytes;            // bytes of one A box: (R+2) * PW * 128
  int b_block_bytes;         // BN/2 * 128
  int WP8, stg_bytes;        // STAGED: pixels per staged image row (W rounded up to 8), bytes of one staging buffer
  float act_lo, act_hi;
  int dbg;                   // PCV_IGEMM3_DBG throughput experiments: 1 skip MMA issue, 2 skip epilogue math+stores, 4 skip A loads
  int sub16;                 // grouped layers with <= 16 channels per group: the 64 x 64 block-diagonal weight block is four
                             // 16 x 16 diagonal blocks, multiplied as four N = 16 MMAs (one K = 16 step each) instead of four
                             // N = 64 MMAs over mostly-zero weights; a CTA's share of sub-block j is rows 16 j + 8 rank .. + 8
};

__device__ __forceinline__ void tma2_load_4d(const CUtensorMap* m, uint32_t mbar_cluster_addr, void* dst, int c0,
                                             int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

constexpr int I3_MAX_NA = 6;
constexpr int I3_THREADS = 384;   // 4 control warps + 8 epilogue warps

template <int BN, bool STAGED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I3_THREADS, 1)
igemm3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const __grid_constant__ CUtensorMap tmOut, const Igemm3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nb_blocks = 9 * p.cblocks * p.G;
  uint8_t* sB = smem;                                         // resident weights: nb_blocks x [BN/2 x 128 B]
  uint8_t* sA = sB + ((nb_blocks * p.b_block_bytes + 1023) & ~1023);
  uint8_t* sStg = sA + p.NA * p.a_buf_bytes;                  // STAGED: 2 output staging buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + (STAGED ? 2 * p.stg_bytes : 0));
  uint64_t* full = bars;                         // [NA]  leader's copy is live
  uint64_t* empty = bars + I3_MAX_NA;            // [NA]  per CTA, multicast commit
  uint64_t* b_full = bars + 2 * I3_MAX_NA;       // [1]   leader's copy
  uint64_t* tmem_full = b_full + 1;              // [2]   per CTA, multicast commit
  uint64_t* tmem_empty = tmem_full + 2;          // [2]   leader's copy, 8 arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  // role index: physical warps 8-11 are the control roles 0-3 (the scheduler prefers high warp ids), physical warps
  // 0-7 the epilogue roles 4-11 (TMEM lane quarter = role & 3 = physical warp & 3)
  const int pw = threadIdx.x >> 5;
  const int warp = pw >= 8 ? pw - 8 : pw + 4;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int pair_tiles = ((p.num_tiles + 1) >> 1) * p.G;
  const int acc_cols = p.NMB * BN;               // TMEM columns of one accumulator buffer
  const uint32_t tmem_cols = 2 * acc_cols <= 256 ? 256u : 512u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (STAGED) tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.NA; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 16);   // 8 epilogue warps x 2 CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_ptr, tmem_cols);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (tmem_base != 0) __trap();   // one CTA per SM (227 KB smem) => the allocation starts at column 0; the MMA warp relies on it
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================== producer (both CTAs) =====================================
    {   // whole warp; TMA / mbarrier instructions under elect.sync (uniform-register operands)
      const uint32_t bfull_leader = mapa_u32(smem_u32(b_full), 0);
      if (rank == 0 && elect_one()) mbar_arrive_expect_tx(b_full, 2 * nb_blocks * p.b_block_bytes);
      for (int b = 0; b < nb_blocks; ++b) {
        const int g = b / (9 * p.cblocks), kb = b - g * 9 * p.cblocks;
        if (p.sub16) {   // four 8-row boxes: this CTA's half of each 16 x 16 diagonal sub-block
          for (int j = 0; j < 4; ++j)
            if (elect_one())
              tma2_load_2d(&tmB, bfull_leader, sB + b * p.b_block_bytes + j * 1024, kb * BLOCK_K,
                           g * BN + 16 * j + 8 * static_cast<int>(rank));
        } else if (elect_one()) {
          tma2_load_2d(&tmB, bfull_leader, sB + b * p.b_block_bytes, kb * BLOCK_K, g * BN + static_cast<int>(rank) * (BN / 2));
        }
      }
      pdl_wait();   // weights above do not depend on the previous kernel; the activations below do
      int slot = 0;
      uint32_t phase = 0;
      for (int t = pair; t < pair_tiles; t += npairs) {
        const int sp = t / p.G, g = t - sp * p.G;
        int tile = 2 * sp + static_cast<int>(rank);
        if (tile >= p.num_tiles) tile = p.num_tiles - 1;   // phantom tile of an odd tail: recompute the last one, stores masked
        const int img = tile / p.tiles_per_img;
        const int h0 = (tile - img * p.tiles_per_img) * p.R;
        for (int cb = 0; cb < p.cblocks; ++cb) {
          mbar_wait(&empty[slot], phase ^ 1);
          const uint32_t full_leader = mapa_u32(smem_u32(&full[slot]), 0);
          if (elect_one()) {
            if (p.dbg & 4) {
              if (rank == 0) mbar_arrive(&full[slot]);
            } else {
              if (rank == 0) mbar_arrive_expect_tx(&full[slot], 2 * p.a_tx_bytes);
              tma2_load_4d(&tmA, full_leader, sA + slot * p.a_buf_bytes, (g + cb) * BLOCK_K, -1, h0 - 1, img);
            }
          }
          if (++slot == p.NA) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =====================================
    // The WHOLE warp runs this loop with warp-uniform values and only the tcgen05 instructions are predicated on
    // elect.sync: ptxas then keeps descriptors in uniform registers.  (Inside an `if (lane == 0)` region every
    // UTCHMMA was wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY "waterfall" of ~13 dependent instructions, which
    // capped the issue rate at one MMA per ~120 cycles whatever its N.)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_e16(2 * BLOCK_M, BN);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA)), b_lo0 = smem_desc_lo(smem_u32(sB));
      const uint32_t b_step = p.b_block_bytes >> 4;
      mbar_wait(b_full, 0);
      tc_fence_after();
      int slot = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair; t < pair_tiles; t += npairs, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[buf], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = buf * acc_cols;     // TMEM base is 0: this CTA owns the SM's whole TMEM (checked above)
        const int g = t % p.G;
        for (int cb = 0; cb < p.cblocks; ++cb) {
          mbar_wait(&full[slot], phase);
          tc_fence_after();
          const uint32_t a_buf = a_lo0 + slot * (p.a_buf_bytes >> 4);
          // All 9 taps x 4 K-steps of one M-block inside ONE elect.sync region with compile-time tap indices: descriptors
          // are formed with uniform adds and the 36 UTCHMMAs issue back to back.  (One region per (tap, M-block) - 4 MMAs
          // behind ~30 R2UR / UMOV instructions - left the N = 64 / N = 128 MMAs issue-bound at 78 / 123 cycles each,
          // against 60 for the stem kernel's shared-memory-bound stream.)
          const uint32_t row_step = p.PW * (128 >> 4);
          const uint32_t tap_step = p.cblocks * b_step;
          const uint32_t b_cb = b_lo0 + (g * 9 * p.cblocks + cb) * b_step;
          for (int mb = 0; mb < ((p.dbg & 1) ? 0 : p.NMB); ++mb) {
            const uint32_t a_mb = a_buf + mb * (BLOCK_M * 128 >> 4);
            const uint32_t d_mb = d_tmem + mb * BN;
            if (p.sub16) {
              // four 16 x 16 diagonal sub-blocks: MMA j multiplies K step j of the pixels by rows 16 j .. 16 j + 15 of the
              // weights into accumulator columns 16 j .. 16 j + 15
              constexpr uint32_t idesc16 = make_idesc_e16(2 * BLOCK_M, 16);
              if (elect_one()) {
#pragma unroll
                for (int fr = 0; fr < 3; ++fr) {
                  const uint32_t a_row = a_mb + fr * row_step;
#pragma unroll
                  for (int fs = 0; fs < 3; ++fs) {
                    const uint32_t a_lo = a_row + fs * (128 >> 4);
                    const uint32_t b_lo = b_cb + (fr * 3 + fs) * tap_step;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                      umma2_bf16_lohi(d_mb + 16 * j, a_lo + 2 * j, b_lo + j * (1024 >> 4) + 2 * j, idesc16, (fr | fs) != 0 ? 1u : 0u);
                  }
                }
              }
            } else if (elect_one()) {
#pragma unroll
              for (int fr = 0; fr < 3; ++fr) {
                const uint32_t a_row = a_mb + fr * row_step;
#pragma unroll
                for (int fs = 0; fs < 3; ++fs) {
                  const uint32_t a_lo = a_row + fs * (128 >> 4);
                  const uint32_t b_lo = b_cb + (fr * 3 + fs) * tap_step;
#pragma unroll
                  for (int k = 0; k < BLOCK_K / 16; ++k)
                    umma2_bf16_lohi(d_mb, a_lo + 2 * k, b_lo + 2 * k, idesc, (fr | fs | k) != 0 ? 1u : (cb != 0 ? 1u : 0u));
                }
              }
            }
          }
          if (elect_one()) umma2_commit(&empty[slot], 0x3);
          if (++slot == p.NA) {
            slot = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma2_commit(&tmem_full[buf], 0x3);
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue (both CTAs) =====================================
    // eight warps: warp w owns TMEM lane quarter (w & 3) and every second 32-column chunk starting at (w - 4) >> 2
    pdl_wait();   // the output buffer may alias a tensor the previous kernel is still reading
    const int q4 = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q4 * 32 + lane;
    const bool relu = p.act_lo == 0.f, capped = p.act_hi != INFINITY;   // clamp family only (igemm3_try_make)
    const uint32_t cap2 = pack_e16x2(p.act_hi, p.act_hi);
    const uint32_t tmem_empty_leader0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    const uint32_t tmem_empty_leader1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
    const bool storer = STAGED && warp == 4 && lane == 0;   // owns every bulk store group of this CTA
    int it = 0;
    for (int t = pair; t < pair_tiles; t += npairs, ++it) {
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      uint8_t* stg = sStg + (STAGED ? buf * p.stg_bytes : 0);   // free: see the wait before named barrier 1 below
      const int sp = t / p.G, g = t - sp * p.G;
      const int tile = 2 * sp + static_cast<int>(rank);
      const bool tile_ok = tile < p.num_tiles;
      const int img = tile / p.tiles_per_img;
      const int h0 = (tile - img * p.tiles_per_img) * p.R;

      mbar_wait(&tmem_full[buf], acc_phase);
      tc_fence_after();

      for (int mb = 0; mb < ((p.dbg & 2) ? 0 : p.NMB); ++mb) {
        const int q = mb * BLOCK_M + row;
        const int r = q / p.PW;
        const int c = q - r * p.PW;
        const bool ok = tile_ok && c < p.W && r < p.R && (h0 + r) < p.H;
        e16* dst = p.out + (static_cast<size_t>(img * p.H + h0 + r) * p.W + c) * p.out_pitch + g * BN;
#pragma unroll 1
        for (int j = half; j < BN / 32; j += 2) {
          uint32_t acc[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + buf * acc_cols + mb * BN + j * 32, acc);
          // bias loads issued under the TMEM load (the wait below is a compiler barrier for memory operations)
          const float4* bias4 = reinterpret_cast<const float4*>(p.bias + g * BN + j * 32);
          float4 b4[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) b4[i] = __ldg(bias4 + i);
          tmem_ld_wait_regs(acc);
          // scalar fp32 adds on purpose: packing the tcgen05.ld registers into 64-bit operands for add.f32x2 cost more
          // instructions than it saved (measured on the igemm2 epilogue)
          float v[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[4 * i + 0] = __uint_as_float(acc[4 * i + 0]) + b4[i].x;
            v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + b4[i].y;
            v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + b4[i].z;
            v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + b4[i].w;
          }
          uint32_t o[16];
          clamp_pack32(v, o, relu, capped, cap2);
          if (ok) {
            if (STAGED) {
              const uint32_t srow = r * p.WP8 + c;   // staged pixel index; staged image rows start on 1024-byte boundaries
              const uint32_t dstp = smem_u32(stg) + srow * (BN * 2);
              const uint32_t sw = srow & 7u;         // SWIZZLE_128B chunk XOR (BN = 64: one 128-byte row per pixel)
#pragma unroll
              for (int i = 0; i < 4; ++i)
                sts128(dstp + (((j * 4 + i) ^ sw) << 4), o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            } else {
              uint4* d4 = reinterpret_cast<uint4*>(dst + j * 32);
#pragma unroll
              for (int i = 0; i < 4; ++i) d4[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tmem_empty[buf]);
        else mbar_arrive_cluster(buf ? tmem_empty_leader1 : tmem_empty_leader0);
      }
      if (STAGED) {
        fence_proxy_async_smem();                   // this thread's st.shared -> visible to the TMA (async proxy)
        if (storer) tma_store_wait_read<0>();       // the previous tile's stores have left the OTHER buffer (next tile's)
        named_bar_sync(1, 256);                     // all 8 epilogue warps: tile staged, other buffer free
        if (storer && tile_ok) {
          for (int r = 0; r < p.R; ++r)
            if (h0 + r < p.H) tma_store_4d(&tmOut, stg + r * p.WP8 * (BN * 2), g * BN, 0, h0 + r, img);
          tma_store_commit();
        }
      }
    }
    if (storer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct Igemm3Op : Op {
  CUtensorMap tmA, tmB, tmOut;
  Igemm3Params p;
  int bn, grid, smem_bytes;
  bool staged = false;
  cudaError_t launch(cudaStream_t s) override;
};

template <int BN, bool STAGED>
static cudaError_t launch_i3(const Igemm3Op& op, cudaStream_t s) {
  static std::atomic<uint64_t> attr_done{0};   // per device (see runtime.h)
  if (cudaError_t e = set_max_smem_once(igemm3_kernel<BN, STAGED>, 232448, attr_done)) return e;
  return launch_pdl(igemm3_kernel<BN, STAGED>, dim3(op.grid), dim3(I3_THREADS), op.smem_bytes, s, op.tmA, op.tmB,
                    op.tmOut, op.p);
}

cudaError_t Igemm3Op::launch(cudaStream_t s) {
  g_launches++;
  if (bn == 64) return staged ? launch_i3<64, true>(*this, s) : launch_i3<64, false>(*this, s);
  return launch_i3<128, false>(*this, s);
}

// Returns PCV_ERR_UNSUPPORTED (message untouched) when the layer is outside this kernel's domain.
int igemm3_try_make(const pcv_conv_desc& d, const void* x, const void* w, const float* bias, const void* res, void* y,
                    Op** out) {
  static const bool enabled = [] {
    const char* e = getenv("PCV_IGEMM_HALO");
    return !(e && e[0] == '0');
  }();
  const int in_pitch = pitch_or(d.in_pitch, d.Cin), out_pitch = pitch_or(d.out_pitch, d.Cout);
  const bool grouped = d.groups > 1;
  if (grouped && (d.Cin != d.Cout || d.Cin % 64 != 0 || 64 % (d.Cin / d.groups) != 0)) return PCV_ERR_UNSUPPORTED;
  if (!enabled || res || d.kh != 3 || d.kw != 3 || d.stride != 1 || d.dil != 1 || d.pad != 1 ||
      d.flags != 0 || d.in_row_pitch != 0 || d.Cin % 64 != 0 || (!grouped && d.Cout != 64 && d.Cout != 128) ||
      in_pitch % 8 != 0 || out_pitch % 8 != 0 || d.W + 2 > 256 || d.W < (grouped ? 14 : 24) || d.act > PCV_ACT_RELU6 ||
      reinterpret_cast<uintptr_t>(x) % 16 != 0 || reinterpret_cast<uintptr_t>(y) % 16 != 0 ||
      reinterpret_cast<uintptr_t>(w) % 16 != 0)
    return PCV_ERR_UNSUPPORTED;
  // grouped (ResNeXt, resnext.py:44-55): every 64-channel block is an independent 64 -> 64 convolution with
  // block-diagonal weights; all G blocks' weights stay resident
  const int G = grouped ? d.Cin / 64 : 1;
  const int BN = grouped ? 64 : d.Cout, PW = d.W + 2, cblocks = grouped ? 1 : d.Cin / 64;
  const int b_block = BN / 2 * 128;
  const int b_bytes = round_up(9 * cblocks * G * b_block, 1024);
  const int budget = 232448 - 1024 - 256 - b_bytes;
  if (budget <= 0) return PCV_ERR_UNSUPPORTED;
  // rows per tile: maximise (useful MMA rows) x (wave efficiency) among the tiles that fit; BN = 64 first tries to fit
  // two output staging buffers as well (STAGED, see the header) and falls back to direct stores when no tile fits
  static const bool stage_enabled = [] {
    const char* e = getenv("PCV_IGEMM3_STAGED");
    return !(e && e[0] == '0');
  }();
  const int pairs = sm_count() / 2;
  const int WP8 = round_up(d.W, 8);
  double best = 0.0;
  int bestR = 0, bestNA = 0, bestNMB = 0, best_buf = 0, best_stg = 0;
  bool staged = false;
  for (int pass = (BN == 64 && stage_enabled) ? 0 : 1; pass < 2 && bestR == 0; ++pass) {
    staged = pass == 0;
    for (int R = 1; R <= std::min(d.H, 32); ++R) {
      const int Q = R * PW, NMB = ceil_div(Q, BLOCK_M);
      if (2 * NMB * BN > 512 || R + 2 > 256) continue;
      const int buf = round_up((NMB * BLOCK_M + 2 * PW + 2) * 128, 1024);
      const int stg = staged ? round_up(R * WP8 * BN * 2, 1024) : 0;
      const int NA = std::min(I3_MAX_NA, (budget - 2 * stg) / buf);
      if (NA < 2) continue;
      const int tiles = d.N * ceil_div(d.H, R);
      const int pair_tiles = (tiles + 1) / 2 * G;
      const double wave = static_cast<double>(pair_tiles) / (ceil_div(pair_tiles, pairs) * pairs);
      const double rows = static_cast<double>(d.H) * d.W / (static_cast<double>(ceil_div(d.H, R)) * NMB * BLOCK_M);
      const double halo = static_cast<double>(R) / (R + 2);
      const double work = static_cast<double>(NMB) * cblocks * 36 * (BN / 2);   // tensor-pipe cycles per tile
      const double score = wave * rows * (0.9 + 0.1 * halo) * work / (work + 500.0);  // ~500-cycle handshake per tile
      if (score > best) {
        best = score; bestR = R; bestNA = NA; bestNMB = NMB; best_buf = buf; best_stg = stg;
      }
    }
  }
  if (bestR == 0) return PCV_ERR_UNSUPPORTED;

  auto op = std::make_unique<Igemm3Op>();
  Igemm3Params& p = op->p;
  p.bias = bias;
  p.out = reinterpret_cast<e16*>(y);
  p.out_pitch = out_pitch;
  p.N = d.N; p.H = d.H; p.W = d.W; p.Cout = d.Cout;
  p.R = bestR; p.PW = PW; p.NMB = bestNMB; p.cblocks = cblocks; p.G = G;
  p.tiles_per_img = ceil_div(d.H, bestR);
  p.num_tiles = d.N * p.tiles_per_img;
  p.NA = bestNA;
  p.a_buf_bytes = best_buf;
  p.a_tx_bytes = (bestR + 2) * PW * 128;
  p.b_block_bytes = b_block;
  {
    // PCV_IGEMM3_SUB16=0: the plain N = 64 block-diagonal MMAs
    const char* e = getenv("PCV_IGEMM3_SUB16");
    const int cg = grouped ? d.Cin / d.groups : 0;
    p.sub16 = (grouped && cg <= 16 && 16 % cg == 0 && !(e && e[0] == '0')) ? 1 : 0;
  }
  p.WP8 = WP8;
  p.stg_bytes = best_stg;
  p.act_lo = (d.act == PCV_ACT_RELU || d.act == PCV_ACT_RELU6) ? 0.f : -INFINITY;
  p.act_hi = d.act == PCV_ACT_RELU6 ? 6.f : INFINITY;
  {
    const char* e = getenv("PCV_IGEMM3_DBG");
    p.dbg = e ? atoi(e) : 0;
  }
  op->bn = BN;
  op->staged = staged;
  op->smem_bytes = 1024 + b_bytes + bestNA * best_buf + 2 * best_stg + 256;
  op->grid = 2 * std::min((p.num_tiles + 1) / 2 * G, pairs);

  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  {
    cuuint64_t dims[4] = {(cuuint64_t)d.Cin, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.N};
    cuuint64_t strides[3] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)d.W * in_pitch * 2, (cuuint64_t)d.H * d.W * in_pitch * 2};
    cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)PW, (cuuint32_t)(bestR + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&op->tmA, TMAP_E16, 4, const_cast<void*>(x), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled (halo A) failed (%d)", (int)r);
  }
  {
    const uint64_t kpad = 9ull * cblocks * BLOCK_K;   // grouped: [Cout, 9 * 64] block-diagonal rows
    cuuint64_t dims[2] = {kpad, (cuuint64_t)d.Cout};
    cuuint64_t strides[1] = {kpad * 2};
    cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)(p.sub16 ? 8 : BN / 2)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&op->tmB, TMAP_E16, 2, const_cast<void*>(w), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled (halo B) failed (%d)", (int)r);
  }
  op->tmOut = op->tmB;
  if (staged) {   // output rows leave through 4-D TMA stores: box = one image row (W pixels x 64 channels)
    cuuint64_t dims[4] = {(cuuint64_t)d.Cout, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.N};
    cuuint64_t strides[3] = {(cuuint64_t)out_pitch * 2, (cuuint64_t)d.W * out_pitch * 2, (cuuint64_t)d.H * d.W * out_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)BN, (cuuint32_t)d.W, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&op->tmOut, TMAP_E16, 4, y, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled (halo out) failed (%d)", (int)r);
  }
  char nm[160];
  snprintf(nm, sizeof nm, "conv_tc3 3x3 s1 d1 g%d %d->%d @%dx%d bn=%d halo R=%d mb=%d na=%d%s", d.groups, d.Cin, d.Cout, d.H,
           d.W, BN, bestR, bestNMB, bestNA, staged ? " staged" : "");
  op->name = nm;
  const double M = static_cast<double>(d.N) * d.H * d.W;
  op->flops = 2.0 * M * d.Cout * (d.Cin / d.groups) * 9;
  op->bytes = 2.0 * d.N * d.Cin * d.H * d.W + 2.0 * M * d.Cout + 2.0 * d.Cout * (d.Cin / d.groups) * 9 + 4.0 * d.Cout;
  *out = op.release();
  return PCV_OK;
}

}  // namespace PCV_TIER
}  // namespace pcv
