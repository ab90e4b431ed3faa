// Cross-layer fusion of a bottleneck's tail (ResBottleneck.forward, resnet.py:136-140, with the unit's add + ReLU,
// resnet.py:221-229):
//     y = relu( conv3_1x1( relu(conv2_3x3(x)) ) + identity )          64*k -> 64 -> Cout (256), stride 1
// in ONE kernel.  The 64-channel intermediate never exists in HBM: it goes TMEM -> registers -> shared memory (as the
// K-major SWIZZLE_128B A operand of the second GEMM) -> tensor core.  For ResNet-50 stage 1 (56x56, bs256) that removes a
// 103 MB write + 103 MB read and one launch per unit: the two kernels it replaces took 0.071 + 0.158 ms, the fused
// kernel's HBM bound (x 103 MB + identity 411 MB + y 411 MB) is 0.143 ms.
//
// MEASURED (B200, round 2, profiles/README.md): 0.224 ms - parity green, but no faster than the two kernels (0.229 ms).
// The reason is the shared-memory port, not HBM: per 2-row tile a CTA moves ~554 KB through shared memory (halo fill 30,
// GEMM1 operand reads 36 MMAs x 6 KB = 216, Y 16 + GEMM2 operands 64, identity in / result out through the staging ring
// 4 x 57) = 6.5 k clk at the ~85 B/clk the port sustains under mixed tcgen05 / LDS / STS / TMA traffic (DESIGN fact 4),
// against 5.5 k clk of HBM time: the N = 64 GEMM that was tensor/smem-bound on its own and the staged epilogue that was
// HBM-bound on its own ADD on the one port they now share.  The plan therefore records it only on request
// (plan.set_fuse_tail / PCV_FUSE_TAIL=1); the parity tests always exercise it.
//
// Structure (CTA pair, cta_group::2, 12 warps per CTA, persistent over tiles of R = 2 output rows of one image):
//   GEMM1  the halo kernel of conv_igemm3.cu: one 4-D TMA box per tile, nine taps = nine shifted descriptor views,
//          W2 resident in shared memory, D1 (256 x 64 fp32 across the pair) in TMEM, double buffered;
//   EPI1   8 warps: tcgen05.ld D1 -> +bias2 -> ReLU -> bf16 -> st.shared into the Y tile [128 rows x 128 B] per CTA
//          (double buffered);
//   GEMM2  Y (both CTAs' halves = 256 rows) x W3 (resident, each CTA half of every 128-column unit), K = 64: one
//          256 x 128 MMA group per unit into D2, double buffered (TMEM: 2 x 64 + 2 x 128 = 384 columns);
//   EPI2   tcgen05.ld D2 -> +bias3 -> +identity -> ReLU -> bf16, through a ring of 64-column staging slabs laid out as
//          image rows: the staging manager warp prefetches the identity slab with 4-D TMA row loads and writes the result
//          with 4-D TMA row stores (the ring of conv_igemm2.cu), so the epilogue warps only ever wait for data.
// Both the MMA warp and the epilogue warps are software pipelined by one tile: the tensor pipe sees G1(t+1), G2(t),
// G1(t+2), ... and the epilogue warps run EPI1(t+1), EPI2(t), EPI1(t+2), ... so neither waits for the EPI1 -> GEMM2 ->
// EPI2 round trip of a tile; the HBM streams (identity in, result out) are kept in flight by the manager warp.
#include "igemm_common.cuh"

namespace pcv {
namespace PCV_TIER {

struct FusedParams {
  const float* bias2;        // conv2 (3x3) folded-BN bias [64]
  const float* bias3;        // conv3 (1x1) folded-BN bias [Cout]
  int N, H, W, Cout;
  int R, PW;                 // output rows per tile (2), padded width W + 2
  int cblocks;               // Cin / 64
  int NU;                    // 128-column units of Cout
  int tiles_per_img, num_tiles;
  int NA, a_buf_bytes, a_tx_bytes;
  int WP8, slab_bytes, NSTG; // staged image row = WP8 pixels x 128 B; slab = R rows; ring depth
  float act_hi;              // final activation cap (ReLU: +inf, ReLU6: 6)
};

__device__ __forceinline__ void fx_tma2_load_4d(const CUtensorMap* m, uint32_t mbar_cluster_addr, void* dst, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void fx_tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

constexpr int FX_MAX_NA = 4, FX_MAX_NSTG = 8;
constexpr int FX_THREADS = 384;
constexpr int FX_D2_COL0 = 128;     // TMEM: D1 buffers at columns 0 / 64, D2 buffers at 128 / 256
constexpr int FX_TMEM_COLS = 512;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FX_THREADS, 1)
igemm3x_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmOut,
               const __grid_constant__ CUtensorMap tmRes, const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nb1 = 9 * p.cblocks;                         // resident W2 blocks of [32 x 128 B]
  constexpr int B1_BLOCK = 32 * 128, B2_UNIT = 64 * 128;
  uint8_t* sB1 = smem;
  uint8_t* sB2 = sB1 + ((nb1 * B1_BLOCK + 1023) & ~1023);   // NU units of [64 x 128 B]
  uint8_t* sY = sB2 + p.NU * B2_UNIT;                       // 2 x [128 x 128 B]: A operand of GEMM2, double buffered
  uint8_t* sA = sY + 2 * BLOCK_M * 128;
  uint8_t* sStg = sA + p.NA * p.a_buf_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + p.NSTG * p.slab_bytes);
  uint64_t* full = bars;                        // [NA]   leader's copy live
  uint64_t* empty = full + FX_MAX_NA;           // [NA]   per CTA (multicast commit)
  uint64_t* b_full = empty + FX_MAX_NA;         // [1]    leader
  uint64_t* d1_full = b_full + 1;               // [2]    per CTA (multicast commit)
  uint64_t* d1_empty = d1_full + 2;             // [2]    leader, 16 arrivals
  uint64_t* y_full = d1_empty + 2;              // [2]    leader, 16 arrivals: both CTAs' Y tiles written
  uint64_t* d2_full = y_full + 2;               // [2]    per CTA (multicast commit)
  uint64_t* d2_empty = d2_full + 2;             // [2]    leader, 16 arrivals
  uint64_t* stg_free = d2_empty + 2;            // [NSTG] (unused: every slab carries an identity load)
  uint64_t* res_full = stg_free + FX_MAX_NSTG;  // [NSTG] identity slab landed
  uint64_t* stg_full = res_full + FX_MAX_NSTG;  // [NSTG] result written (4 warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(stg_full + FX_MAX_NSTG);

  const int pw = threadIdx.x >> 5;
  const int warp = pw >= 8 ? pw - 8 : pw + 4;   // roles: 0 producer, 1 MMA, 2 TMEM, 3 staging manager, 4-11 epilogue
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int pair_tiles = (p.num_tiles + 1) >> 1;
  const int my_tiles = pair < pair_tiles ? (pair_tiles - pair + npairs - 1) / npairs : 0;
  const int SLABS = 2 * p.NU;                   // 64-column slabs per tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmB2);
    tma_prefetch_desc(&tmOut);
    tma_prefetch_desc(&tmRes);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.NA; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&y_full[i], 16);
      mbar_init(&d1_full[i], 1);
      mbar_init(&d1_empty[i], 16);
      mbar_init(&d2_full[i], 1);
      mbar_init(&d2_empty[i], 16);
    }
    for (int i = 0; i < p.NSTG; ++i) {
      mbar_init(&stg_free[i], 1);
      mbar_init(&res_full[i], 1);
      mbar_init(&stg_full[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_ptr, FX_TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (tmem_base != 0) __trap();   // one CTA per SM: the allocation starts at column 0 (the MMA warp relies on it)
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================== producer (both CTAs): resident weights, then one halo box per tile ==========
    const uint32_t bfull_leader = mapa_u32(smem_u32(b_full), 0);
    if (rank == 0 && elect_one()) mbar_arrive_expect_tx(b_full, 2 * (nb1 * B1_BLOCK + p.NU * B2_UNIT));
    for (int b = 0; b < nb1; ++b)
      if (elect_one()) tma2_load_2d(&tmB1, bfull_leader, sB1 + b * B1_BLOCK, b * BLOCK_K, static_cast<int>(rank) * 32);
    for (int u = 0; u < p.NU; ++u)
      if (elect_one()) tma2_load_2d(&tmB2, bfull_leader, sB2 + u * B2_UNIT, 0, u * 128 + static_cast<int>(rank) * 64);
    pdl_wait();   // the weights do not depend on the previous kernel; the activations do
    int slot = 0;
    uint32_t phase = 0;
    for (int t = pair; t < pair_tiles; t += npairs) {
      int tile = 2 * t + static_cast<int>(rank);
      if (tile >= p.num_tiles) tile = p.num_tiles - 1;   // phantom tile of an odd tail: recomputed, never stored
      const int img = tile / p.tiles_per_img;
      const int h0 = (tile - img * p.tiles_per_img) * p.R;
      for (int cb = 0; cb < p.cblocks; ++cb) {
        mbar_wait(&empty[slot], phase ^ 1);
        const uint32_t full_leader = mapa_u32(smem_u32(&full[slot]), 0);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&full[slot], 2 * p.a_tx_bytes);
          fx_tma2_load_4d(&tmA, full_leader, sA + slot * p.a_buf_bytes, cb * BLOCK_K, -1, h0 - 1, img);
        }
        if (++slot == p.NA) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA): G1(0), then per tile G1(t+1), G2(t) =================
    if (rank == 0) {
      constexpr uint32_t idesc1 = make_idesc_e16(2 * BLOCK_M, 64);
      constexpr uint32_t idesc2 = make_idesc_e16(2 * BLOCK_M, 128);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA)), b1_lo0 = smem_desc_lo(smem_u32(sB1));
      const uint32_t y_lo = smem_desc_lo(smem_u32(sY)), b2_lo0 = smem_desc_lo(smem_u32(sB2));
      const uint32_t row_step = p.PW * (128 >> 4);
      const uint32_t tap_step = p.cblocks * (B1_BLOCK >> 4);
      mbar_wait(b_full, 0);
      tc_fence_after();
      int slot = 0;
      uint32_t phase = 0;
      int u_cnt = 0;   // GEMM2 units issued so far (D2 buffer = u_cnt & 1)
      auto gemm1 = [&](int it) {
        const int buf = it & 1;
        mbar_wait(&d1_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = buf * 64;
        for (int cb = 0; cb < p.cblocks; ++cb) {
          mbar_wait(&full[slot], phase);
          tc_fence_after();
          const uint32_t a_buf = a_lo0 + slot * (p.a_buf_bytes >> 4);
          const uint32_t b_cb = b1_lo0 + cb * (B1_BLOCK >> 4);
          if (elect_one()) {
#pragma unroll
            for (int fr = 0; fr < 3; ++fr) {
              const uint32_t a_row = a_buf + fr * row_step;
#pragma unroll
              for (int fs = 0; fs < 3; ++fs) {
                const uint32_t a_lo = a_row + fs * (128 >> 4);
                const uint32_t b_lo = b_cb + (fr * 3 + fs) * tap_step;
#pragma unroll
                for (int k = 0; k < BLOCK_K / 16; ++k)
                  umma2_bf16_lohi(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc1, (fr | fs | k) != 0 ? 1u : (cb != 0 ? 1u : 0u));
              }
            }
          }
          if (elect_one()) umma2_commit(&empty[slot], 0x3);
          if (++slot == p.NA) {
            slot = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma2_commit(&d1_full[buf], 0x3);
      };
      if (my_tiles > 0) gemm1(0);
      for (int it = 0; it < my_tiles; ++it) {
        if (it + 1 < my_tiles) gemm1(it + 1);
        // GEMM2 of tile `it`: Y(it) is complete in both CTAs
        mbar_wait(&y_full[it & 1], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t y_it = y_lo + (it & 1) * (BLOCK_M * 128 >> 4);
        for (int u = 0; u < p.NU; ++u, ++u_cnt) {
          const int b2 = u_cnt & 1;
          mbar_wait(&d2_empty[b2], ((u_cnt >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = FX_D2_COL0 + b2 * 128;
          const uint32_t b_lo = b2_lo0 + u * (B2_UNIT >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) umma2_bf16_lohi(d_tmem, y_it + 2 * k, b_lo + 2 * k, idesc2, k ? 1u : 0u);
            umma2_commit(&d2_full[b2], 0x3);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 3) {
    // ===================================== staging manager: identity slabs in, result slabs out (image rows) ============
    pdl_wait();
    const int total_rounds = my_tiles * SLABS;
    auto round_coords = [&](int r, int& col0, int& h0, int& img, bool& ok) {
      const int it = r / SLABS;
      const int tile = 2 * (pair + it * npairs) + static_cast<int>(rank);
      ok = tile < p.num_tiles;
      const int tl = ok ? tile : p.num_tiles - 1;
      img = tl / p.tiles_per_img;
      h0 = (tl - img * p.tiles_per_img) * p.R;
      col0 = (r - it * SLABS) * 64;
    };
    auto refill = [&](int r) {
      if (r >= total_rounds) return;
      const int s = r % p.NSTG;
      int col0, h0, img;
      bool ok;
      round_coords(r, col0, h0, img, ok);
      if (elect_one()) {
        mbar_arrive_expect_tx(&res_full[s], p.R * p.W * 128);
        for (int rr = 0; rr < p.R; ++rr)
          fx_tma_load_4d(&tmRes, &res_full[s], sStg + s * p.slab_bytes + rr * p.WP8 * 128, col0, 0, h0 + rr, img);
      }
    };
    for (int r = 0; r < p.NSTG; ++r) refill(r);
    for (int r = 0; r < total_rounds; ++r) {
      const int s = r % p.NSTG;
      mbar_wait(&stg_full[s], (r / p.NSTG) & 1);
      if (elect_one()) {
        int col0, h0, img;
        bool ok;
        round_coords(r, col0, h0, img, ok);
        if (ok)
          for (int rr = 0; rr < p.R; ++rr)
            tma_store_4d(&tmOut, sStg + s * p.slab_bytes + rr * p.WP8 * 128, col0, 0, h0 + rr, img);
        tma_store_commit();
        tma_store_wait_read<1>();   // the stores of round r-1 have left their slab
      }
      __syncwarp();
      if (r >= 1) refill(r - 1 + p.NSTG);
    }
    if (elect_one()) tma_store_wait_read<0>();
    __syncwarp();
    if (elect_one()) tma_store_wait_all<0>();
  } else if (warp >= 4) {
    // ===================================== epilogue warps: EPI1(t) then EPI2(t) ====================================
    const int q4 = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q4 * 32 + lane;              // accumulator row = tile pixel q = r * PW + c
    const int r_img = row / p.PW, c_img = row - r_img * p.PW;
    const bool px_ok = c_img < p.W && r_img < p.R;
    const uint32_t srow = r_img * p.WP8 + c_img; // staged pixel index
    const uint32_t sw_stg = (srow & 7u) << 4, sw_y = (row & 7u) << 4;
    const uint32_t y_u32 = smem_u32(sY) + row * 128;
    const uint32_t stg_u32 = smem_u32(sStg);
    const bool capped = p.act_hi != INFINITY;
    const uint32_t cap2 = pack_e16x2(p.act_hi, p.act_hi);
    const uint32_t yf0 = mapa_u32(smem_u32(&y_full[0]), 0), yf1 = mapa_u32(smem_u32(&y_full[1]), 0);
    const uint32_t d1e0 = mapa_u32(smem_u32(&d1_empty[0]), 0), d1e1 = mapa_u32(smem_u32(&d1_empty[1]), 0);
    const uint32_t d2e0 = mapa_u32(smem_u32(&d2_empty[0]), 0), d2e1 = mapa_u32(smem_u32(&d2_empty[1]), 0);
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    // ---- EPI1(it): D1 -> relu(+bias2) -> Y[it & 1] (bf16, K-major SWIZZLE_128B rows of 64 channels) ----
    auto epi1 = [&](int it) {
      const int buf = it & 1;
      mbar_wait(&d1_full[buf], (it >> 1) & 1);
      tc_fence_after();
      uint32_t acc[32];
      tmem_ld_32x32(tmem_base + lane_base + buf * 64 + half * 32, acc);
      const float4* bias4 = reinterpret_cast<const float4*>(p.bias2 + half * 32);
      float4 b4[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) b4[i] = __ldg(bias4 + i);
      tmem_ld_wait_regs(acc);
      float v[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[4 * i + 0] = __uint_as_float(acc[4 * i + 0]) + b4[i].x;
        v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + b4[i].y;
        v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + b4[i].z;
        v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + b4[i].w;
      }
      uint32_t o[16];
      clamp_pack32(v, o, true, false, 0u);
      const uint32_t yb = y_u32 + buf * (BLOCK_M * 128);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        sts128(yb + ((static_cast<uint32_t>((half * 4 + i) << 4)) ^ sw_y), o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
      tc_fence_before();
      fence_proxy_async_smem();   // Y is read by tcgen05.mma (async proxy)
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) {
          mbar_arrive(&d1_empty[buf]);
          mbar_arrive(&y_full[buf]);
        } else {
          mbar_arrive_cluster(buf ? d1e1 : d1e0);
          mbar_arrive_cluster(buf ? yf1 : yf0);
        }
      }
    };
    int u_cnt = 0, round = 0;
    if (my_tiles > 0) epi1(0);
    for (int it = 0; it < my_tiles; ++it) {
      // software pipeline: Y(it+1) is produced BEFORE the second epilogue of tile `it`, so GEMM2(it+1) runs under EPI2(it)
      // and no epilogue warp ever waits for the EPI1 -> GEMM2 -> EPI2 round trip (Y is double buffered for this)
      if (it + 1 < my_tiles) epi1(it + 1);
      // ---- EPI2: per 128-column unit, this warp's 64-column slab (half) ----
      for (int u = 0; u < p.NU; ++u, ++u_cnt) {
        const int b2 = u_cnt & 1;
        const int my_round = round + 2 * u + half;
        const int s = my_round % p.NSTG;
        const uint32_t slab_u32 = stg_u32 + s * p.slab_bytes + srow * 128;
        mbar_wait(&d2_full[b2], (u_cnt >> 1) & 1);
        tc_fence_after();
        mbar_wait(&res_full[s], (my_round / p.NSTG) & 1);
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          uint32_t acc[32];
          tmem_ld_32x32(tmem_base + lane_base + FX_D2_COL0 + b2 * 128 + half * 64 + j * 32, acc);
          const float4* bias4 = reinterpret_cast<const float4*>(p.bias3 + u * 128 + half * 64 + j * 32);
          float4 b4[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) b4[i] = __ldg(bias4 + i);
          uint4 r4[4];
          if (px_ok) {
#pragma unroll
            for (int c = 0; c < 4; ++c) r4[c] = lds128(slab_u32 + ((static_cast<uint32_t>((j * 4 + c) << 4)) ^ sw_stg));
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) r4[c] = make_uint4(0u, 0u, 0u, 0u);
          }
          tmem_ld_wait_regs(acc);
          float v[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[4 * i + 0] = __uint_as_float(acc[4 * i + 0]) + b4[i].x;
            v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + b4[i].y;
            v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + b4[i].z;
            v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + b4[i].w;
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            v[8 * c + 0] += e16lo(r4[c].x);
            v[8 * c + 1] += e16hi(r4[c].x);
            v[8 * c + 2] += e16lo(r4[c].y);
            v[8 * c + 3] += e16hi(r4[c].y);
            v[8 * c + 4] += e16lo(r4[c].z);
            v[8 * c + 5] += e16hi(r4[c].z);
            v[8 * c + 6] += e16lo(r4[c].w);
            v[8 * c + 7] += e16hi(r4[c].w);
          }
          uint32_t o[16];
          clamp_pack32(v, o, true, capped, cap2);
          if (px_ok) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              sts128(slab_u32 + ((static_cast<uint32_t>((j * 4 + c) << 4)) ^ sw_stg), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
          }
        }
        // D2 columns of this warp are in registers / smem: release the buffer, hand the slab to the manager
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(&d2_empty[b2]);
          else mbar_arrive_cluster(b2 ? d2e1 : d2e0);
          mbar_arrive(&stg_full[s]);
        }
      }
      round += SLABS;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, FX_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct FusedOp : Op {
  CUtensorMap tmA, tmB1, tmB2, tmOut, tmRes;
  FusedParams p;
  int grid, smem_bytes;
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    static std::atomic<uint64_t> attr_done{0};
    if (cudaError_t e = set_max_smem_once(igemm3x_kernel, 232448, attr_done)) return e;
    return launch_pdl(igemm3x_kernel, dim3(grid), dim3(FX_THREADS), smem_bytes, s, tmA, tmB1, tmB2, tmOut, tmRes, p);
  }
};

struct FusedGeom {
  int R, PW, WP8, cblocks, NU, NA, NSTG, a_buf, slab, b1_bytes, smem;
};

static bool fused_geom(const pcv_conv_desc& d2, const pcv_conv_desc& d3, FusedGeom* g) {
  // conv2: dense 3x3 stride 1 pad 1, Cin % 64 == 0 -> 64 channels, ReLU;  conv3: dense 1x1 stride 1, 64 -> Cout (% 128 == 0)
  if (d2.kh != 3 || d2.kw != 3 || d2.stride != 1 || d2.pad != 1 || d2.dil != 1 || d2.groups != 1 || d2.Cout != 64 ||
      d2.Cin % 64 != 0 || d2.Cin > 256 || d2.act != PCV_ACT_RELU || d2.flags != 0 || d2.in_row_pitch != 0)
    return false;
  if (d3.kh != 1 || d3.kw != 1 || d3.stride != 1 || d3.pad != 0 || d3.groups != 1 || d3.Cin != 64 || d3.Cout % 128 != 0 ||
      d3.Cout > 1024 || (d3.act != PCV_ACT_RELU && d3.act != PCV_ACT_RELU6) || d3.flags != 0 || d3.in_row_pitch != 0)
    return false;
  if (d2.N != d3.N || d2.H != d3.H || d2.W != d3.W) return false;
  if (pitch_or(d2.in_pitch, d2.Cin) % 8 || pitch_or(d3.out_pitch, d3.Cout) % 8 || pitch_or(d3.res_pitch, d3.Cout) % 8) return false;
  g->R = 2;
  g->PW = d2.W + 2;
  if (g->R * g->PW > BLOCK_M || d2.H % g->R != 0 || d2.W < 16) return false;   // one M-block per tile, whole tiles only
  g->WP8 = round_up(d2.W, 8);
  g->cblocks = d2.Cin / 64;
  g->NU = d3.Cout / 128;
  g->b1_bytes = round_up(9 * g->cblocks * 32 * 128, 1024);
  g->a_buf = round_up((BLOCK_M + 2 * g->PW + 2) * 128, 1024);
  g->slab = g->R * g->WP8 * 128;                        // multiple of 1024 (WP8 % 8 == 0)
  const int fixed = 1024 + g->b1_bytes + g->NU * 64 * 128 + 2 * BLOCK_M * 128 + 512;
  g->NSTG = 4;
  g->NA = std::min(FX_MAX_NA, (232448 - fixed - g->NSTG * g->slab) / g->a_buf);
  if (g->NA < 2) return false;
  g->NSTG = std::min(FX_MAX_NSTG, (232448 - fixed - g->NA * g->a_buf) / g->slab);
  g->smem = fixed + g->NA * g->a_buf + g->NSTG * g->slab;
  return true;
}

int fused_tail_ok(const pcv_conv_desc& d2, const pcv_conv_desc& d3) {
  FusedGeom g;
  return fused_geom(d2, d3, &g) ? 1 : 0;
}

int fused_tail_make(const pcv_conv_desc& d2, const pcv_conv_desc& d3, const void* x, const void* w2, const float* bias2,
                    const void* w3, const float* bias3, const void* res, void* y, Op** out) {
  FusedGeom g;
  if (!fused_geom(d2, d3, &g)) return fail(PCV_ERR_UNSUPPORTED, "bottleneck tail outside the fused kernel's domain");
  PCV_REQUIRE(x && w2 && w3 && bias2 && bias3 && res && y, "NULL tensor pointer");
  for (const void* ptr : {x, w2, w3, res, static_cast<const void*>(y)})
    PCV_REQUIRE(reinterpret_cast<uintptr_t>(ptr) % 16 == 0, "fused tail operands must be 16-byte aligned");
  const int in_pitch = pitch_or(d2.in_pitch, d2.Cin), out_pitch = pitch_or(d3.out_pitch, d3.Cout);
  const int res_pitch = pitch_or(d3.res_pitch, d3.Cout);
  auto op = std::make_unique<FusedOp>();
  FusedParams& p = op->p;
  p.bias2 = bias2; p.bias3 = bias3;
  p.N = d2.N; p.H = d2.H; p.W = d2.W; p.Cout = d3.Cout;
  p.R = g.R; p.PW = g.PW; p.cblocks = g.cblocks; p.NU = g.NU;
  p.tiles_per_img = d2.H / g.R;
  p.num_tiles = d2.N * p.tiles_per_img;
  p.NA = g.NA; p.a_buf_bytes = g.a_buf; p.a_tx_bytes = (g.R + 2) * g.PW * 128;
  p.WP8 = g.WP8; p.slab_bytes = g.slab; p.NSTG = g.NSTG;
  p.act_hi = d3.act == PCV_ACT_RELU6 ? 6.f : INFINITY;
  op->smem_bytes = g.smem;
  op->grid = 2 * std::min((p.num_tiles + 1) / 2, sm_count() / 2);
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  auto enc4 = [&](CUtensorMap* tm, const void* base, int C, int pitch, int boxw, int boxh, const char* what) -> int {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)d2.W, (cuuint64_t)d2.H, (cuuint64_t)d2.N};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)d2.W * pitch * 2, (cuuint64_t)d2.H * d2.W * pitch * 2};
    cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)boxw, (cuuint32_t)boxh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(tm, TMAP_E16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled (fused %s) failed (%d)", what, (int)r);
    return PCV_OK;
  };
  if (int rc = enc4(&op->tmA, x, d2.Cin, in_pitch, g.PW, g.R + 2, "halo A")) return rc;
  if (int rc = enc4(&op->tmOut, y, d3.Cout, out_pitch, d2.W, 1, "out")) return rc;
  if (int rc = enc4(&op->tmRes, res, d3.Cout, res_pitch, d2.W, 1, "identity")) return rc;
  const uint64_t k1 = 9ull * g.cblocks * BLOCK_K;
  if (int rc = make_tiled_2d(&op->tmB1, w2, k1, 64, k1 * 2, BLOCK_K, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = make_tiled_2d(&op->tmB2, w3, BLOCK_K, d3.Cout, BLOCK_K * 2, BLOCK_K, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  char nm[176];
  snprintf(nm, sizeof nm, "conv_tc3x fused 3x3 %d->64 + 1x1 64->%d +res @%dx%d R=%d na=%d nstg=%d", d2.Cin, d3.Cout, d2.H, d2.W,
           g.R, g.NA, g.NSTG);
  op->name = nm;
  const double M = static_cast<double>(d2.N) * d2.H * d2.W;
  op->flops = 2.0 * M * 64 * d2.Cin * 9 + 2.0 * M * d3.Cout * 64;
  // fused group ("the bound is recomputed with the fused group's bytes"): x + identity + y + both weights
  op->bytes = 2.0 * M * d2.Cin + 2.0 * 2.0 * M * d3.Cout + 2.0 * (64.0 * d2.Cin * 9 + 64.0 * d3.Cout) + 4.0 * (64 + d3.Cout);
  *out = op.release();
  return PCV_OK;
}

}  // namespace PCV_TIER
}  // namespace pcv
