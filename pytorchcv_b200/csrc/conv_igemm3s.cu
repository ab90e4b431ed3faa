// Space-to-depth STEM convolution (k x k stride 2 on a <= 4-channel image) with a shared-memory halo tile of 32-byte
// pixel rows.
//
// Why: the s2d stem (include/pcv_b200.h) is a T x T stride-1 convolution (T = k/2 + 1) over 16-channel s2d pixels.
// Served by the im2col kernel through an overlapping-window view it re-reads every s2d pixel T*T/ (T) = T times
// through the L2 -> SMEM fabric as 128-byte rows (ResNet 7x7: 1.64 GB of fill for a 0.1 GB tensor; the stem was the
// single most expensive launch of the ResNet-50 step at 0.22 ms against an 0.08 ms HBM bound).  Here the tile's input
// box - (R+T-1) rows x PW pixels x 16 channels, ONE 4-D tiled TMA load, SWIZZLE_32B - lands in shared memory as a
// dense list of 32-byte pixel rows = the canonical K-major SWIZZLE_32B operand with "row" = s2d pixel index, and the
// A operand of tap (fr, fs) is the same buffer with the descriptor start advanced by (fr*PW + fs) * 32 bytes (the
// swizzle is a function of the absolute shared-memory address, as for the 128-byte rows of conv_igemm3.cu).  One
// K = 16 MMA per tap: 16 MMAs (7x7) or 4 (3x3) per 256 output pixels; fill traffic = (R+T-1)/R x the s2d tensor.
//
// Roles / pairing / TMEM double buffering are those of conv_igemm3.cu: CTA pair (cta_group::2, 256 x BN MMA), each CTA
// its own tile and half of the output channels' weights (resident), 8 epilogue warps.
//
// Epilogue stores: a thread owns one output pixel (TMEM lane), so direct stores put 32 different 128-byte lines behind
// every STG (ncu: LSU wavefronts 60 % of peak, half-written sectors, the store queue back-pressuring the epilogue;
// 0.246 ms).  Instead the tile is staged in shared memory as R image rows of round_up(Wo, 8) swizzled pixel rows and
// leaves through one 4-D TMA store per image row (garbage columns c >= Wo are never written), double-buffered so the
// stores of tile i drain while tile i+1 is computed.
//
// POOL = true fuses the 3x3 stride-2 pad-1 max pool that follows the stem in every ResNet-style init block
// (resnet.py:255-258): the conv tile (R even rows) is only STAGED, the R/2 pooled rows are reduced from the staged rows
// plus the last row of the previous tile (still intact in the other staging buffer), and only the pooled rows are
// written (4x fewer output bytes; the 64 x 112 x 112 stem output - 0.4 GB per 256 images, written once and read once -
// never exists).  For the carry row to be there, every CTA walks a CONTIGUOUS range of tiles, preceded by one warm-up
// tile (computed and staged, nothing stored).
#include "igemm_common.cuh"

namespace pcv {
namespace PCV_TIER {

struct StemParams {
  const float* bias;
  e16* out;
  int out_pitch;
  int N, Ho, Wo, Cout;
  int R, PW, NMB, T;         // output rows per tile, s2d row width (Wo + T - 1), 128-row M-blocks per tile, taps per axis
  int tiles_per_img, num_tiles;
  int NA;                    // A ring depth
  int a_buf_bytes;           // smem stride of one A buffer (multiple of 1024, includes the over-read slack)
  int a_tx_bytes;            // bytes of one A box: (R+T-1) * PW * 32
  int WP8, stg_bytes;        // staging: pixels per staged image row (Wo rounded up to 8), bytes of one staging buffer
  int tpc;                   // POOL: tiles per CTA (contiguous range), excluding the warm-up tile
  int Hp, Wp, WPP8;          // POOL: pooled map size, pooled pixels per staged pooled row (Wp rounded up to 8)
  float act_lo, act_hi;
  int dbg;                   // PCV_STEM_DBG throughput experiments: 1 skip MMA issue, 2 skip epilogue loads+math+staging, 4 skip A loads, 8 skip the TMA stores, 16 skip the pool pass
};

__device__ __forceinline__ void tma2_load_4d_s(const CUtensorMap* m, uint32_t mbar_cluster_addr, void* dst, int c0,
                                               int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

constexpr int ST_MAX_NA = 6;
constexpr int ST_THREADS = 384;   // 4 control warps + 8 epilogue warps
constexpr int ST_ROW = 32;        // bytes per s2d pixel (16 bf16 channels)

template <int BN, int T, bool POOL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ST_THREADS, 1)
stem_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int B_BLOCK = BN / 2 * ST_ROW;                    // one tap's weights for this CTA's half of the channels
  constexpr int ntaps = T * T;
  uint8_t* sB = smem;                                         // resident weights: ntaps x [BN/2 x 32 B]
  uint8_t* sA = sB + ((ntaps * B_BLOCK + 1023) & ~1023);
  uint8_t* sStg = sA + p.NA * p.a_buf_bytes;                  // 2 output staging buffers (POOL: 1 buffer of R/2 vertically pooled rows)
  uint8_t* sPool = sStg + (POOL ? 1 : 2) * p.stg_bytes;       // POOL: 2 staging buffers of R/2 pooled rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPool + (POOL ? 2 * (p.R / 2) * p.WPP8 * BN * 2 : 0));
  uint64_t* full = bars;                         // [NA]  leader's copy is live
  uint64_t* empty = bars + ST_MAX_NA;            // [NA]  per CTA, multicast commit
  uint64_t* b_full = bars + 2 * ST_MAX_NA;       // [1]   leader's copy
  uint64_t* tmem_full = b_full + 1;              // [2]   per CTA, multicast commit
  uint64_t* tmem_empty = tmem_full + 2;          // [2]   leader's copy, 16 arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int pw = threadIdx.x >> 5;
  const int warp = pw >= 8 ? pw - 8 : pw + 4;    // control roles 0-3 on physical warps 8-11, epilogue roles 4-11
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int pair_tiles = (p.num_tiles + 1) >> 1;
  // Tile schedule, identical in every role.  Default: pair-interleaved (iteration i of this CTA = tile
  // 2*(pair + i*npairs) + rank).  POOL: CTA c owns tiles [c*tpc, (c+1)*tpc) and runs tpc + 1 iterations, the first one
  // being the warm-up tile c*tpc - 1; both CTAs of a pair run the same number of iterations (one MMA serves both).
  const int n_iter = POOL ? p.tpc + 1 : (pair < pair_tiles ? (pair_tiles - pair + npairs - 1) / npairs : 0);
  auto sched = [&](int it, int& tile, bool& live) {
    if (POOL) {
      tile = (2 * pair + static_cast<int>(rank)) * p.tpc - 1 + it;
      live = it >= 1 && tile < p.num_tiles;
      tile = max(0, min(tile, p.num_tiles - 1));
    } else {
      tile = 2 * (pair + it * npairs) + static_cast<int>(rank);
      live = tile < p.num_tiles;
      if (!live) tile = p.num_tiles - 1;   // phantom tile of an odd tail: recompute the last one, stores masked
    }
  };
  const int acc_cols = p.NMB * BN;               // TMEM columns of one accumulator buffer
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * acc_cols) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
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
  if (tmem_base != 0) __trap();   // one CTA per SM => the allocation starts at column 0; the MMA warp relies on it
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================== producer (both CTAs) =====================================
    const uint32_t bfull_leader = mapa_u32(smem_u32(b_full), 0);
    if (rank == 0 && elect_one()) mbar_arrive_expect_tx(b_full, 2 * ntaps * B_BLOCK);
    for (int fr = 0; fr < T; ++fr)
      for (int fs = 0; fs < T; ++fs)
        if (elect_one())
          tma2_load_2d(&tmB, bfull_leader, sB + (fr * T + fs) * B_BLOCK, fr * BLOCK_K + fs * 16,
                       static_cast<int>(rank) * (BN / 2));
    pdl_wait();   // the weights do not depend on the previous kernel; the s2d image does
    int slot = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_iter; ++it) {
      int tile;
      bool live;
      sched(it, tile, live);
      const int img = tile / p.tiles_per_img;
      const int h0 = (tile - img * p.tiles_per_img) * p.R;
      mbar_wait(&empty[slot], phase ^ 1);
      const uint32_t full_leader = mapa_u32(smem_u32(&full[slot]), 0);
      if (elect_one()) {
        if (p.dbg & 4) {
          if (rank == 0) mbar_arrive(&full[slot]);
        } else {
          if (rank == 0) mbar_arrive_expect_tx(&full[slot], 2 * p.a_tx_bytes);
          tma2_load_4d_s(&tmA, full_leader, sA + slot * p.a_buf_bytes, 0, 0, h0, img);
        }
      }
      if (++slot == p.NA) {
        slot = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =====================================
    if (rank == 0) {   // whole warp, warp-uniform values; tcgen05 instructions under elect.sync
      constexpr uint32_t idesc = make_idesc_e16(2 * BLOCK_M, BN);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA)), b_lo0 = smem_desc_lo(smem_u32(sB));
      mbar_wait(b_full, 0);
      tc_fence_after();
      int slot = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[buf], acc_phase ^ 1);
        tc_fence_after();
        mbar_wait(&full[slot], phase);
        tc_fence_after();
        const uint32_t d_tmem = buf * acc_cols;
        const uint32_t a_buf = a_lo0 + slot * (p.a_buf_bytes >> 4);
        // all T*T taps of one M-block inside ONE elect.sync region with compile-time tap offsets: the descriptors stay in
        // uniform registers and the UTCHMMAs issue back to back (one elect region per MMA cost ~100 cycles per MMA,
        // 3x the N = 64 MMA itself: the first version of this kernel was issue-bound at 0.17 ms)
        const uint32_t row_step = p.PW * (ST_ROW >> 4);
        for (int mb = 0; mb < ((p.dbg & 1) ? 0 : p.NMB); ++mb) {
          const uint32_t a_mb = a_buf + mb * (BLOCK_M * ST_ROW >> 4);
          const uint32_t d_mb = d_tmem + mb * BN;
          if (elect_one()) {
#pragma unroll
            for (int fr = 0; fr < T; ++fr) {
              const uint32_t a_row = a_mb + fr * row_step;
#pragma unroll
              for (int fs = 0; fs < T; ++fs)
                umma2_bf16_lohi_h(d_mb, a_row + fs * (ST_ROW >> 4), b_lo0 + (fr * T + fs) * (B_BLOCK >> 4), idesc,
                                  (fr | fs) != 0 ? 1u : 0u, SMEM_DESC_HI_SW32);
            }
          }
        }
        if (elect_one()) {
          umma2_commit(&empty[slot], 0x3);
          umma2_commit(&tmem_full[buf], 0x3);
        }
        if (++slot == p.NA) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
  } else if (POOL && warp >= 4) {
    // ============================ epilogue with the max pool fused (both CTAs) ============================
    // The pooled kernel runs with PW = 128 = BLOCK_M (the TMA box is 128 pixels wide whatever the s2d row width; the
    // extra columns are out-of-bounds zero fill), so M-block mb IS conv row mb of the tile and TMEM lane c IS conv
    // column c: a thread sees the same column of all R rows.  The vertical half of the separable 3x3 / stride-2 max
    // therefore happens in REGISTERS on the packed bf16 values (max commutes with the monotonic ReLU + rounding),
    // the row above the tile being carried in registers from the previous tile (contiguous tile schedule; warm-up
    // tile), and only the R/2 vertically pooled rows are staged.  Pass 2 takes the horizontal 3-max at even columns
    // from that buffer and hands the pooled rows to the TMA.  Shared-memory traffic of the pooling: 71 KB per tile
    // instead of 164 KB - this kernel is bound by the shared-memory port (MMA operand reads included).
    // Padding: row -1 of the image is skipped, column -1 is replaced by column 0 (max is idempotent).
    pdl_wait();   // the output buffer may alias a tensor the previous kernel is still reading
    const int q4 = warp & 3;
    const int j = (warp - 4) >> 2;                 // this warp's 32-channel chunk, the same for every row
    const int col = q4 * 32 + lane;                // conv column of this thread
    const bool col_ok = col < p.Wo;
    const float act_lo = p.act_lo, act_hi = p.act_hi;
    const bool relu_only = act_lo == 0.f && act_hi == INFINITY;
    const uint32_t tmem_empty_leader0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    const uint32_t tmem_empty_leader1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
    constexpr int ROWB = BN * 2;                   // 128 bytes per staged pixel (BN = 64)
    float bias_r[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + j * 32) + i);
      bias_r[4 * i + 0] = b.x; bias_r[4 * i + 1] = b.y; bias_r[4 * i + 2] = b.z; bias_r[4 * i + 3] = b.w;
    }
    const uint32_t t_off0 = (static_cast<uint32_t>(q4 * 32) << 16) + j * 32;
    const bool storer = warp == 4 && lane == 0;    // owns every bulk store group of this CTA
    const uint32_t sV_u32 = smem_u32(sStg);        // [R/2][WP8] vertically pooled pixels
    const uint32_t sPool_u32 = smem_u32(sPool);    // 2 x [R/2][WPP8] pooled pixels
    const int tid = (warp - 4) * 32 + lane;        // 0..255 among the epilogue threads
    const uint32_t v_rowb = p.WP8 * ROWB, p_rowb = p.WPP8 * ROWB;
    const uint32_t pool_buf_bytes = (p.R / 2) * p_rowb;
    // this thread's staged pixel: column `col`, chunks j*4 .. j*4+3, SWIZZLE_128B phase = col & 7 (WP8 % 8 == 0)
    const uint32_t v_dst = sV_u32 + col * ROWB;
    const uint32_t v_sw = col & 7u;
    uint32_t prev[16];                             // packed row above the next row (carried across tiles)
#pragma unroll
    for (int i = 0; i < 16; ++i) prev[i] = 0u;
    for (int it = 0; it < n_iter; ++it) {
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int tile;
      bool live;
      sched(it, tile, live);
      const int img = tile / p.tiles_per_img;
      const int h0 = (tile - img * p.tiles_per_img) * p.R;

      mbar_wait(&tmem_full[buf], acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + buf * acc_cols + t_off0;
      uint32_t acc[2][32];
      uint32_t v[16];
      if (!(p.dbg & 2)) {
        tmem_ld_32x32(tbase, acc[0]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {              // R <= 4 rows = M-blocks
          if (r < p.R) {
            tmem_ld_wait_regs(acc[r & 1]);
            if (r + 1 < 4 && r + 1 < p.R) tmem_ld_32x32(tbase + (r + 1) * BN, acc[(r + 1) & 1]);
            const uint32_t* a = acc[r & 1];
            uint32_t o[16];
            if (relu_only) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                o[i] = pack_relu_e16x2(__uint_as_float(a[2 * i]) + bias_r[2 * i], __uint_as_float(a[2 * i + 1]) + bias_r[2 * i + 1]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                o[i] = pack_e16x2(fminf(fmaxf(__uint_as_float(a[2 * i]) + bias_r[2 * i], act_lo), act_hi),
                                   fminf(fmaxf(__uint_as_float(a[2 * i + 1]) + bias_r[2 * i + 1], act_lo), act_hi));
            }
            if ((r & 1) == 0) {                    // row 2pr: start pooled row pr with the row above (if inside the image)
              const bool top = r == 0 && h0 == 0;
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = top ? o[i] : hmax2_e16(prev[i], o[i]);
            } else {                               // row 2pr+1 completes it
              if (col_ok) {
                const uint32_t dstp = v_dst + (r >> 1) * v_rowb;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  sts128(dstp + (((j * 4 + i) ^ v_sw) << 4), hmax2_e16(v[4 * i], o[4 * i]), hmax2_e16(v[4 * i + 1], o[4 * i + 1]),
                         hmax2_e16(v[4 * i + 2], o[4 * i + 2]), hmax2_e16(v[4 * i + 3], o[4 * i + 3]));
              }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) prev[i] = o[i];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tmem_empty[buf]);
        else mbar_arrive_cluster(buf ? tmem_empty_leader1 : tmem_empty_leader0);
      }
      if (storer) tma_store_wait_read<1>();         // the pooled rows of tile it-2 have left this sPool buffer
      named_bar_sync(1, 256);                       // vertically pooled rows staged by all 8 warps; sPool[buf] free
      const uint32_t pool_u32 = sPool_u32 + buf * pool_buf_bytes;
      if (live && !(p.dbg & 16)) {
        constexpr int CG = BN / 8;                  // 16-byte channel groups per pixel
        const int per_row = p.Wp * CG;
        for (int idx = tid; idx < (p.R / 2) * per_row; idx += 256) {
          const int pr = idx >= per_row ? 1 : 0;    // R / 2 <= 2 pooled rows
          const int rem = idx - pr * per_row;
          const uint32_t cg = rem & (CG - 1);
          const uint32_t px = rem >> 3;
          const uint32_t c1 = 2 * px, c0 = c1 == 0 ? 0 : c1 - 1, c2 = c1 + 1;   // c2 <= Wo - 1: Wo is even
          const uint32_t rowp = sV_u32 + pr * v_rowb;
          const uint4 a = lds128(rowp + c0 * ROWB + ((cg ^ (c0 & 7u)) << 4));
          const uint4 b = lds128(rowp + c1 * ROWB + ((cg ^ (c1 & 7u)) << 4));
          const uint4 c = lds128(rowp + c2 * ROWB + ((cg ^ (c2 & 7u)) << 4));
          sts128(pool_u32 + pr * p_rowb + px * ROWB + ((cg ^ (px & 7u)) << 4), hmax3_e16(a.x, b.x, c.x),
                 hmax3_e16(a.y, b.y, c.y), hmax3_e16(a.z, b.z, c.z), hmax3_e16(a.w, b.w, c.w));
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(2, 256);                       // pooled rows staged; the vertical buffer may be overwritten
      if (storer && live && !(p.dbg & 8)) {
        for (int pr = 0; pr < p.R / 2; ++pr)
          tma_store_4d(&tmOut, sPool + buf * pool_buf_bytes + pr * p_rowb, 0, 0, (h0 >> 1) + pr, img);
        tma_store_commit();
      }
    }
    if (storer) tma_store_wait_all<0>();
  } else if (!POOL && warp >= 4) {
    // ===================================== epilogue (both CTAs) =====================================
    // eight warps: warp w owns TMEM lane quarter (w & 3) and every second (M-block, 32-column chunk) work item
    // w = half + 2k.  With CH = BN/32 chunks per M-block that is always the SAME chunk j (CH = 2: j = half, mb = k;
    // CH = 1: j = 0, mb = half + 2k), so its 32 bias values live in registers for the whole kernel, and a thread's
    // pixel (TMEM lane) -> staged-row mapping is tile-invariant and precomputed per k (no divisions in the tile loop).
    pdl_wait();   // the output buffer may alias a tensor the previous kernel is still reading
    const int q4 = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q4 * 32 + lane;
    const float act_lo = p.act_lo, act_hi = p.act_hi;
    const bool relu_only = act_lo == 0.f && act_hi == INFINITY;
    const uint32_t tmem_empty_leader0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    const uint32_t tmem_empty_leader1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
    constexpr int CH = BN / 32;     // 32-column chunks per M-block
    constexpr int ROWB = BN * 2;    // bytes of one staged pixel
    constexpr int MAXK = 4;         // work items per warp and tile: ceil(NMB * CH / 2) <= 4 (2 * NMB * BN <= 512 columns)
    const int j = half % CH;
    const int nk = (p.NMB * CH - half + 1) >> 1;
    float bias_r[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + j * 32) + i);
      bias_r[4 * i + 0] = b.x; bias_r[4 * i + 1] = b.y; bias_r[4 * i + 2] = b.z; bias_r[4 * i + 3] = b.w;
    }
    int st_off[MAXK], st_r[MAXK];   // byte offset of this thread's staged pixel (-1: padding column / beyond R), its local row
    const uint32_t t_off0 = (static_cast<uint32_t>(q4 * 32) << 16) + (CH == 2 ? 0 : half) * BN + j * 32;
    constexpr uint32_t T_STEP = CH == 2 ? BN : 2 * BN;   // TMEM columns between this warp's consecutive work items
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
      const int mb = CH == 2 ? k : half + 2 * k;
      const int q = mb * BLOCK_M + row;
      const int r = q / p.PW;
      const int c = q - r * p.PW;
      const uint32_t srow = r * p.WP8 + c;   // staged pixel index; staged image rows start on swizzle-atom boundaries
      st_off[k] = (k < nk && c < p.Wo && r < p.R) ? static_cast<int>(srow * ROWB) : -1;
      st_r[k] = r;
    }
    const bool storer = warp == 4 && lane == 0;   // owns every bulk store group of this CTA
    const uint32_t sStg_u32 = smem_u32(sStg);
    for (int it = 0; it < n_iter; ++it) {
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int tile;
      bool live;
      sched(it, tile, live);
      const bool tile_ok = live;
      const int img = tile / p.tiles_per_img;
      const int h0 = (tile - img * p.tiles_per_img) * p.R;
      uint8_t* stg = sStg + buf * p.stg_bytes;   // free: the storer waited for tile it-2's stores before barrier 1 of tile it-1
      const uint32_t stg_u32 = sStg_u32 + buf * p.stg_bytes;

      mbar_wait(&tmem_full[buf], acc_phase);
      tc_fence_after();

      // TMEM -> registers -> (+bias, activation, bf16) -> swizzled staging; the tcgen05.ld of item k+1 is in flight while
      // item k is converted
      const uint32_t tbase = tmem_base + buf * acc_cols;
      uint32_t acc[2][32];
      if (nk > 0 && !(p.dbg & 2)) tmem_ld_32x32(tbase + t_off0, acc[0]);
#pragma unroll
      for (int k = 0; k < MAXK; ++k) {
        if (k < nk && !(p.dbg & 2)) {
          tmem_ld_wait_regs(acc[k & 1]);
          if (k + 1 < MAXK && k + 1 < nk) tmem_ld_32x32(tbase + t_off0 + (k + 1) * T_STEP, acc[(k + 1) & 1]);
          const uint32_t* a = acc[k & 1];
          uint32_t o[16];
          if (relu_only) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              o[i] = pack_relu_e16x2(__uint_as_float(a[2 * i]) + bias_r[2 * i], __uint_as_float(a[2 * i + 1]) + bias_r[2 * i + 1]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              o[i] = pack_e16x2(fminf(fmaxf(__uint_as_float(a[2 * i]) + bias_r[2 * i], act_lo), act_hi),
                                 fminf(fmaxf(__uint_as_float(a[2 * i + 1]) + bias_r[2 * i + 1], act_lo), act_hi));
          }
          if (st_off[k] >= 0 && tile_ok && h0 + st_r[k] < p.Ho) {
            const uint32_t dstp = stg_u32 + st_off[k];
            // SWIZZLE_128B (128-byte pixels) / SWIZZLE_64B (64-byte pixels) chunk XOR from the staged pixel index
            const uint32_t sw = BN == 64 ? ((st_off[k] >> 7) & 7u) : ((st_off[k] >> 7) & 3u);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              sts128(dstp + (((j * 4 + i) ^ sw) << 4), o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tmem_empty[buf]);
        else mbar_arrive_cluster(buf ? tmem_empty_leader1 : tmem_empty_leader0);
      }
      fence_proxy_async_smem();                   // this thread's st.shared -> visible to the TMA (async proxy)
      if (storer) tma_store_wait_read<0>();       // tile it-1's stores have left the OTHER buffer (next tile's)
      named_bar_sync(1, 256);                     // all 8 epilogue warps: tile staged, other buffer free
      if (storer && tile_ok && !(p.dbg & 10)) {
        for (int r = 0; r < p.R; ++r)
          if (h0 + r < p.Ho) tma_store_4d(&tmOut, stg + r * p.WP8 * ROWB, 0, 0, h0 + r, img);
        tma_store_commit();
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
struct StemOp : Op {
  CUtensorMap tmA, tmB, tmOut;
  StemParams p;
  int bn, grid, smem_bytes;
  bool pool = false;
  cudaError_t launch(cudaStream_t s) override;
};

template <int BN, int T, bool POOL>
static cudaError_t launch_stem(const StemOp& op, cudaStream_t s) {
  static std::atomic<uint64_t> attr_done{0};   // per device (see runtime.h)
  if (cudaError_t e = set_max_smem_once(stem_halo_kernel<BN, T, POOL>, 232448, attr_done)) return e;
  return launch_pdl(stem_halo_kernel<BN, T, POOL>, dim3(op.grid), dim3(ST_THREADS), op.smem_bytes, s, op.tmA, op.tmB, op.tmOut,
                    op.p);
}

cudaError_t StemOp::launch(cudaStream_t s) {
  g_launches++;
  if (pool) {   // BN = 64 only (stem_pool_geometry)
    switch (p.T) {
      case 2: return launch_stem<64, 2, true>(*this, s);
      case 3: return launch_stem<64, 3, true>(*this, s);
      default: return launch_stem<64, 4, true>(*this, s);
    }
  }
  switch (p.T) {
    case 2: return bn == 32 ? launch_stem<32, 2, false>(*this, s) : launch_stem<64, 2, false>(*this, s);
    case 3: return bn == 32 ? launch_stem<32, 3, false>(*this, s) : launch_stem<64, 3, false>(*this, s);
    default: return bn == 32 ? launch_stem<32, 4, false>(*this, s) : launch_stem<64, 4, false>(*this, s);
  }
}

// Fused max pool (PCV_CONV_POOL3S2): rows per tile R for a Ho x Wo conv map with T x T taps, or 0.  The pooled kernel
// needs one conv row per 128-row M-block (PW = 128 >= Wo + T - 1: TMEM lane == column), an even R = NMB <= 4 that
// divides Ho, and maps wide enough that padding every row to 128 pixels beats the separate pool kernel's extra pass.
constexpr int ST_POOL_PW = BLOCK_M;
static int stem_pool_rows(int Ho, int Wo, int T, int Cout) {
  if (Cout != 64 || Ho % 2 != 0 || Wo % 2 != 0 || Wo + T - 1 > ST_POOL_PW || Wo < 64) return 0;
  return Ho % 4 == 0 ? 4 : 2;
}
int stem_pool_ok(int C, int H, int W, int k, int Cout) {
  if (C < 1 || C > 4 || (k != 3 && k != 5 && k != 7) || H % 2 != 0 || W % 2 != 0) return 0;
  return stem_pool_rows(H / 2, W / 2, k / 2 + 1, Cout) > 0;
}

// The s2d stem as recorded by the plan compiler (plan.py::_conv_s2d): kh = T, kw = 1, Cin = T*16 through an
// overlapping-window view (PCV_CONV_IN_OVERLAP) of the [N, Ho+T-1, Wo+T-1, 16] s2d tensor; weights packed by
// igemm_pack as [Cout, T * 64] with K index = fr*64 + fs*16 + ch.  Returns PCV_ERR_UNSUPPORTED (message untouched)
// when the layer is outside this kernel's domain.
int stem_halo_try_make(const pcv_conv_desc& d, const void* x, const void* w, const float* bias, const void* res,
                       void* y, Op** out) {
  static const bool enabled = [] {
    const char* e = getenv("PCV_STEM_HALO");
    return !(e && e[0] == '0');
  }();
  const int T = d.kh;
  const int out_pitch = pitch_or(d.out_pitch, d.Cout);
  const bool pool = (d.flags & PCV_CONV_POOL3S2) != 0;
  if (!enabled || res || !(d.flags & PCV_CONV_IN_OVERLAP) || (d.flags & ~(PCV_CONV_IN_OVERLAP | PCV_CONV_POOL3S2)) || d.kw != 1 || T < 2 ||
      T > 4 || d.Cin != T * 16 || d.in_pitch != 16 || d.stride != 1 || d.pad != 0 || d.dil != 1 || d.groups != 1 ||
      (d.Cout != 32 && d.Cout != 64) || d.in_row_pitch != (d.W + T - 1) * 16 || out_pitch % 8 != 0 ||
      d.W + T - 1 > 256 || d.act > PCV_ACT_RELU6 || reinterpret_cast<uintptr_t>(x) % 16 != 0 ||
      reinterpret_cast<uintptr_t>(y) % 16 != 0 || reinterpret_cast<uintptr_t>(w) % 16 != 0)
    return PCV_ERR_UNSUPPORTED;
  const int BN = d.Cout, Wo = d.W, Ho = d.H - T + 1, PW = Wo + T - 1, rows = d.H;
  const int ntaps = T * T;
  const int b_bytes = round_up(ntaps * (BN / 2) * ST_ROW, 1024);
  const int budget = 232448 - 1024 - 256 - b_bytes;
  const int pairs = sm_count() / 2;
  const int WP8 = round_up(Wo, 8);
  double best = 0.0;
  int bestR = 0, bestNA = 0, bestNMB = 0, best_buf = 0, best_stg = 0;
  const int poolR = pool ? stem_pool_rows(Ho, Wo, T, d.Cout) : 0;
  if (pool && poolR == 0) return PCV_ERR_UNSUPPORTED;
  if (pool) {
    // one conv row per M-block: the A box is ST_POOL_PW pixels wide (columns past the s2d row are zero-filled by the TMA)
    const int R = poolR, NMB = R;
    const int buf = round_up((NMB * BLOCK_M + (T - 1) * ST_POOL_PW + T) * ST_ROW, 1024);
    const int stg = round_up((R / 2) * WP8 * BN * 2, 1024);                                // vertically pooled rows
    const int pooled = 2 * (R / 2) * round_up(Wo / 2, 8) * BN * 2;                         // 2 buffers of pooled rows
    const int NA = std::min(ST_MAX_NA, (budget - stg - round_up(pooled, 1024)) / buf);
    if (NA < 2) return PCV_ERR_UNSUPPORTED;
    bestR = R; bestNA = NA; bestNMB = NMB; best_buf = buf; best_stg = stg;
  }
  const int pool_bytes = pool ? round_up(2 * (poolR / 2) * round_up(Wo / 2, 8) * BN * 2, 1024) : 0;
  for (int R = 1; !pool && R <= std::min(Ho, 64); ++R) {
    const int Q = R * PW, NMB = ceil_div(Q, BLOCK_M);
    if (2 * NMB * BN > 512 || R + T - 1 > 256) continue;
    const int buf = round_up((NMB * BLOCK_M + (T - 1) * PW + T) * ST_ROW, 1024);
    const int stg = round_up(R * WP8 * BN * 2, 1024);
    const int NA = std::min(ST_MAX_NA, (budget - 2 * stg) / buf);
    if (NA < 2) continue;
    const int tiles = d.N * ceil_div(Ho, R);
    const int pair_tiles = (tiles + 1) / 2;
    const double wave = static_cast<double>(pair_tiles) / (ceil_div(pair_tiles, pairs) * pairs);
    const double useful = static_cast<double>(Ho) * Wo / (static_cast<double>(ceil_div(Ho, R)) * NMB * BLOCK_M);
    const double halo = static_cast<double>(R) / (R + T - 1);
    const double score = wave * useful * (0.8 + 0.2 * halo);
    if (score > best) {
      best = score; bestR = R; bestNA = NA; bestNMB = NMB; best_buf = buf; best_stg = stg;
    }
  }
  if (bestR == 0) return PCV_ERR_UNSUPPORTED;

  auto op = std::make_unique<StemOp>();
  StemParams& p = op->p;
  p.bias = bias;
  p.out = reinterpret_cast<e16*>(y);
  p.out_pitch = out_pitch;
  p.N = d.N; p.Ho = Ho; p.Wo = Wo; p.Cout = d.Cout;
  const int PWK = pool ? ST_POOL_PW : PW;   // pixels per staged A row (the kernel's row pitch)
  p.R = bestR; p.PW = PWK; p.NMB = bestNMB; p.T = T;
  p.tiles_per_img = ceil_div(Ho, bestR);
  p.num_tiles = d.N * p.tiles_per_img;
  p.NA = bestNA;
  p.a_buf_bytes = best_buf;
  p.a_tx_bytes = (bestR + T - 1) * PWK * ST_ROW;
  p.WP8 = WP8;
  p.stg_bytes = best_stg;
  p.act_lo = (d.act == PCV_ACT_RELU || d.act == PCV_ACT_RELU6) ? 0.f : -INFINITY;
  p.act_hi = d.act == PCV_ACT_RELU6 ? 6.f : INFINITY;
  {
    const char* e = getenv("PCV_STEM_DBG");
    p.dbg = e ? atoi(e) : 0;
  }
  op->bn = BN;
  op->pool = pool;
  op->smem_bytes = 1024 + b_bytes + bestNA * best_buf + (pool ? 1 : 2) * best_stg + pool_bytes + 256;
  op->grid = 2 * std::min((p.num_tiles + 1) / 2, pairs);
  p.tpc = ceil_div(p.num_tiles, op->grid);
  p.Hp = Ho / 2; p.Wp = Wo / 2; p.WPP8 = round_up(Wo / 2, 8);

  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  {
    cuuint64_t dims[4] = {16, (cuuint64_t)PW, (cuuint64_t)rows, (cuuint64_t)d.N};
    cuuint64_t strides[3] = {ST_ROW, (cuuint64_t)PW * ST_ROW, (cuuint64_t)rows * PW * ST_ROW};
    cuuint32_t box[4] = {16, (cuuint32_t)PWK, (cuuint32_t)(bestR + T - 1), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&op->tmA, TMAP_E16, 4, const_cast<void*>(x), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled (stem A) failed (%d)", (int)r);
  }
  {
    const uint64_t kpad = static_cast<uint64_t>(T) * BLOCK_K;
    cuuint64_t dims[2] = {kpad, (cuuint64_t)d.Cout};
    cuuint64_t strides[1] = {kpad * 2};
    cuuint32_t box[2] = {16, (cuuint32_t)(BN / 2)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&op->tmB, TMAP_E16, 2, const_cast<void*>(w), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled (stem B) failed (%d)", (int)r);
  }
  {   // output rows leave through 4-D TMA stores: box = one image row (Wo pixels x BN channels)
    const int Hy = pool ? Ho / 2 : Ho, Wy = pool ? Wo / 2 : Wo;   // y is the pooled map when the max pool is fused
    cuuint64_t dims[4] = {(cuuint64_t)d.Cout, (cuuint64_t)Wy, (cuuint64_t)Hy, (cuuint64_t)d.N};
    cuuint64_t strides[3] = {(cuuint64_t)out_pitch * 2, (cuuint64_t)Wy * out_pitch * 2, (cuuint64_t)Hy * Wy * out_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)BN, (cuuint32_t)Wy, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&op->tmOut, TMAP_E16, 4, y, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, BN == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled (stem out) failed (%d)", (int)r);
  }
  char nm[160];
  snprintf(nm, sizeof nm, "conv_stem%s s2d %dx%d taps x16ch ->%d @%dx%d bn=%d halo R=%d mb=%d na=%d", pool ? "+maxpool3s2" : "",
           T, T, d.Cout, Ho, Wo, BN, bestR, bestNMB, bestNA);
  op->name = nm;
  const double M = static_cast<double>(d.N) * Ho * Wo;
  op->flops = 2.0 * M * d.Cout * ntaps * 16;
  // algorithmic bytes: the s2d tensor once (what this kernel's input really is) + the output + weights + bias
  // (fused max pool: the conv map never reaches HBM, only the pooled map does)
  op->bytes = 2.0 * d.N * rows * PW * 16 + 2.0 * (pool ? M / 4 : M) * d.Cout + 2.0 * d.Cout * ntaps * 16 + 4.0 * d.Cout;
  *out = op.release();
  return PCV_OK;
}

}  // namespace PCV_TIER
}  // namespace pcv
