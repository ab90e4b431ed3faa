// CTA-pair variant of the implicit-GEMM convolution: tcgen05.mma.cta_group::2, 256 x BN output tile per cluster.
//
// Why: with 128 x 128 single-CTA tiles every SM has to pull 32 KB of operands from L2 per 256 MMA cycles, and the
// L2 -> SMEM fabric (measured ~11.7 TB/s on B200, round-1 profiles) saturates long before the tensor pipe or HBM.
// Pairing the two SMs of a TPC halves that traffic: each CTA loads its own 128 pixel rows of A and only HALF of the
// B (weight) tile; one tcgen05.mma (M = 256, N = BN <= 256) issued by the leader CTA consumes A/B from both CTAs'
// shared memory and writes 128 x BN accumulators into EACH CTA's TMEM.  BN = 256 also halves how often an A tile is
// re-read for wide layers.
//
// Roles per CTA (8 warps) are those of conv_igemm.cu; differences:
//   * cluster of 2 (rank 0 = leader).  Both producers issue `cp.async.bulk.tensor...cta_group::2` loads that
//     complete on the LEADER's full barrier (count 1: the leader's arrive.expect_tx covers both CTAs' bytes).  Only the leader's warp 1 issues MMAs; `tcgen05.commit...multicast::cluster` (mask 0b11)
//     releases the smem stage / publishes the accumulator in both CTAs.
//   * the accumulator-free barrier lives in the leader and counts the 4 epilogue warps of BOTH CTAs (peer arrives
//     through mapa + mbarrier.arrive.shared::cluster).
//   * the epilogue's staging ring works on 64-column sub-tiles (16 KB): residual TMA load -> math -> TMA store per
//     sub-tile, NSTG slots, up to PENDING stores in flight, so the residual of sub-tile u+NSTG-PENDING is prefetched
//     while sub-tile u is computed and no per-tile bubble remains.
//   * the tile width is NS*64 <= BN output columns (template parameter, chosen per layer on the host so that
//     ceil(Cout/64) splits into equal tiles): Cout = 144 runs as ONE 192-wide tile (N = 192 MMAs, 3 epilogue rounds), not
//     as a 256-wide tile with a quarter of the epilogue work spent on padding columns; 576 = 3 x 192, 960 = 5 x 192.
//     BN only sizes the shared-memory stages and the TMEM double buffer.
#include "igemm_common.cuh"

namespace pcv {
namespace PCV_TIER {

constexpr int NUM_THREADS2 = 384;   // 4 control warps (physical warps 8-11) + 8 epilogue warps

template <int BN, int NS = BN / 64>
struct Pair {
  static constexpr int TW = NS * 64;                        // tile width in output columns (<= BN)
  static constexpr int HALF_N = TW / 2;                     // weight rows each CTA of the pair loads per K block
  static constexpr int B_STAGE_BYTES = (BN / 2) * BLOCK_K * 2;   // stage stride (sized for the full width)
  static constexpr int B_TX_BYTES = HALF_N * BLOCK_K * 2;        // bytes one weight box really delivers
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int SUB_COLS = 64;
  static constexpr int SUB_BYTES = BLOCK_M * SUB_COLS * 2;  // 16 KiB
  static constexpr int NSUB = NS;                           // sub-tiles per output tile
  static constexpr int MAX_STAGES = 8, MAX_NSTG = 8;
  static constexpr int NUM_BARS = 2 * MAX_STAGES + 4 + 3 * MAX_NSTG;
  static constexpr uint32_t TMEM_COLS = 2 * BN;             // double-buffered accumulator (256 or 512 columns)
  // shared memory: [stages x A][stages x B][nstg x staging][barriers]; the split is chosen per layer at run time
  // (long-K tensor-bound layers want a deep operand ring, short-K bandwidth-bound layers a deep staging ring)
  static constexpr int SMEM_LIMIT = 232448;
  __host__ __device__ static constexpr int bytes(int stages, int ksub, int nstg) {
    return stages * ksub * STAGE_BYTES + nstg * SUB_BYTES + NUM_BARS * 8 + 16 + 1024;
  }
};

template <int BN, int NS, bool GATE, bool DUAL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS2, 1)
igemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
              const IgemmParams p) {
  using L = Pair<BN, NS>;
  constexpr int TW = L::TW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STAGES = p.stages, NSTG = p.nstg, KSUB = p.ksub;
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * KSUB * A_STAGE_BYTES;
  uint8_t* sStg = sB + STAGES * KSUB * L::B_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + NSTG * L::SUB_BYTES);
  uint64_t* full = bars;                            // [STAGES]  (leader's copy is the live one)
  uint64_t* empty = bars + L::MAX_STAGES;           // [STAGES]  per CTA, released by the multicast commit
  uint64_t* tmem_full = bars + 2 * L::MAX_STAGES;   // [2]       per CTA, multicast commit
  uint64_t* tmem_empty = tmem_full + 2;             // [2]       leader's copy, 8 arrivals
  uint64_t* stg_free = tmem_empty + 2;              // [NSTG]  staging manager -> epilogue: slot may be overwritten
  uint64_t* res_full = stg_free + L::MAX_NSTG;      // [NSTG]  residual TMA -> epilogue: residual sub-tile landed
  uint64_t* stg_full = res_full + L::MAX_NSTG;      // [NSTG]  epilogue (4 warps) -> staging manager: result written
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

  // Warp roles.  The scheduler arbitrates highest-warp-id-first, so the latency-critical single-lane roles (TMA
  // producer, MMA issuer, staging manager) take the HIGH warp ids 4-7 and the four epilogue warps the low ids 0-3
  // (TMEM lane quarter = physical warp id & 3 either way).  `warp` below is the ROLE index: 0 producer, 1 MMA,
  // 2 TMEM allocator, 3 staging manager, 4-7 epilogue.
  const int pw = threadIdx.x >> 5;
  const int warp = pw >= 8 ? pw - 8 : pw + 4;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();        // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int pair_tiles_m = (p.tiles_m + 1) >> 1;
  const int num_tiles = pair_tiles_m * p.tiles_n;  // pair tiles: 256 rows x BN columns

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if (p.has_res || DUAL) tma_prefetch_desc(&tmRes);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);   // leader's arrive.expect_tx covers both CTAs' bytes
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 16);   // 8 epilogue warps x 2 CTAs
    }
    for (int i = 0; i < NSTG; ++i) {
      mbar_init(&stg_free[i], 1);
      mbar_init(&res_full[i], 1);
      mbar_init(&stg_full[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_ptr, L::TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();   // peer barriers initialised, both TMEM allocations done
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (tmem_base != 0) __trap();   // one CTA per SM => the allocation starts at column 0; the MMA warp relies on it
  pdl_launch_dependents();        // the next kernel may start its prologue as SMs drain ...
  pdl_wait();                     // ... and this one touches activations only after its predecessor completed

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) =====================================
    // One full/empty handshake per STAGE; a stage holds KSUB 64-channel sub-blocks (KSUB chosen per layer so that
    // a stage carries >= ~512 tensor-pipe cycles of work: the single-thread handshake costs ~400-600 cycles).
    // whole warp, warp-uniform values; only the TMA / mbarrier instructions are predicated on elect.sync so that ptxas
    // keeps addresses and coordinates in uniform registers (an `if (lane == 0)` region wraps every UTMALDG / UTCHMMA
    // in a ~13-instruction ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += npairs) {
        const int pm = t / p.tiles_n;
        const int n_tile = t - pm * p.tiles_n;
        int m0 = (2 * pm + static_cast<int>(rank)) * BLOCK_M;
        if (m0 >= p.M) m0 = 0;  // phantom second tile of an odd tail pair: load anything valid, its store is clipped
        const int img = m0 / p.HoWo;
        const int rem = m0 - img * p.HoWo;
        const int ho = rem / p.Wo;
        const int wo = rem - ho * p.Wo;
        const int w0 = wo * p.stride - p.pad;
        const int h0 = ho * p.stride - p.pad;
        const int b_row = n_tile * TW + static_cast<int>(rank) * L::HALF_N;
        const int c_base = p.grouped ? n_tile * p.g_in_span : 0;
        int cb = 0, fr = 0, fs = 0;
        for (int kb = 0; kb < p.num_kblocks; kb += KSUB) {
          const int nsub = min(KSUB, p.num_kblocks - kb);
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t full_leader = mapa_u32(smem_u32(&full[stage]), 0);
          // only the leader arrives; the peer's bytes may land first (tx-count goes negative, the phase cannot
          // complete before the leader's arrival), exactly the CUTLASS 2-SM pipeline protocol
          if (rank == 0 && elect_one()) mbar_arrive_expect_tx(&full[stage], 2 * nsub * (A_STAGE_BYTES + L::B_TX_BYTES));
          uint8_t* a_dst = sA + stage * KSUB * A_STAGE_BYTES;
          uint8_t* b_dst = sB + stage * KSUB * L::B_STAGE_BYTES;
          for (int j = 0; j < nsub; ++j) {
            if (DUAL && kb + j >= p.kb_split) {
              // second source (pcv_conv1x1_dual): the unit's input under the projection shortcut's (strided) 1x1 conv - its
              // tensor map sits in the residual slot, its weights follow the first source's along K
              if (elect_one()) {
                const int c2 = (kb + j - p.kb_split) * BLOCK_K;
                if (p.a_mode2 == 1) {
                  tma2_load_im2col_4d(&tmRes, full_leader, a_dst, c2, wo * p.stride2, ho * p.stride2, img, 0, 0);
                } else {
                  tma2_load_2d(&tmRes, full_leader, a_dst, c2, m0);
                }
                tma2_load_2d(&tmB, full_leader, b_dst, (kb + j) * BLOCK_K, b_row);
              }
              a_dst += A_STAGE_BYTES;
              b_dst += L::B_STAGE_BYTES;
              continue;
            }
            if (elect_one()) {
              if (p.a_mode == 1) {
                tma2_load_im2col_4d(&tmA, full_leader, a_dst, c_base + cb * BLOCK_K, w0, h0, img,
                                    static_cast<uint16_t>(fs * p.dil), static_cast<uint16_t>(fr * p.dil));
              } else {
                tma2_load_2d(&tmA, full_leader, a_dst, c_base + cb * BLOCK_K, m0);
              }
              tma2_load_2d(&tmB, full_leader, b_dst, (kb + j) * BLOCK_K, b_row);
            }
            a_dst += A_STAGE_BYTES;
            b_dst += L::B_STAGE_BYTES;
            if (++cb == p.cblocks) {
              cb = 0;
              if (++fs == p.kw) {
                fs = 0;
                ++fr;
              }
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =====================================
    if (rank == 0) {   // whole warp (see the producer's note); tcgen05 instructions under elect.sync
      constexpr uint32_t idesc = make_idesc_e16(2 * BLOCK_M, TW);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA)), b_lo0 = smem_desc_lo(smem_u32(sB));
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair; t < num_tiles; t += npairs, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[buf], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = buf * BN;   // TMEM base is 0 (one CTA per SM, checked after the allocation)
        uint32_t accum = 0;
        for (int kb = 0; kb < p.num_kblocks; kb += KSUB) {
          const int nsub = min(KSUB, p.num_kblocks - kb);
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          uint32_t a_lo = a_lo0 + stage * KSUB * (A_STAGE_BYTES >> 4);
          uint32_t b_lo = b_lo0 + stage * KSUB * (L::B_STAGE_BYTES >> 4);
          for (int j = 0; j < nsub; ++j) {
            uint32_t d_acc = d_tmem, acc0 = accum;
            if constexpr (GATE && DUAL) {
              // gated unit with a projection shortcut: the SE gate multiplies conv3's half of the sum only, so the shortcut's
              // k-blocks accumulate into a SECOND accumulator (columns TW .. 2 TW of the buffer) that the epilogue adds un-gated
              static_assert(2 * TW <= BN, "two accumulators per TMEM buffer");
              if (kb + j >= p.kb_split) d_acc = d_tmem + TW;
              if (kb + j == p.kb_split) acc0 = 0;
            }
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / 16; ++k) umma2_bf16_lohi(d_acc, a_lo + 2 * k, b_lo + 2 * k, idesc, k ? 1u : acc0);
            }
            accum = 1;
            a_lo += A_STAGE_BYTES >> 4;
            b_lo += L::B_STAGE_BYTES >> 4;
          }
          if (elect_one()) umma2_commit(&empty[stage], 0x3);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma2_commit(&tmem_full[buf], 0x3);
      }
    }
  } else if (warp == 3) {
    // ===================================== staging manager (both CTAs) =====================================
    // Owns every TMA operation on the staging ring: stores the finished 64-column sub-tiles and refills freed slots
    // with the residual of the sub-tile that will use them NSTG rounds later.  The epilogue warps never wait for a
    // store to drain or for each other - they only wait for data (res_full) or a free slot (stg_free) - so the
    // per-round latency chain (bar.sync + store issue + read-out wait) that paced every short-K layer is gone.
    const int my_tiles = pair < num_tiles ? (num_tiles - pair + npairs - 1) / npairs : 0;
    const int total_rounds = my_tiles * L::NSUB;
    auto round_coords = [&](int r, int& col0, int& row0) {
      const int t = pair + (r / L::NSUB) * npairs;
      const int pm = t / p.tiles_n;
      const int n_tile = t - pm * p.tiles_n;
      col0 = n_tile * TW + (r % L::NSUB) * L::SUB_COLS;
      row0 = (2 * pm + static_cast<int>(rank)) * BLOCK_M;
    };
    auto refill = [&](int r) {   // slot of round r is free: fetch its residual, or tell the epilogue it may write
      if (r >= total_rounds) return;
      const int s = r % NSTG;
      if (elect_one()) {
        if (p.has_res) {
          int col0, row0;
          round_coords(r, col0, row0);
          mbar_arrive_expect_tx(&res_full[s], L::SUB_BYTES);
          tma_load_2d(&tmRes, &res_full[s], sStg + s * L::SUB_BYTES, col0, row0);
        } else {
          mbar_arrive(&stg_free[s]);
        }
      }
    };
    for (int r = 0; r < NSTG; ++r) refill(r);
    for (int r = 0; r < total_rounds; ++r) {
      const int s = r % NSTG;
      mbar_wait(&stg_full[s], (r / NSTG) & 1);
      if (elect_one()) {   // elect.sync is deterministic: one lane owns every bulk group
        int col0, row0;
        round_coords(r, col0, row0);
        tma_store_2d(&tmOut, sStg + s * L::SUB_BYTES, col0, row0);
        tma_store_commit();
        tma_store_wait_read<1>();   // the store of round r-1 has left its slot
      }
      __syncwarp();
      if (r >= 1) refill(r - 1 + NSTG);
    }
    if (elect_one()) tma_store_wait_read<0>();
    __syncwarp();
    if (elect_one()) tma_store_wait_all<0>();
  } else if (warp >= 4) {
    // ===================================== epilogue (both CTAs) =====================================
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;   // eight epilogue warps: (lane quarter, 32-column half)
    const int row = q * 32 + lane;
    const bool fancy_act = p.act > PCV_ACT_RELU6;
    const bool relu = p.act_lo == 0.f, capped = p.act_hi != INFINITY;   // the clamp family: none / ReLU / ReLU6
    const uint32_t cap2 = pack_e16x2(p.act_hi, p.act_hi);
    const uint32_t sStg_u32 = smem_u32(sStg);
    const uint32_t tmem_empty_leader0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    const uint32_t tmem_empty_leader1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
    int it = 0;
    int slot = 0;
    uint32_t sphase = 0;
    for (int t = pair; t < num_tiles; t += npairs, ++it) {
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int pm = t / p.tiles_n;
      const int n_tile = t - pm * p.tiles_n;
      const int n0 = n_tile * TW;

      mbar_wait(&tmem_full[buf], acc_phase);
      tc_fence_after();

#pragma unroll 1
      for (int sub = 0; sub < L::NSUB; ++sub) {
        const uint32_t stg_u32 = sStg_u32 + slot * L::SUB_BYTES;
        // the slot is ready when its residual has landed (which implies the previous store has left it) or, without
        // a residual, when the staging manager has released it
        mbar_wait(p.has_res ? &res_full[slot] : &stg_free[slot], sphase);
        {
          const int h = half;
          const int col = sub * L::SUB_COLS + h * 32;
          uint32_t acc[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + col, acc);
          // bias and residual loads issued under the TMEM load (the wait below is a compiler barrier for memory operations)
          const float4* bias4 = reinterpret_cast<const float4*>(p.bias + n0 + col);
          float4 b4[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)   // columns past Cout (tile wider than the layer) are clipped by the store: no read there
            b4[i] = (n0 + col + 4 * i < p.Cout) ? __ldg(bias4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          const uint32_t row_off = row * 128 + h * 64;
          const uint32_t sw = (row & 7u) << 4;   // SWIZZLE_128B: 16-byte chunk index XOR (row mod 8)
          uint4 r4[4];
          if (p.has_res) {
#pragma unroll
            for (int c = 0; c < 4; ++c) r4[c] = lds128(stg_u32 + ((row_off + c * 16) ^ sw));
          }
          // SE scale in the epilogue (PCV_CONV_SE_GATE): the row's image picks the gate vector (the rows of a warp mostly share
          // it: broadcast loads), fetched under the TMEM load like the bias
          // (a template parameter, not a run-time branch: with the gate code in the common instantiation the kernel grew from
          // 114 to 150 registers and ResNet-50 lost 1.4 %)
          float4 g4[8];
          if constexpr (GATE) {
            const int m = (2 * pm + static_cast<int>(rank)) * BLOCK_M + row;
            const int img = min(m / p.HoWo, p.n_img - 1);
            const float4* gp = reinterpret_cast<const float4*>(p.gate + static_cast<size_t>(img) * p.Cout + n0 + col);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              g4[i] = (n0 + col + 4 * i < p.Cout) ? __ldg(gp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
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
          if constexpr (GATE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[4 * i + 0] *= g4[i].x; v[4 * i + 1] *= g4[i].y; v[4 * i + 2] *= g4[i].z; v[4 * i + 3] *= g4[i].w;
            }
          }
          if constexpr (GATE && DUAL) {
            // + the projection shortcut from the second accumulator (and its own folded bias), outside the gate
            uint32_t acc2[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + TW + col, acc2);
            const float4* bias24 = reinterpret_cast<const float4*>(p.bias2 + n0 + col);
            tmem_ld_wait_regs(acc2);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b2 = (n0 + col + 4 * i < p.Cout) ? __ldg(bias24 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
              v[4 * i + 0] += __uint_as_float(acc2[4 * i + 0]) + b2.x;
              v[4 * i + 1] += __uint_as_float(acc2[4 * i + 1]) + b2.y;
              v[4 * i + 2] += __uint_as_float(acc2[4 * i + 2]) + b2.z;
              v[4 * i + 3] += __uint_as_float(acc2[4 * i + 3]) + b2.w;
            }
          }
          if (p.has_res) {
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
          }
          uint32_t o[16];
          if (fancy_act) {
            fast_act_n(v, p.act, p.act_a);
            clamp_pack32(v, o, false, false, 0u);
          } else {
            clamp_pack32(v, o, relu, capped, cap2);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
            sts128(stg_u32 + ((row_off + c * 16) ^ sw), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
        }
        if (sub == L::NSUB - 1) {
          // all accumulator columns of this tile are in registers / smem: release the TMEM buffer to the leader's MMA
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&tmem_empty[buf]);
            else mbar_arrive_cluster(buf ? tmem_empty_leader1 : tmem_empty_leader0);
          }
        }
        fence_proxy_async_smem();   // this thread's st.shared -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&stg_full[slot]);   // 4 arrivals (one per epilogue warp) hand the slot to the manager
        if (++slot == NSTG) {
          slot = 0;
          sphase ^= 1;
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // no CTA may exit (or free TMEM) while its peer can still multicast into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, L::TMEM_COLS);
  }
}

template <int BN, int NS, bool GATE, bool DUAL = false>
static cudaError_t launch_pair_g(int grid, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                                 const CUtensorMap& tmRes, const IgemmParams& p, cudaStream_t s) {
  using L = Pair<BN, NS>;
  static std::atomic<uint64_t> attr_done{0};   // per device (see runtime.h)
  if (cudaError_t e = set_max_smem_once(igemm2_kernel<BN, NS, GATE, DUAL>, L::SMEM_LIMIT, attr_done)) return e;
  return launch_pdl(igemm2_kernel<BN, NS, GATE, DUAL>, dim3(grid), dim3(NUM_THREADS2), L::bytes(p.stages, p.ksub, p.nstg), s, tmA,
                    tmB, tmOut, tmRes, p);
}
// the gated epilogue (PCV_CONV_SE_GATE) is its own instantiation of the full-width tiles that SE units' last 1x1 convs use
template <int BN, int NS>
static cudaError_t launch_pair(int grid, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                               const CUtensorMap& tmRes, const IgemmParams& p, cudaStream_t s) {
  if constexpr (BN == 256 && NS == 2) {   // gated dual-source conv: two 128-column accumulators per TMEM buffer
    if (p.gate != nullptr && p.kb_split > 0) return launch_pair_g<BN, NS, true, true>(grid, tmA, tmB, tmOut, tmRes, p, s);
  }
  if constexpr (NS * 64 == BN) {
    if (p.kb_split == 0 && p.gate != nullptr) return launch_pair_g<BN, NS, true>(grid, tmA, tmB, tmOut, tmRes, p, s);
    if constexpr (BN == 256) {   // the dual-source producer is instantiated for the tile the bottleneck tails use
      if (p.kb_split > 0 && p.gate == nullptr) return launch_pair_g<BN, NS, false, true>(grid, tmA, tmB, tmOut, tmRes, p, s);
    }
  }
  if (p.gate != nullptr || p.kb_split > 0) return cudaErrorInvalidValue;   // igemm_make keeps such layers on full-width 256 tiles
  return launch_pair_g<BN, NS, false>(grid, tmA, tmB, tmOut, tmRes, p, s);
}

// Per-layer shared-memory split: K sub-blocks per stage (amortises the per-stage handshake over >= ~512 MMA cycles),
// operand-ring depth, staging slots.  Long-K (tensor-bound) layers get a deep operand ring and a 2-slot staging ring;
// short-K (bandwidth-bound) layers a shallow operand ring and a deep staging ring for residual prefetch.
void igemm2_pick_smem(int bn, int num_kblocks, bool has_res, int taps, int* stages, int* ksub, int* nstg) {
  const int stage_bytes = A_STAGE_BYTES + (bn / 2) * BLOCK_K * 2;
  const int sub_bytes = BLOCK_M * 64 * 2;
  const int fixed = (2 * 8 + 4 + 2 * 8) * 8 + 16 + 1024;
  const int budget = 232448 - fixed;
  int ks = bn >= 256 ? 2 : 4;                       // MMA cycles per 64-K sub-block: 512 (bn 256), 256, 128
  ks = std::max(1, std::min(ks, num_kblocks));
  int ns = num_kblocks >= 8 ? 2 : (has_res ? 6 : 4);
  int st = std::min(8, (budget - ns * sub_bytes) / (ks * stage_bytes));
  while (st < 2 && ks > 1) {
    ks /= 2;
    st = std::min(8, (budget - ns * sub_bytes) / (ks * stage_bytes));
  }
  if (num_kblocks < 8) st = std::min(st, std::max(2, 4 / ks + 1));
  ns = std::max(2, std::min(8, (budget - st * ks * stage_bytes) / sub_bytes));
  if (taps == 1 && num_kblocks >= 8 && bn == 256) {
    // long-K pointwise layers are paced by the epilogue round trip (store read-out + barrier), not by operand
    // latency: measured sweep (profiles/README.md) prefers one k-block per stage and >= 4 staging slots
    ks = 1;
    ns = has_res ? 6 : 4;
    st = std::min(8, (budget - ns * sub_bytes) / stage_bytes);
  }
  if (const char* e = getenv("PCV_IGEMM2_STAGES")) st = atoi(e);
  if (const char* e = getenv("PCV_IGEMM2_KSUB")) ks = atoi(e);
  if (const char* e = getenv("PCV_IGEMM2_NSTG")) ns = atoi(e);
  *stages = st;
  *ksub = ks;
  *nstg = ns;
}

cudaError_t launch_igemm2(int bn, int grid, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                          const CUtensorMap& tmRes, const IgemmParams& p, cudaStream_t s) {
  switch (bn) {
    case 256:
      if (p.nsubs == 3) return launch_pair<256, 3>(grid, tmA, tmB, tmOut, tmRes, p, s);
      if (p.nsubs == 2) return launch_pair<256, 2>(grid, tmA, tmB, tmOut, tmRes, p, s);
      return launch_pair<256, 4>(grid, tmA, tmB, tmOut, tmRes, p, s);
    case 128: return launch_pair<128, 2>(grid, tmA, tmB, tmOut, tmRes, p, s);
    default: return launch_pair<64, 1>(grid, tmA, tmB, tmOut, tmRes, p, s);
  }
}

}  // namespace PCV_TIER
}  // namespace pcv
