// The fp32 tier's dense / grouped convolution on the tensor cores (PCV_CONV_F32_SPLIT): fp32 in, fp32 weights, fp32 out,
// evaluated as a 3-way bf16 split
//     x = x0 + x1 + x2,   w = w0 + w1 + w2      (each part the bf16 rounding of what the earlier parts left over:
//                                                3 x 8 = 24 significant bits, i.e. the fp32 value itself)
//     x*w = x0 w0 + x1 w0 + x2 w0 + x0 w1 + x1 w1 + x0 w2 + (terms <= 2^-24 |x w|, dropped)
// Every kept term is an EXACT bf16 x bf16 product accumulated in fp32, so the result has the accuracy of an fp32 FMA chain
// (SURVEY 7.2 rules TF32 out for the <= 1e-4 tier: 10-bit operands give 4e-4 ... 2e-2 end to end) at tcgen05 rate: the
// CUDA-core kernel this replaces ran ResNet-18 bs8 at 8 TFLOP/s, 1 % of the machine.
//
// How the split rides on the implicit GEMM  D[m, n] = sum_k A[m, k] B[n, k]:
//   * a pre-kernel writes the split activation  xs[pixel, part * C + c]  (bf16, 3 C channels per pixel) into the op's
//     workspace; for stems (C <= 16, k x k taps) it writes the im2col row instead, xs[m, part * K + tap * C + c], and the
//     GEMM runs as a 1 x 1 convolution over K = taps * C channels (TMA im2col gathers of 16-byte pixels crawl);
//   * the K axis is widened, not the kernel's inner loop: per filter tap the k-blocks run over weight part q = 0, 1, 2 and,
//     inside q, over 64-channel windows of xs; the packed weight block of (q, window) holds w_q[o, c] where the window's
//     channel belongs to activation part a <= 2 - q, zero elsewhere.  With C < 64 several parts share a window
//     (C = 8 im2col'ed 7 x 7: K = 392, 39 k-blocks instead of 6 x 49);
//   * tcgen05 accumulates in fp32 but TRUNCATES each add (measured: the error of one long TMEM accumulation chain grows
//     linearly, ~1.7e-8 per MMA; 1728 MMAs = 2.9e-5): the chain is cut into chunks of F3_CHUNK_KB k-blocks whose results
//     the epilogue warps sum in registers with round-to-nearest fp32 adds (two-level accumulation);
//   * small-M layers (ResNet-18 stage 4 at batch 8: 4 M-tiles) are split along K over several CTAs; partial tiles go to the
//     workspace and the LAST CTA to finish a tile sums them in fixed order s = 0..S-1 (deterministic) and runs the epilogue.
//
// Where the time goes (ResNet-18 bs8, 1.42 ms/step; PCV_F3_DBG decomposition, profiles/README.md): per op ~25 us are fixed
// (two launches, TMEM / barrier prologue, split-K hand-over) and each k-block costs ~0.35 us of which 0.25 us remain with NO
// loads and NO MMAs issued - the producer / MMA-warp barrier hand-over per stage (a tcgen05.commit per 4 small MMAs), not
// bandwidth or math; busy-polling the barriers instead of try_wait changes nothing.  Grouping several k-blocks per stage
// (the KSUB of conv_igemm2.cu) and emitting the split activation from the producing conv's epilogue are the next steps.
#include "igemm_common.cuh"

namespace pcv {
namespace PCV_TIER {

constexpr int F3_PARTS = 3;
constexpr int F3_NBUF = 4;           // TMEM accumulator buffers (one chunk each): the MMA warp runs up to 4 chunks ahead of the epilogue
// warp 0 producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, then the epilogue warps: 4 for BN = 32 (one per TMEM lane
// quarter), 8 for BN >= 64 (lane quarter x column half: the chunk sums and the final bias / residual / store pass are the
// kernel's critical path once the loads and MMAs are hidden, so the work per warp is halved)
template <int BN>
struct F3Cfg {
  static constexpr int EPI_WARPS = BN >= 64 ? 8 : 4;
  static constexpr int COLS_PER_WARP = BN / (EPI_WARPS / 4);
  static constexpr int THREADS = 128 + EPI_WARPS * 32;
};
constexpr int F3_CHUNK_KB = 4;       // k-blocks (16 MMAs) per TMEM accumulation chain: ~3e-7 of truncation error, i.e. the
                                     // level of an fp32 FMA chain (ResNet-18 end to end: 3e-6 with 4, 1.4e-5 with 16, 4.6e-5 unchunked)

struct F3Params {
  const float* bias;
  float* out;
  const float* res;
  float* partial;        // [tile][S][128][BN] fp32 (S > 1)
  unsigned int* counter; // [tiles] arrivals per tile (zeroed by the pre-kernel of every launch)
  int M, Cout, out_pitch, res_pitch;
  int HoWo, Wo, stride, pad, dil, kw;
  int num_kblocks, tiles_m, tiles_n;
  int act, has_res;
  float act_lo, act_hi, act_a;
  int a_mode, grouped, g_in_span;
  int kq[3], per_tap, cstride;
  int S, kb_per_split;   // split-K factor and k-blocks per split
  int chunk_kb;          // k-blocks per TMEM accumulation chain (F3_CHUNK_KB; PCV_F3_CHUNK overrides for experiments)
  int vec_ok;            // out / residual rows are 16-byte aligned: float4 epilogue
  int ksub;              // k-blocks per full / empty hand-over (smem slot = ksub consecutive stages); divides chunk_kb
  int dbg;               // PCV_F3_DBG throughput experiments (WRONG results): 1 skip the A loads, 2 skip the B loads, 4 skip MMA issue
};

template <int BN>
struct F3Smem {
  static constexpr int STAGES = BN == 32 ? 10 : (BN == 64 ? 8 : 6);
  static constexpr int B_STAGE = BN * BLOCK_K * 2;
  static constexpr int OFF_B = STAGES * A_STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_B + STAGES * B_STAGE;
  static constexpr int NUM_BARS = 2 * STAGES + 2 * F3_NBUF;
  static constexpr int BYTES = OFF_BAR + NUM_BARS * 8 + 32 + 1024;
};

__device__ __forceinline__ float f3_act(float x, int act, float lo, float hi, float a) {
  if (act <= PCV_ACT_RELU6) return fminf(fmaxf(x, lo), hi);
  if (act == PCV_ACT_LEAKY_RELU) return x >= 0.f ? x : x * a;
  if (act == PCV_ACT_SIGMOID) return 1.f / (1.f + expf(-x));
  if (act == PCV_ACT_SWISH) return x / (1.f + expf(-x));
  if (act == PCV_ACT_HSWISH) return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) / 6.f;
  return fminf(fmaxf(x + 3.f, 0.f), 6.f) / 6.f;   // h-sigmoid
}

template <int BN>
__global__ void __launch_bounds__(F3Cfg<BN>::THREADS, 1)
f32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const F3Params p) {

  using L = F3Smem<BN>;
  constexpr int STAGES = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const long long t_entry = clock64();   // PCV_F3_DBG & 16: CTA 0 prints a timeline of its MMA warp (clocks)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + L::OFF_B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = tmem_full + F3_NBUF;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);
  int* last_flag = reinterpret_cast<int*>(tmem_ptr + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_units = p.tiles_m * p.tiles_n * p.S;
  constexpr uint32_t TMEM_COLS = F3_NBUF * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < F3_NBUF; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], F3Cfg<BN>::EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const long long t_prologue = clock64();
  pdl_launch_dependents();   // the next op's pre-kernel may start as SMs drain ...
  pdl_wait();                // ... and this kernel reads the split activation only after its own pre-kernel has completed
  const long long t_pdl = clock64();

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    const int nslots = STAGES / p.ksub;
    int slot = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int tile = u / p.S, s = u - tile * p.S;
      const int m_tile = tile / p.tiles_n, n_tile = tile - m_tile * p.tiles_n;
      const int m0 = m_tile * BLOCK_M;
      const int img = m0 / p.HoWo;
      const int rem = m0 - img * p.HoWo;
      const int ho = rem / p.Wo;
      const int wo = rem - ho * p.Wo;
      const int w0 = wo * p.stride - p.pad;
      const int h0 = ho * p.stride - p.pad;
      const int c_base = p.grouped ? n_tile * p.g_in_span : 0;
      const int kb0 = s * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.num_kblocks);
      // position of k-block kb0 inside (tap, weight part q, window cb)
      const int tap = kb0 / p.per_tap;
      int r = kb0 - tap * p.per_tap;
      int q = 0;
      while (r >= p.kq[q]) r -= p.kq[q++];
      int cb = r;
      int fr = tap / p.kw, fs = tap - fr * p.kw;
      for (int kb = kb0; kb < kb1; kb += p.ksub) {
        // one hand-over per slot of `ksub` k-blocks: the barrier round trip, not the loads, paces this loop (see the header)
        const int nsub = min(p.ksub, kb1 - kb);
        mbar_wait(&empty[slot], phase ^ 1);
        if (elect_one())
          mbar_arrive_expect_tx(&full[slot], nsub * (((p.dbg & 1) ? 0 : A_STAGE_BYTES) + ((p.dbg & 2) ? 0 : L::B_STAGE)));
        for (int j = 0; j < nsub; ++j) {
          const int stage = slot * p.ksub + j;
          if (elect_one()) {
            uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
            if (!(p.dbg & 1)) {
              if (p.a_mode == 1) {
                tma_load_im2col_4d(&tmA, &full[slot], a_dst, c_base + cb * p.cstride, w0, h0, img,
                                   static_cast<uint16_t>(fs * p.dil), static_cast<uint16_t>(fr * p.dil));
              } else {
                tma_load_2d(&tmA, &full[slot], a_dst, c_base + cb * p.cstride, m0);
              }
            }
            if (!(p.dbg & 2)) tma_load_2d(&tmB, &full[slot], sB + stage * L::B_STAGE, (kb + j) * BLOCK_K, n_tile * BN);
          }
          if (++cb == p.kq[q]) {
            cb = 0;
            if (++q == F3_PARTS) {
              q = 0;
              if (++fs == p.kw) {
                fs = 0;
                ++fr;
              }
            }
          }
        }
        if (++slot == nslots) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer: one TMEM chain per chunk of <= F3_CHUNK_KB k-blocks ====
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BN);
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA)), b_lo0 = smem_desc_lo(smem_u32(sB));
    const int nslots = STAGES / p.ksub;
    int slot = 0;
    uint32_t phase = 0;
    int it = 0;   // chunk counter: TMEM buffer = it & 1
    long long t_first_full = 0, t_unit0_end = 0;
    int units_done = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++units_done) {
      const int s = u % p.S;
      const int kb0 = s * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.num_kblocks);
      if (units_done == 1) t_unit0_end = clock64();
      for (int c0 = kb0; c0 < kb1; c0 += p.chunk_kb, ++it) {
        const int c1 = min(c0 + p.chunk_kb, kb1);
        const int buf = it % F3_NBUF;
        mbar_wait(&tmem_empty[buf], ((it / F3_NBUF) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = c0; kb < c1; kb += p.ksub) {
          const int nsub = min(p.ksub, c1 - kb);
          mbar_wait(&full[slot], phase);
          tc_fence_after();
          if (t_first_full == 0) t_first_full = clock64();
          for (int j = 0; j < nsub; ++j) {
            const int stage = slot * p.ksub + j;
            const uint32_t a_lo = a_lo0 + stage * (A_STAGE_BYTES >> 4);
            const uint32_t b_lo = b_lo0 + stage * (L::B_STAGE >> 4);
            if (elect_one() && !(p.dbg & 4)) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / 16; ++k)
                umma_bf16_lohi(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb + j != c0 || k != 0) ? 1u : 0u);
            }
          }
          if (elect_one()) umma_commit(&empty[slot]);
          if (++slot == nslots) {
            slot = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tmem_full[buf]);
      }
    }
    if ((p.dbg & 16) && blockIdx.x == 0 && lane == 0) {
      const long long t_end = clock64();
      printf("f32x3<%d> grid=%d units=%d (mine %d) kb/unit=%d S=%d | prologue %lld pdl_wait %lld first_full %lld unit0 %lld loop_total %lld clk\n",
             BN, gridDim.x, num_units, units_done, p.kb_per_split < p.num_kblocks ? p.kb_per_split : p.num_kblocks, p.S,
             t_prologue - t_entry, t_pdl - t_prologue, t_first_full - t_pdl, (units_done > 1 ? t_unit0_end : t_end) - t_first_full,
             t_end - t_pdl);
    }
  } else if (warp >= 4) {
    // ===================================== epilogue: chunk sums in registers (RN fp32), then bias / residual / act ====
    constexpr int CW = F3Cfg<BN>::COLS_PER_WARP, EPI_THREADS = F3Cfg<BN>::EPI_WARPS * 32;
    const int q4 = warp & 3;
    const int c_off = ((warp - 4) >> 2) * CW;   // this warp's columns of the tile: [c_off, c_off + CW)
    const int row = q4 * 32 + lane;
    int it = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int tile = u / p.S, s = u - tile * p.S;
      const int m_tile = tile / p.tiles_n, n_tile = tile - m_tile * p.tiles_n;
      const int m = m_tile * BLOCK_M + row;
      const int n0 = n_tile * BN + c_off;
      const int kb0 = s * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.num_kblocks);
      float run[CW];
#pragma unroll
      for (int i = 0; i < CW; ++i) run[i] = 0.f;
      for (int c0 = kb0; c0 < kb1; c0 += p.chunk_kb, ++it) {
        const int buf = it % F3_NBUF;
        mbar_wait(&tmem_full[buf], (it / F3_NBUF) & 1);
        tc_fence_after();
        if (CW == 64) {   // both 32-column loads in flight under one wait
          uint32_t acc0[32], acc1[32];
          const uint32_t ta = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + buf * BN + c_off;
          tmem_ld_32x32(ta, acc0);
          tmem_ld_32x32(ta + 32, acc1);
          tmem_ld_wait_regs(acc0);
          tmem_ld_wait_regs(acc1);
#pragma unroll
          for (int i = 0; i < 32; ++i) run[i] += __uint_as_float(acc0[i]);
#pragma unroll
          for (int i = 0; i < 32; ++i) run[(CW == 64 ? 32 : 0) + i] += __uint_as_float(acc1[i]);
        } else {
          uint32_t acc[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + buf * BN + c_off, acc);
          tmem_ld_wait_regs(acc);
#pragma unroll
          for (int i = 0; i < 32; ++i) run[i] += __uint_as_float(acc[i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      }
      if (p.S > 1) {
        // split-K: publish this split's partial tile; the last split to arrive sums all of them in fixed order
        float* mine = p.partial + (static_cast<size_t>(u) * BLOCK_M + row) * BN + c_off;
#pragma unroll
        for (int i = 0; i < CW; i += 4) *reinterpret_cast<float4*>(mine + i) = make_float4(run[i], run[i + 1], run[i + 2], run[i + 3]);
        __threadfence();
        named_bar_sync(1, EPI_THREADS);
        if (threadIdx.x == 128) {
          const unsigned int prev = atomicAdd(p.counter + tile, 1u);
          *last_flag = (prev == static_cast<unsigned int>(p.S - 1));
          if (prev == static_cast<unsigned int>(p.S - 1)) p.counter[tile] = 0;
          __threadfence();
        }
        named_bar_sync(1, EPI_THREADS);
        const bool last = *last_flag != 0;
        named_bar_sync(1, EPI_THREADS);   // everyone has read the flag before the next unit may rewrite it
        if (!last) continue;
#pragma unroll
        for (int i = 0; i < CW; ++i) run[i] = 0.f;
        for (int ss = 0; ss < p.S; ++ss) {
          const float* src = p.partial + ((static_cast<size_t>(tile) * p.S + ss) * BLOCK_M + row) * BN + c_off;
#pragma unroll
          for (int i = 0; i < CW; i += 4) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(src + i));
            run[i] += v.x; run[i + 1] += v.y; run[i + 2] += v.z; run[i + 3] += v.w;
          }
        }
      }
      if (m < p.M) {
        const int ncol = min(CW, p.Cout - n0);
        float* op = p.out + static_cast<size_t>(m) * p.out_pitch + n0;
        const float* rp = p.has_res ? p.res + static_cast<size_t>(m) * p.res_pitch + n0 : nullptr;
        if (p.vec_ok && (ncol & 3) == 0) {
          // 16-byte loads / stores: a thread owns CW contiguous floats of its output row (scalar 4-byte stores of 32 lanes
          // in 32 different rows were measured at ~30 k clk per tile - more than the tile's whole main loop)
#pragma unroll
          for (int i = 0; i < CW; i += 4) {
            if (i < ncol) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
              float4 v = make_float4(run[i] + b4.x, run[i + 1] + b4.y, run[i + 2] + b4.z, run[i + 3] + b4.w);
              if (rp) {
                const float4 r4 = *reinterpret_cast<const float4*>(rp + i);
                v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
              }
              v.x = f3_act(v.x, p.act, p.act_lo, p.act_hi, p.act_a); v.y = f3_act(v.y, p.act, p.act_lo, p.act_hi, p.act_a);
              v.z = f3_act(v.z, p.act, p.act_lo, p.act_hi, p.act_a); v.w = f3_act(v.w, p.act, p.act_lo, p.act_hi, p.act_a);
              *reinterpret_cast<float4*>(op + i) = v;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < CW; ++i) {
            if (i < ncol) {
              float v = run[i] + __ldg(p.bias + n0 + i);
              if (rp) v += rp[i];
              op[i] = f3_act(v, p.act, p.act_lo, p.act_hi, p.act_a);
            }
          }
        }
      }
    }
  }

  if ((p.dbg & 16) && blockIdx.x == 0 && threadIdx.x == 128)
    printf("f32x3<%d> epilogue done at %lld clk after pdl_wait\n", BN, clock64() - t_pdl);
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------
// pre-kernel: fp32 activation -> [part0 | part1 | part2] bf16 (optionally as explicit im2col rows), counters zeroed
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split3(float v, float& p0, float& p1, float& p2) {
  p0 = __bfloat162float(__float2bfloat16(v));
  const float r1 = v - p0;   // exact: p0 holds the leading 8 bits of v
  p1 = __bfloat162float(__float2bfloat16(r1));
  p2 = __bfloat162float(__float2bfloat16(r1 - p1));
}

__device__ __forceinline__ void split_store8(const float (&v)[8], __nv_bfloat16* dst, size_t part_stride) {
  float p[3][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split3(v[i], p[0][i], p[1][i], p[2][i]);
#pragma unroll
  for (int q = 0; q < F3_PARTS; ++q)
    *reinterpret_cast<uint4*>(dst + q * part_stride) = make_uint4(pack_bf16x2(p[q][0], p[q][1]), pack_bf16x2(p[q][2], p[q][3]),
                                                                  pack_bf16x2(p[q][4], p[q][5]), pack_bf16x2(p[q][6], p[q][7]));
}

__global__ void __launch_bounds__(256)
f3_split_kernel(long long pixels, int C, int in_pitch, const float* __restrict__ x, __nv_bfloat16* __restrict__ xs,
                unsigned int* __restrict__ counter, int ncounters) {
  const long long gid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  pdl_launch_dependents();   // the GEMM kernel's prologue (TMEM allocation, barriers, descriptor prefetch) overlaps this kernel
  if (gid < ncounters) counter[gid] = 0;
  const int cv = C >> 3;
  const long long total = pixels * cv;
  for (long long idx = gid; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = idx / cv;
    const int c8 = static_cast<int>(idx - pix * cv) << 3;
    const float4 a = *reinterpret_cast<const float4*>(x + pix * in_pitch + c8);
    const float4 b = *reinterpret_cast<const float4*>(x + pix * in_pitch + c8 + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    split_store8(v, xs + pix * (F3_PARTS * C) + c8, C);
  }
}

// stems: explicit im2col row per output pixel, K = taps * C channels per part (C <= 16)
__global__ void __launch_bounds__(256)
f3_split_im2col_kernel(int N, int H, int W, int C, int in_pitch, int kh, int kw, int stride, int pad, int dil, int Ho, int Wo,
                       const float* __restrict__ x, __nv_bfloat16* __restrict__ xs, unsigned int* __restrict__ counter,
                       int ncounters) {
  const long long gid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  pdl_launch_dependents();
  if (gid < ncounters) counter[gid] = 0;
  const int cv = C >> 3, taps = kh * kw;
  const int K = taps * C;
  const long long total = static_cast<long long>(N) * Ho * Wo * taps * cv;
  for (long long idx = gid; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = idx;
    const int c8 = static_cast<int>(r % cv) << 3; r /= cv;
    const int tap = static_cast<int>(r % taps); r /= taps;
    const int wo = static_cast<int>(r % Wo); r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    const int fr = tap / kw, fs = tap - fr * kw;
    const int h = ho * stride - pad + fr * dil, w = wo * stride - pad + fs * dil;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (h >= 0 && h < H && w >= 0 && w < W) {
      const float* px = x + ((static_cast<size_t>(n) * H + h) * W + w) * in_pitch + c8;
      const float4 a = *reinterpret_cast<const float4*>(px);
      const float4 b = *reinterpret_cast<const float4*>(px + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    const size_t m = (static_cast<size_t>(n) * Ho + ho) * Wo + wo;
    split_store8(v, xs + m * (F3_PARTS * K) + tap * C + c8, K);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct F3Geom {
  bool im2col;       // explicit im2col rows (stems): the GEMM is a 1x1 conv over K = taps * Cin channels
  int gC, gtaps;     // channels per part / filter taps AS THE GEMM SEES THEM
  int kq[3], per_tap, cstride;
  int Ho, Wo, M, bn, tiles, S, kb_per_split, num_kblocks;
  size_t xs_bytes, partial_bytes, counter_bytes;
};

static F3Geom f3_geom(const pcv_conv_desc& d) {
  F3Geom g;
  const int taps = d.kh * d.kw;
  g.im2col = d.groups == 1 && taps > 1 && d.Cin <= 16;
  g.gC = g.im2col ? taps * d.Cin : d.Cin;
  g.gtaps = g.im2col ? 1 : taps;
  for (int q = 0; q < 3; ++q) g.kq[q] = d.groups > 1 ? F3_PARTS - q : ceil_div((F3_PARTS - q) * g.gC, BLOCK_K);
  g.per_tap = g.kq[0] + g.kq[1] + g.kq[2];
  g.cstride = d.groups > 1 ? d.Cin : BLOCK_K;
  g.Ho = conv_out(d.H, d.kh, d.stride, d.pad, d.dil);
  g.Wo = conv_out(d.W, d.kw, d.stride, d.pad, d.dil);
  g.M = d.N * g.Ho * g.Wo;
  g.num_kblocks = g.gtaps * g.per_tap;
  const int tiles_m = ceil_div(g.M, BLOCK_M);
  // tile width 64 (32 for narrow layers); small-M layers get their parallelism from split-K below.  A 128-wide variant
  // (8 epilogue warps, 6 stages) exists and was measured SLOWER (ResNet-18 bs8: 2.05 vs 1.52 ms/step): the kernel is bound
  // by k-block round trips per SM, and fewer, larger units with a shallower ring lose more than the halved A traffic wins.
  // PCV_F3_BN=128 selects it for experiments.
  const int sms = sm_count();
  g.bn = (d.groups > 1 || d.Cout > 32) ? 64 : 32;
  if (const char* e = getenv("PCV_F3_BN")) {
    const int v = atoi(e);
    if (d.groups == 1 && (v == 32 || v == 64 || (v == 128 && d.Cout >= 128))) g.bn = v;
  }
  g.tiles = tiles_m * ceil_div(d.Cout, g.bn);
  // split-K when the tiles alone cannot fill the machine; at least 2 chunks of work per split
  int S = 1;
  if (g.tiles < sms) S = std::min(ceil_div(3 * sms, 2 * g.tiles), std::max(1, g.num_kblocks / (2 * F3_CHUNK_KB)));
  S = std::max(1, std::min(S, 32));
  g.kb_per_split = ceil_div(g.num_kblocks, S);
  if (S > 1) g.kb_per_split = round_up(g.kb_per_split, F3_CHUNK_KB);   // splits start on chunk (and slot) boundaries
  g.S = ceil_div(g.num_kblocks, g.kb_per_split);
  const size_t rows = g.im2col ? static_cast<size_t>(g.M) : static_cast<size_t>(d.N) * d.H * d.W;
  g.xs_bytes = (rows * F3_PARTS * g.gC * 2 + 255) & ~static_cast<size_t>(255);
  g.partial_bytes = g.S > 1 ? static_cast<size_t>(g.tiles) * g.S * BLOCK_M * g.bn * 4 : 0;
  g.counter_bytes = (static_cast<size_t>(g.tiles) * 4 + 255) & ~static_cast<size_t>(255);
  return g;
}

__global__ void f3_pack_kernel(const float* __restrict__ w, const float* __restrict__ conv_bias, const float* __restrict__ g,
                               const float* __restrict__ b, const float* __restrict__ mean, const float* __restrict__ var,
                               float eps, int Cout, int Cin, int groups, int taps, int im2col, int gC, int gtaps, int kq0,
                               int kq1, int kq2, int in_span, __nv_bfloat16* __restrict__ wp, float* __restrict__ bias_out,
                               int bias_len) {
  const int per_tap = (kq0 + kq1 + kq2) * BLOCK_K;
  const int kpad = gtaps * per_tap;
  const size_t total = static_cast<size_t>(Cout) * kpad;
  const int cin_g = Cin / groups, cout_g = Cout / groups;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(idx / kpad);
    const int kp = static_cast<int>(idx - static_cast<size_t>(o) * kpad);
    const int gtap = kp / per_tap;
    const int r = kp - gtap * per_tap;
    const int j = r / BLOCK_K, within = r - j * BLOCK_K;
    const int q = j < kq0 ? 0 : (j < kq0 + kq1 ? 1 : 2);
    const int wdw = j - (q == 0 ? 0 : (q == 1 ? kq0 : kq0 + kq1));
    const float scale = g ? g[o] * rsqrtf(var[o] + eps) : 1.f;
    float val = 0.f;
    if (groups == 1) {
      const int c_abs = wdw * BLOCK_K + within;   // channel of the split tensor: part a, GEMM channel cg
      const int a = c_abs / gC;
      const int cg = c_abs - a * gC;
      if (a + q < F3_PARTS) {
        const int tap = im2col ? cg / Cin : gtap;
        const int ci = im2col ? cg - tap * Cin : cg;
        val = w[(static_cast<size_t>(o) * Cin + ci) * taps + tap] * scale;
      }
    } else {
      // window = activation part `wdw` of this output tile's 64-channel group window (block-diagonal weights)
      const int ci = (o / 64) * in_span + within - (o / cout_g) * cin_g;
      if (ci >= 0 && ci < cin_g && within < in_span) val = w[(static_cast<size_t>(o) * cin_g + ci) * taps + gtap] * scale;
    }
    float p0, p1, p2;
    split3(val, p0, p1, p2);
    wp[idx] = __float2bfloat16(q == 0 ? p0 : (q == 1 ? p1 : p2));
  }
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < bias_len; o += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (o < Cout) {
      const float cb = conv_bias ? conv_bias[o] : 0.f;
      v = g ? (cb - mean[o]) * (g[o] * rsqrtf(var[o] + eps)) + b[o] : cb;
    }
    bias_out[o] = v;
  }
}

int igemm_split_supported(const pcv_conv_desc& d, std::string* why) {
  if (!igemm_supported(d, why)) return 0;
  auto no = [&](const char* m) {
    if (why) *why = m;
    return 0;
  };
  if ((d.flags & (PCV_CONV_IN_OVERLAP | PCV_CONV_POOL3S2)) || d.in_row_pitch != 0)
    return no("row-pitched / overlapping input views have no fp32 split form");
  if (pitch_or(d.in_pitch, d.Cin) % 4 != 0) return no("fp32 input pitch must be a multiple of 4");
  if (static_cast<long long>(d.N) * d.H * d.W * F3_PARTS * d.Cin * d.kh * d.kw >= (1ll << 40)) return no("workspace too large");
  return 1;
}

int igemm_split_packed_bytes(const pcv_conv_desc& d, size_t* w_bytes, size_t* b_bytes, size_t* ws_bytes) {
  const F3Geom g = f3_geom(d);
  if (w_bytes) *w_bytes = static_cast<size_t>(d.Cout) * g.num_kblocks * BLOCK_K * 2;
  if (b_bytes) *b_bytes = static_cast<size_t>(round_up(d.Cout, 256)) * 4;
  if (ws_bytes) *ws_bytes = g.xs_bytes + g.partial_bytes + g.counter_bytes;
  return PCV_OK;
}

int igemm_split_pack(const pcv_conv_desc& d, const float* w, const float* conv_bias, const float* g, const float* b,
                     const float* m, const float* v, float eps, void* w_packed, float* bias_out, cudaStream_t s) {
  const F3Geom sg = f3_geom(d);
  const size_t total = static_cast<size_t>(d.Cout) * sg.num_kblocks * BLOCK_K;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 4096));
  f3_pack_kernel<<<blocks, 256, 0, s>>>(w, conv_bias, g, b, m, v, eps, d.Cout, d.Cin, d.groups, d.kh * d.kw, sg.im2col ? 1 : 0,
                                        sg.gC, sg.gtaps, sg.kq[0], sg.kq[1], sg.kq[2],
                                        d.groups > 1 ? 64 * (d.Cin / d.groups) / (d.Cout / d.groups) : 64,
                                        reinterpret_cast<__nv_bfloat16*>(w_packed), bias_out, round_up(d.Cout, 256));
  g_launches++;
  PCV_CHECK_CUDA(cudaGetLastError());
  return PCV_OK;
}

struct F3Op : Op {
  CUtensorMap tmA, tmB;
  F3Params p;
  pcv_conv_desc d;
  F3Geom g;
  const float* x;
  __nv_bfloat16* xs;
  int in_pitch, grid;
  template <int BN>
  cudaError_t run(cudaStream_t s) {
    static std::atomic<uint64_t> attr_done{0};
    if (cudaError_t e = set_max_smem_once(f32x3_kernel<BN>, F3Smem<BN>::BYTES, attr_done)) return e;
    return launch_pdl(f32x3_kernel<BN>, dim3(grid), dim3(F3Cfg<BN>::THREADS), F3Smem<BN>::BYTES, s, tmA, tmB, p);
  }
  cudaError_t launch(cudaStream_t s) override {
    g_launches += 2;
    const int cap = sm_count() * 16;
    if (g.im2col) {
      const long long items = static_cast<long long>(g.M) * d.kh * d.kw * (d.Cin >> 3);
      f3_split_im2col_kernel<<<static_cast<int>(std::min<long long>((items + 255) / 256, cap)), 256, 0, s>>>(
          d.N, d.H, d.W, d.Cin, in_pitch, d.kh, d.kw, d.stride, d.pad, d.dil, g.Ho, g.Wo, x, xs, p.counter, g.tiles);
    } else {
      const long long pixels = static_cast<long long>(d.N) * d.H * d.W;
      const long long items = std::max<long long>(pixels * (d.Cin >> 3), g.tiles);
      f3_split_kernel<<<static_cast<int>(std::min<long long>((items + 255) / 256, cap)), 256, 0, s>>>(
          pixels, d.Cin, in_pitch, x, xs, p.counter, g.tiles);
    }
    if (cudaError_t e = cudaGetLastError()) return e;
    return g.bn == 32 ? run<32>(s) : (g.bn == 64 ? run<64>(s) : run<128>(s));
  }
};

int igemm_split_make(const pcv_conv_desc& d, const void* x, const void* w, const float* bias, const void* res, void* y,
                     void* workspace, Op** out) {
  std::string why;
  if (!igemm_split_supported(d, &why)) return fail(PCV_ERR_UNSUPPORTED, "fp32 split conv: %s", why.c_str());
  PCV_REQUIRE(workspace != nullptr, "PCV_CONV_F32_SPLIT needs a workspace of pcv_conv_workspace_bytes() bytes");
  const F3Geom g = f3_geom(d);
  PCV_REQUIRE(g.Ho > 0 && g.Wo > 0, "conv output is empty (H=%d W=%d k=%d)", d.H, d.W, d.kh);
  PCV_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(w) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(workspace) % 256 == 0,
              "conv operands must be 16-byte (workspace: 256-byte) aligned");
  PCV_REQUIRE(g.tiles * g.S < (1 << 30), "too many work units");
  auto op = std::make_unique<F3Op>();
  op->d = d;
  op->g = g;
  op->x = reinterpret_cast<const float*>(x);
  op->xs = reinterpret_cast<__nv_bfloat16*>(workspace);
  op->in_pitch = pitch_or(d.in_pitch, d.Cin);
  op->launches = 2;
  F3Params& p = op->p;
  p.bias = bias;
  p.out = reinterpret_cast<float*>(y);
  p.res = reinterpret_cast<const float*>(res);
  p.partial = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + g.xs_bytes);
  p.counter = reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(workspace) + g.xs_bytes + g.partial_bytes);
  p.M = g.M;
  p.Cout = d.Cout;
  p.out_pitch = pitch_or(d.out_pitch, d.Cout);
  p.res_pitch = pitch_or(d.res_pitch, d.Cout);
  p.HoWo = g.Ho * g.Wo;
  p.Wo = g.Wo;
  p.stride = d.stride; p.pad = d.pad; p.dil = d.dil; p.kw = g.im2col ? 1 : d.kw;
  p.num_kblocks = g.num_kblocks;
  p.tiles_m = ceil_div(g.M, BLOCK_M);
  p.tiles_n = ceil_div(d.Cout, g.bn);
  p.act = d.act;
  p.act_lo = (d.act == PCV_ACT_RELU || d.act == PCV_ACT_RELU6) ? 0.f : -INFINITY;
  p.act_hi = d.act == PCV_ACT_RELU6 ? 6.f : INFINITY;
  p.act_a = d.act_param;
  p.has_res = res != nullptr;
  p.grouped = d.groups > 1;
  p.g_in_span = p.grouped ? 64 * (d.Cin / d.groups) / (d.Cout / d.groups) : 0;
  for (int q = 0; q < 3; ++q) p.kq[q] = g.kq[q];
  p.per_tap = g.per_tap;
  p.cstride = g.cstride;
  p.S = g.S;
  p.kb_per_split = g.kb_per_split;
  p.chunk_kb = F3_CHUNK_KB;
  if (const char* e = getenv("PCV_F3_CHUNK")) p.chunk_kb = std::max(1, atoi(e));
  p.dbg = 0;
  if (const char* e = getenv("PCV_F3_DBG")) p.dbg = atoi(e);
  p.vec_ok = (p.out_pitch % 4 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0) &&
             (!res || (p.res_pitch % 4 == 0 && reinterpret_cast<uintptr_t>(res) % 16 == 0));
  p.ksub = 4;   // measured on ResNet-18 bs8: 1 -> 1.143, 2 -> 1.048, 4 -> 1.031 ms/step
  if (const char* e = getenv("PCV_F3_KSUB")) p.ksub = std::max(1, atoi(e));
  if (p.chunk_kb % p.ksub != 0 || (p.S > 1 && p.kb_per_split % p.ksub != 0)) p.ksub = 1;   // slots never straddle a chunk or a split
  const bool pointwise = g.im2col || (d.kh * d.kw == 1 && d.stride == 1 && d.pad == 0);
  p.a_mode = pointwise ? 0 : 1;
  // the activation the GEMM reads: the workspace, [rows, 3 * gC] bf16
  const int C3 = F3_PARTS * g.gC;
  int rc;
  if (p.a_mode == 1) {
    pcv_conv_desc ds = d;
    ds.Cin = C3;
    ds.in_pitch = 0;
    rc = make_im2col_4d(&op->tmA, workspace, ds, C3);
  } else {
    rc = make_tiled_2d(&op->tmA, workspace, C3, g.M, (uint64_t)C3 * 2, BLOCK_K, BLOCK_M, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (rc) return rc;
  const uint64_t kpad = (uint64_t)g.num_kblocks * BLOCK_K;
  rc = make_tiled_2d(&op->tmB, w, kpad, d.Cout, kpad * 2, BLOCK_K, g.bn, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  op->grid = std::min(g.tiles * g.S, sm_count());
  char nm[176];
  snprintf(nm, sizeof nm, "conv_tc_f32x3 %dx%d s%d d%d g%d %d->%d @%dx%d bn=%d kb=%d%s%s%s", d.kh, d.kw, d.stride, d.dil,
           d.groups, d.Cin, d.Cout, d.H, d.W, g.bn, g.num_kblocks, g.im2col ? " im2col" : "",
           g.S > 1 ? (" splitK=" + std::to_string(g.S)).c_str() : "", res ? " +res" : "");
  op->name = nm;
  const int taps = d.kh * d.kw;
  const double pin = (taps == 1 && d.stride > 1) ? (double)g.Ho * g.Wo : (double)d.H * d.W;
  op->flops = 2.0 * g.M * d.Cout * (d.Cin / d.groups) * taps;
  op->bytes = 4.0 * d.N * d.Cin * pin + 4.0 * g.M * d.Cout * (res ? 2.0 : 1.0) + 4.0 * d.Cout * (d.Cin / d.groups) * taps +
              4.0 * d.Cout;
  *out = op.release();
  return PCV_OK;
}

}  // namespace PCV_TIER
}  // namespace pcv
