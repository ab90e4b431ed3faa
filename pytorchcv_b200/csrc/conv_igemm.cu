// Dense / grouped convolution as a tcgen05 implicit GEMM (bf16 in, fp32 accumulate in TMEM).
//
// Replaces, for one ConvBlock, the reference's  nn.Conv2d -> BatchNorm2d -> activation  sequence
// (pytorchcv/models/common/conv.py:278-286) plus the unit-level  x + identity ; ReLU  (models/resnet.py:221-229):
// BatchNorm is folded into the packed weights / bias, residual add and activation run in the epilogue.
//
// GEMM view:  D[m, n] = sum_k A[m, k] * B[n, k]
//   m = output pixel (n_img, ho, wo) linearised, M = N*Ho*Wo        (BLOCK_M = 128 rows = 128 TMEM lanes)
//   n = output channel                                                (BLOCK_N = 32 / 64 / 128 TMEM columns)
//   k = (filter tap, input channel), 64 channels per k-block          (BLOCK_K = 64 bf16 = one 128-byte swizzle row)
// A tiles come straight from the NHWC activation tensor through a TMA *im2col* descriptor (one instruction per
// k-block: 128 pixels x 64 channels, halo and padding zero-filled by the TMA unit); 1x1 stride-1 convs use a plain
// 2-D tiled descriptor.  B tiles come from the packed [Cout, taps*Cpad] weight matrix.  Both land in 128B-swizzled
// shared memory and feed tcgen05.mma (M=128, N=BLOCK_N, K=16) issued by a single thread.
//
// Persistent, warp-specialised CTA (one per SM), 8 warps:
//   warp 0  TMA producer (A+B ring of STAGES)          warp 1  MMA issuer (double-buffered TMEM accumulator)
//   warp 2  TMEM allocator                              warp 3  residual prefetcher (TMA load into the staging tile)
//   warps 4-7  epilogue: tcgen05.ld -> +bias (+residual) -> act -> bf16 -> swizzled staging smem -> TMA store
// so the epilogue of tile i overlaps the main loop of tile i+1.
#include "igemm_common.cuh"

namespace pcv {
namespace PCV_TIER {

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int B_STAGE_BYTES = BN * BLOCK_K * 2;
  static constexpr int SUB_COLS = BN >= 64 ? 64 : BN;               // columns per staging sub-tile
  static constexpr int SUB_BYTES = BLOCK_M * SUB_COLS * 2;          // one [128 x SUB_COLS] bf16 sub-tile
  static constexpr int NSUB = BN / SUB_COLS;
  static constexpr int STG_BYTES = BLOCK_M * BN * 2;                // one staging buffer
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + STAGES * A_STAGE_BYTES;
  static constexpr int OFF_STG = OFF_B + STAGES * B_STAGE_BYTES;
  static constexpr int NSTG = 3;                                    // staging ring: residual in -> result out
  static constexpr int OFF_BAR = OFF_STG + NSTG * STG_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 4 + 2 * NSTG;
  static constexpr int TOTAL = OFF_BAR + NUM_BARS * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;                    // slack for manual 1 KiB alignment
};

template <int BN, int STAGES, int OUT_MODE>
__global__ void __launch_bounds__(NUM_THREADS, 2)   // <= 128 registers: the BN = 32 / 3-stage variant runs two CTAs per SM
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
             const IgemmParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint8_t* sA = smem + L::OFF_A;
  uint8_t* sB = smem + L::OFF_B;
  uint8_t* sStg = smem + L::OFF_STG;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* full = bars;                       // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;             // [STAGES]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * STAGES;     // [2]       MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;        // [2]       epilogue -> MMA
  uint64_t* stg_empty = tmem_empty + 2;        // [NSTG]    epilogue -> residual prefetcher
  uint64_t* res_full = stg_empty + L::NSTG;    // [NSTG]    residual TMA -> epilogue
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

  // role index (see conv_igemm2.cu): control roles live in the high physical warp ids, which the scheduler prefers
  const int warp = ((threadIdx.x >> 5) + 4) & 7;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (OUT_MODE == 0) tma_prefetch_desc(&tmOut);
    if (p.has_res && OUT_MODE == 0) tma_prefetch_desc(&tmRes);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    for (int i = 0; i < L::NSTG; ++i) {
      mbar_init(&stg_empty[i], 1);
      mbar_init(&res_full[i], 1);
    }
    fence_mbar_init();
    // staging buffers 0..NSTG-2 start out free; buffer NSTG-1 is released by tile 0's epilogue (see below)
    for (int i = 0; i < L::NSTG - 1; ++i) mbar_arrive(&stg_empty[i]);
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;   // non-zero for the second CTA of an SM (two-CTAs-per-SM variant)
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // whole warp with warp-uniform values; TMA / mbarrier instructions under elect.sync so ptxas keeps addresses in
    // uniform registers (inside `if (lane == 0)` every UTMALDG / UTCHMMA got a ~13-instruction uniformisation loop)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_tile = t / p.tiles_n;
        const int n_tile = t - m_tile * p.tiles_n;
        const int m0 = m_tile * BLOCK_M;
        const int img = m0 / p.HoWo;
        const int rem = m0 - img * p.HoWo;
        const int ho = rem / p.Wo;
        const int wo = rem - ho * p.Wo;
        const int w0 = wo * p.stride - p.pad;
        const int h0 = ho * p.stride - p.pad;
        const int c_base = p.grouped ? n_tile * p.g_in_span : 0;
        int tap = 0, cb = 0, fr = 0, fs = 0;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[stage], A_STAGE_BYTES + L::B_STAGE_BYTES);
            uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
            if (p.a_mode == 1) {
              tma_load_im2col_4d(&tmA, &full[stage], a_dst, c_base + cb * BLOCK_K, w0, h0, img,
                                 static_cast<uint16_t>(fs * p.dil), static_cast<uint16_t>(fr * p.dil));
            } else {
              tma_load_2d(&tmA, &full[stage], a_dst, c_base + cb * BLOCK_K, m0);
            }
            tma_load_2d(&tmB, &full[stage], sB + stage * L::B_STAGE_BYTES, kb * BLOCK_K, n_tile * BN);
          }
          if (++cb == p.cblocks) {
            cb = 0;
            ++tap;
            if (++fs == p.kw) {
              fs = 0;
              ++fr;
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
    // ===================================== MMA issuer =====================================
    {   // whole warp; tcgen05 instructions under elect.sync
      constexpr uint32_t idesc = make_idesc_e16(BLOCK_M, BN);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA)), b_lo0 = smem_desc_lo(smem_u32(sB));
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[buf], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (A_STAGE_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + stage * (L::B_STAGE_BYTES >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k)
              umma_bf16_lohi(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty[stage]);  // frees this smem stage once the MMAs above have read it
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tmem_full[buf]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp == 3) {
    // ===================================== residual prefetcher =====================================
    if (p.has_res && OUT_MODE == 0) {   // whole warp, TMA under elect.sync
      int sbuf = 0;
      uint32_t sphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_tile = t / p.tiles_n;
        const int n_tile = t - m_tile * p.tiles_n;
        mbar_wait(&stg_empty[sbuf], sphase);
        if (elect_one()) {
          mbar_arrive_expect_tx(&res_full[sbuf], L::STG_BYTES);
#pragma unroll
          for (int sub = 0; sub < L::NSUB; ++sub)
            tma_load_2d(&tmRes, &res_full[sbuf], sStg + sbuf * L::STG_BYTES + sub * L::SUB_BYTES,
                        n_tile * BN + sub * L::SUB_COLS, m_tile * BLOCK_M);
        }
        if (++sbuf == L::NSTG) {
          sbuf = 0;
          sphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =====================================
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;       // accumulator row == output pixel within the tile
    const int epi_tid = threadIdx.x - 128;
    constexpr uint32_t ROW_BYTES = L::SUB_COLS * 2;
    constexpr uint32_t SWZ_MASK = ROW_BYTES == 128 ? 7u : (ROW_BYTES == 64 ? 3u : 1u);
    const float act_lo = p.act_lo, act_hi = p.act_hi;
    const bool fancy_act = p.act > PCV_ACT_RELU6;
    const bool relu = p.act_lo == 0.f, capped = p.act_hi != INFINITY;   // the clamp family: none / ReLU / ReLU6
    const uint32_t cap2 = pack_e16x2(p.act_hi, p.act_hi);
    const uint32_t sStg_u32 = smem_u32(sStg);
    int it = 0;
    int sbuf = 0;           // staging ring position of this tile
    uint32_t sphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m_tile = t / p.tiles_n;
      const int n_tile = t - m_tile * p.tiles_n;
      const int m0 = m_tile * BLOCK_M;
      const int n0 = n_tile * BN;
      uint8_t* stg = sStg + sbuf * L::STG_BYTES;
      const uint32_t stg_u32 = sStg_u32 + sbuf * L::STG_BYTES;

      if (p.has_res && OUT_MODE == 0) mbar_wait(&res_full[sbuf], sphase);
      mbar_wait(&tmem_full[buf], acc_phase);
      tc_fence_after();

#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + j * 32, acc);
        // bias (and, staged mode, residual) loads issued under the TMEM load: the wait is a compiler barrier for memory ops
        const float4* bias4 = reinterpret_cast<const float4*>(p.bias + n0 + j * 32);
        float4 b4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)   // columns past Cout are clipped by the store: no read there
          b4[i] = (n0 + j * 32 + 4 * i < p.Cout) ? __ldg(bias4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int col = j * 32;
        const uint32_t sub_u32 = stg_u32 + (col / L::SUB_COLS) * L::SUB_BYTES;
        const uint32_t row_off = row * ROW_BYTES + (col % L::SUB_COLS) * 2;
        uint4 r4[4];
        if (OUT_MODE == 0 && p.has_res) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t off = row_off + c * 16;
            off ^= ((off >> 7) & SWZ_MASK) << 4;
            r4[c] = lds128(sub_u32 + off);
          }
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
        if (OUT_MODE == 0) {
          // staging sub-tile holding columns [j*32, j*32+32): row pitch ROW_BYTES, 16-byte chunks XOR-swizzled
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
          for (int c = 0; c < 4; ++c) {
            uint32_t off = row_off + c * 16;
            off ^= ((off >> 7) & SWZ_MASK) << 4;
            sts128(sub_u32 + off, o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
          }
        } else {
          // direct global stores: any Cout / pitch, bf16 or fp32 output (classifier logits, 21-class heads)
          const int m = m0 + row;
          if (m < p.M) {
            const int ncol = min(32, p.Cout - (n0 + j * 32));
            if (p.has_res) {
              const e16* rp = p.res + static_cast<size_t>(m) * p.res_pitch + n0 + j * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < ncol) v[i] += e16_to_float(rp[i]);
            }
            if (fancy_act) {
              fast_act_n(v, p.act, p.act_a);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fminf(fmaxf(v[i], act_lo), act_hi);
            }
            if (OUT_MODE == 2) {
              float* op = reinterpret_cast<float*>(p.out) + static_cast<size_t>(m) * p.out_pitch + n0 + j * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < ncol) op[i] = v[i];
            } else {
              e16* op =
                  reinterpret_cast<e16*>(p.out) + static_cast<size_t>(m) * p.out_pitch + n0 + j * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < ncol) op[i] = float_to_e16(v[i]);
            }
          }
        }
      }
      // accumulator buffer fully read -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);

      if (OUT_MODE == 0) {
        fence_proxy_async_smem();           // st.shared above -> visible to the TMA (async proxy)
        named_bar_sync(1, EPI_THREADS);
        if (warp == 4 && elect_one()) {   // elect.sync is deterministic: one lane owns every bulk group
#pragma unroll
          for (int sub = 0; sub < L::NSUB; ++sub)
            tma_store_2d(&tmOut, stg + sub * L::SUB_BYTES, n0 + sub * L::SUB_COLS, m0);
          tma_store_commit();
          // the store issued one tile ago has finished reading its buffer: hand that buffer (ring position
          // sbuf-1, i.e. the one tile it+NSTG-1 will use) back to the residual prefetcher
          tma_store_wait_read<1>();
          mbar_arrive(&stg_empty[sbuf == 0 ? L::NSTG - 1 : sbuf - 1]);
        }
        named_bar_sync(1, EPI_THREADS);     // ... and every epilogue thread may overwrite it from now on
      }
      if (++sbuf == L::NSTG) {
        sbuf = 0;
        sphase ^= 1;
      }
    }
    if (OUT_MODE == 0 && warp == 4 && elect_one()) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------
// weight packing: fold BN, cast to bf16, lay out as [Cout, taps * cblocks * 64] (K-major, zero padded)
// ------------------------------------------------------------------------------------------------------------
__global__ void igemm_pack_kernel(const float* __restrict__ w, const float* __restrict__ conv_bias,
                                  const float* __restrict__ g, const float* __restrict__ b,
                                  const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                  int Cout, int Cin, int groups, int taps, int cblocks, int grouped_bn, int in_span,
                                  e16* __restrict__ wp, float* __restrict__ bias_out, int bias_len) {
  const int kpad = taps * cblocks * BLOCK_K;
  const size_t total = static_cast<size_t>(Cout) * kpad;
  const int cin_g = Cin / groups;
  const int cout_g = Cout / groups;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(idx / kpad);
    const int kp = static_cast<int>(idx - static_cast<size_t>(o) * kpad);
    const int tap = kp / (cblocks * BLOCK_K);
    const int cc = kp - tap * (cblocks * BLOCK_K);
    const float scale = g ? g[o] * rsqrtf(var[o] + eps) : 1.f;
    float val = 0.f;
    if (groups == 1) {
      if (cc < Cin) val = w[(static_cast<size_t>(o) * Cin + cc) * taps + tap];
    } else {
      // block-diagonal: the A window of this output channel's N tile starts at input channel (o / bn) * in_span, where
      // in_span = bn * cin_g / cout_g input channels feed the tile's bn / cout_g groups (== bn when Cin/g == Cout/g;
      // SENet's half-width grouped 3x3, senet.py:52-56, has in_span = 32)
      const int ci_abs = (o / grouped_bn) * in_span + cc;
      const int grp = o / cout_g;
      const int ci = ci_abs - grp * cin_g;
      if (ci >= 0 && ci < cin_g && cc < in_span) val = w[(static_cast<size_t>(o) * cin_g + ci) * taps + tap];
    }
    wp[idx] = float_to_e16(val * scale);
  }
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < bias_len; o += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (o < Cout) {
      const float cb = conv_bias ? conv_bias[o] : 0.f;
      if (g) {
        const float scale = g[o] * rsqrtf(var[o] + eps);
        v = (cb - mean[o]) * scale + b[o];
      } else {
        v = cb;
      }
    }
    bias_out[o] = v;
  }
}

int pick_bn(const pcv_conv_desc& d, int tiles_m) {
  if (d.groups > 1) return 64;
  if (d.Cout <= 32) return 32;
  if (d.Cout <= 64) return 64;
  // few output rows (the classifier on pooled features: M = batch): narrower tiles put more CTAs on the weight stream -
  // 2048 -> 1000 at batch 256 was 16 CTAs walking K = 2048 alone (0.030 ms against a 0.001 ms bound)
  const int sms = sm_count();
  if (tiles_m * ceil_div(d.Cout, 128) * 2 <= sms) return tiles_m * ceil_div(d.Cout, 64) * 2 <= sms ? 32 : 64;
  return 128;
}

int igemm_supported(const pcv_conv_desc& d, std::string* why) {
  auto no = [&](const char* m) {
    if (why) *why = m;
    return 0;
  };
  const int in_pitch = pitch_or(d.in_pitch, d.Cin);
  if (in_pitch % 8 != 0 || d.in_row_pitch % 8 != 0)
    return no("input channel / row pitch must be a multiple of 8 (16-byte TMA stride)");
  if (d.Cin % 8 != 0) return no("Cin must be a multiple of 8");
  if (d.kh * d.kw > 49 || d.stride > 8) return no("kernel/stride out of range");
  const int lo = -d.pad, up_w = d.pad - (d.kw - 1) * d.dil, up_h = d.pad - (d.kh - 1) * d.dil;
  if (lo < -128 || up_w < -128 || up_h < -128 || up_w > 127 || up_h > 127) return no("im2col corner out of range");
  if (d.groups > 1) {
    const int cg_in = d.Cin / d.groups, cg_out = d.Cout / d.groups;
    if (64 % cg_out != 0 || d.Cout % 64 != 0) return no("grouped conv needs 64 % (Cout/g) == 0 and Cout % 64 == 0");
    // a 64-wide output tile covers 64 / cg_out groups = in_span input channels, read as (part of) one 64-channel k-block
    const int in_span = 64 * cg_in / cg_out;
    if (cg_in > cg_out || in_span * cg_out != 64 * cg_in || in_span % 8 != 0)
      return no("grouped conv needs Cin/g <= Cout/g with 64*(Cin/g)/(Cout/g) a multiple of 8");
  }
  return 1;
}

int igemm_packed_bytes(const pcv_conv_desc& d, size_t* w_bytes, size_t* b_bytes) {
  const int taps = d.kh * d.kw;
  const int cblocks = d.groups > 1 ? 1 : ceil_div(d.Cin, BLOCK_K);
  *w_bytes = static_cast<size_t>(d.Cout) * taps * cblocks * BLOCK_K * 2;
  *b_bytes = static_cast<size_t>(round_up(d.Cout, 256)) * 4;
  return PCV_OK;
}

int igemm_pack(const pcv_conv_desc& d, const float* w, const float* conv_bias, const float* g, const float* b,
               const float* m, const float* v, float eps, void* w_packed, float* bias_out, cudaStream_t s) {
  const int taps = d.kh * d.kw;
  const int cblocks = d.groups > 1 ? 1 : ceil_div(d.Cin, BLOCK_K);
  const size_t total = static_cast<size_t>(d.Cout) * taps * cblocks * BLOCK_K;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 4096));
  igemm_pack_kernel<<<blocks, 256, 0, s>>>(w, conv_bias, g, b, m, v, eps, d.Cout, d.Cin, d.groups, taps, cblocks, 64,
                                           d.groups > 1 ? 64 * (d.Cin / d.groups) / (d.Cout / d.groups) : 64,
                                           reinterpret_cast<e16*>(w_packed), bias_out,
                                           round_up(d.Cout, 256));
  g_launches++;
  PCV_CHECK_CUDA(cudaGetLastError());
  return PCV_OK;
}

// ------------------------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------------------------
int make_tiled_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                         uint32_t box_inner, uint32_t box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, TMAP_E16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu stride=%llu box=%ux%u", (int)r,
                (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_stride_bytes, box_inner,
                box_rows);
  return PCV_OK;
}

int make_im2col_4d(CUtensorMap* tm, const void* base, const pcv_conv_desc& d, int in_pitch) {
  EncodeIm2colFn fn = encode_im2col_fn();
  if (!fn) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeIm2col entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)d.Cin, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.N};
  const cuuint64_t row_pitch = d.in_row_pitch > 0 ? (cuuint64_t)d.in_row_pitch : (cuuint64_t)d.W * in_pitch;
  cuuint64_t strides[3] = {(cuuint64_t)in_pitch * 2, row_pitch * 2, (cuuint64_t)d.H * row_pitch * 2};
  int lower[2] = {-d.pad, -d.pad};                                              // {W, H}
  int upper[2] = {d.pad - (d.kw - 1) * d.dil, d.pad - (d.kh - 1) * d.dil};      // {W, H}
  cuuint32_t estr[4] = {1, (cuuint32_t)d.stride, (cuuint32_t)d.stride, 1};
  CUresult r = fn(tm, TMAP_E16, 4, const_cast<void*>(base), dims, strides, lower, upper,
                  BLOCK_K, BLOCK_M, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(PCV_ERR_CUDA, "cuTensorMapEncodeIm2col failed (%d): C=%d W=%d H=%d N=%d pitch=%d pad=%d k=%d dil=%d s=%d",
                (int)r, d.Cin, d.W, d.H, d.N, in_pitch, d.pad, d.kw, d.dil, d.stride);
  return PCV_OK;
}

struct IgemmOp : Op {
  CUtensorMap tmA, tmB, tmOut, tmRes;
  IgemmParams p;
  int bn, grid;
  bool pair = false;  // cta_group::2 kernel (conv_igemm2.cu)
  bool twin = false;  // BN = 32, short K: 3-stage variant, two CTAs per SM (64 TMEM columns and ~86 KB of smem each)
  cudaError_t launch(cudaStream_t s) override;
};

template <int BN, int STAGES, int OUT_MODE>
static cudaError_t launch_variant(const IgemmOp& op, cudaStream_t s) {
  using L = SmemLayout<BN, STAGES>;
  static std::atomic<uint64_t> attr_done{0};   // per device (see runtime.h)
  if (cudaError_t e = set_max_smem_once(igemm_kernel<BN, STAGES, OUT_MODE>, L::DYN_BYTES, attr_done)) return e;
  return launch_pdl(igemm_kernel<BN, STAGES, OUT_MODE>, dim3(op.grid), dim3(NUM_THREADS), L::DYN_BYTES, s, op.tmA, op.tmB,
                    op.tmOut, op.tmRes, op.p);
}

template <int BN, int STAGES>
static cudaError_t launch_bn(const IgemmOp& op, cudaStream_t s) {
  switch (op.p.out_mode) {
    case 0: return launch_variant<BN, STAGES, 0>(op, s);
    case 1: return launch_variant<BN, STAGES, 1>(op, s);
    default: return launch_variant<BN, STAGES, 2>(op, s);
  }
}

cudaError_t IgemmOp::launch(cudaStream_t s) {
  g_launches++;
  if (pair) return launch_igemm2(bn, grid, tmA, tmB, tmOut, tmRes, p, s);
  switch (bn) {
    case 32: return twin ? launch_bn<32, 3>(*this, s) : launch_bn<32, 6>(*this, s);
    case 64: return launch_bn<64, 6>(*this, s);
    default: return launch_bn<128, 4>(*this, s);
  }
}

// Is `d` served by the CTA-pair kernel (the one with the PCV_CONV_SE_GATE epilogue)?  Mirrors igemm_make's choice for aligned
// operands: a dense 1x1 / k x k layer outside the halo and stem kernels, bf16/fp16 output through TMA stores.
int igemm_gate_ok(const pcv_conv_desc& d) {
  std::string why;
  if (!igemm_supported(d, &why) || d.groups != 1 || (d.flags & (PCV_CONV_OUT_F32 | PCV_CONV_IN_OVERLAP | PCV_CONV_POOL3S2))) return 0;
  if (d.kh != 1 || d.kw != 1 || d.stride != 1 || d.pad != 0) return 0;   // the unit's last 1x1 conv
  const long long M = static_cast<long long>(d.N) * d.H * d.W;
  if (pitch_or(d.out_pitch, d.Cout) % 8 || pitch_or(d.res_pitch, d.Cout) % 8 || d.Cout % 8 || d.Cout < 64) return 0;
  const char* e = getenv("PCV_IGEMM_2CTA");
  return !(e && e[0] == '0') && ceil_div(static_cast<int>(M), BLOCK_M) >= 2;
}

// Dual-source 1x1 conv (pcv_conv1x1_dual): y = act(W1 x1 + W2 x2[::s] + b) as ONE GEMM over the K-concatenated operands - a
// bottleneck's last 1x1 conv with the unit's projection shortcut folded in.  `d` is the stride-1 conv over x1, `d2` the
// (possibly strided) 1x1 conv over x2; both land on the same output grid.  Served by the CTA-pair kernel's 256-wide tile.
int igemm_dual_ok(const pcv_conv_desc& d, const pcv_conv_desc& d2) {
  std::string why;
  if (!igemm_supported(d, &why) || !igemm_supported(d2, &why)) return 0;
  auto plain1x1 = [](const pcv_conv_desc& c) {
    return c.kh == 1 && c.kw == 1 && c.pad == 0 && c.groups == 1 && c.dil == 1 && c.in_row_pitch == 0 && c.Cin % 8 == 0 &&
           pitch_or(c.in_pitch, c.Cin) % 8 == 0;
  };
  if (!plain1x1(d) || !plain1x1(d2) || d.stride != 1 || d2.stride < 1 || d2.stride > 2) return 0;
  // (PCV_CONV_SE_GATE on d: the gated variant - the shortcut's half of the sum lives in a second accumulator, outside the gate)
  if ((d.flags & ~PCV_CONV_SE_GATE) != 0 || d2.flags != 0 || d2.act != PCV_ACT_NONE) return 0;
  if (d.N != d2.N || d.Cout != d2.Cout || conv_out(d2.H, 1, d2.stride, 0, 1) != d.H || conv_out(d2.W, 1, d2.stride, 0, 1) != d.W)
    return 0;
  if (d.Cout <= 128 || d.Cout % 8 || pitch_or(d.out_pitch, d.Cout) % 8) return 0;   // the 256-wide pair tile
  const char* e = getenv("PCV_IGEMM_2CTA");
  const long long M = static_cast<long long>(d.N) * d.H * d.W;
  return !(e && e[0] == '0') && ceil_div(static_cast<int>(M), BLOCK_M) >= 2;
}

int igemm_make(const pcv_conv_desc& d, const void* x, const void* w, const float* bias, const void* res, void* y,
               Op** out, const float* gate, const IgemmDual* dual) {
  std::string why;
  if (!igemm_supported(d, &why)) return fail(PCV_ERR_UNSUPPORTED, "tcgen05 conv: %s", why.c_str());
  if (dual) {
    PCV_REQUIRE(igemm_dual_ok(d, *dual->d2) && res == nullptr,
                "pcv_conv1x1_dual: layer pair outside the dual-source kernel's domain (ask pcv_conv1x1_dual_ok)");
    PCV_REQUIRE((gate != nullptr) == (dual->bias2 != nullptr), "the gated dual-source conv takes the shortcut's bias separately");
    PCV_REQUIRE(dual->x2 && reinterpret_cast<uintptr_t>(dual->x2) % 16 == 0, "second source must be 16-byte aligned");
  }
  if (gate == nullptr || !(d.flags & PCV_CONV_SE_GATE)) {
    PCV_REQUIRE(gate == nullptr && !(d.flags & PCV_CONV_SE_GATE), "PCV_CONV_SE_GATE needs the gate in `workspace` (and only then)");
  } else {
    PCV_REQUIRE(igemm_gate_ok(d), "PCV_CONV_SE_GATE: this layer is not served by the gated-epilogue kernel (ask pcv_conv_se_gate_ok)");
    PCV_REQUIRE(reinterpret_cast<uintptr_t>(gate) % 16 == 0, "the SE gate must be 16-byte aligned");
  }
  if (gate == nullptr && dual == nullptr) {
    const int rcs = stem_halo_try_make(d, x, w, bias, res, y, out);   // s2d stem with a 32-byte-row halo tile
    if (rcs != PCV_ERR_UNSUPPORTED) return rcs;
    if (d.flags & PCV_CONV_POOL3S2)
      return fail(PCV_ERR_UNSUPPORTED, "PCV_CONV_POOL3S2: this stem cannot take the fused max pool (ask pcv_stem_s2d_pool_ok)");
    const int rc3 = igemm3_try_make(d, x, w, bias, res, y, out);   // 3x3 stride-1 layers with a smem halo tile
    if (rc3 != PCV_ERR_UNSUPPORTED) return rc3;
  }
  const int Ho = conv_out(d.H, d.kh, d.stride, d.pad, d.dil);
  const int Wo = conv_out(d.W, d.kw, d.stride, d.pad, d.dil);
  PCV_REQUIRE(Ho > 0 && Wo > 0, "conv output is empty (H=%d W=%d k=%d)", d.H, d.W, d.kh);
  const int in_pitch = pitch_or(d.in_pitch, d.Cin);
  const int out_pitch = pitch_or(d.out_pitch, d.Cout);
  const int res_pitch = pitch_or(d.res_pitch, d.Cout);
  const int taps = d.kh * d.kw;
  const bool grouped = d.groups > 1;

  auto op = std::make_unique<IgemmOp>();
  op->bn = pick_bn(d, ceil_div(d.N * Ho * Wo, BLOCK_M));
  IgemmParams& p = op->p;
  p.bias = bias;
  p.out = y;
  p.res = reinterpret_cast<const e16*>(res);
  p.M = d.N * Ho * Wo;
  p.Cout = d.Cout;
  p.out_pitch = out_pitch;
  p.res_pitch = res_pitch;
  p.HoWo = Ho * Wo;
  p.Wo = Wo;
  p.stride = d.stride;
  p.pad = d.pad;
  p.dil = d.dil;
  p.kw = d.kw;
  p.cblocks = grouped ? 1 : ceil_div(d.Cin, BLOCK_K);
  p.num_kblocks = taps * p.cblocks;
  p.kb_split = p.a_mode2 = p.stride2 = 0;
  p.bias2 = dual ? dual->bias2 : nullptr;
  if (dual) {
    p.kb_split = p.cblocks;
    p.num_kblocks += ceil_div(dual->d2->Cin, BLOCK_K);
    p.stride2 = dual->d2->stride;
    p.a_mode2 = dual->d2->stride > 1 ? 1 : 0;
  }
  p.tiles_m = ceil_div(p.M, BLOCK_M);
  p.act = d.act;
  p.act_lo = (d.act == PCV_ACT_RELU || d.act == PCV_ACT_RELU6) ? 0.f : -INFINITY;
  p.act_hi = d.act == PCV_ACT_RELU6 ? 6.f : INFINITY;
  p.act_a = d.act_param;
  p.gate = gate;
  p.n_img = d.N;
  p.has_res = res != nullptr;
  p.grouped = grouped;
  p.g_in_span = grouped ? 64 * (d.Cin / d.groups) / (d.Cout / d.groups) : 0;
  p.stages = p.nstg = 0;
  p.ksub = 1;
  {
    const char* e = getenv("PCV_IGEMM_DBG");
    p.dbg = e ? atoi(e) : 0;
  }
  const bool pointwise = (taps == 1 && d.stride == 1 && d.pad == 0 && d.in_row_pitch == 0 &&
                          !(d.flags & PCV_CONV_IN_OVERLAP));
  p.a_mode = (pointwise && !(d.flags & PCV_CONV_A_IM2COL)) ? 0 : 1;
  const bool tma_out = !(d.flags & PCV_CONV_OUT_F32) && (out_pitch % 8 == 0) && (!res || res_pitch % 8 == 0) &&
                       (reinterpret_cast<uintptr_t>(y) % 16 == 0) && (reinterpret_cast<uintptr_t>(res) % 16 == 0);
  p.out_mode = tma_out ? 0 : ((d.flags & PCV_CONV_OUT_F32) ? 2 : 1);
  PCV_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(w) % 16 == 0,
              "conv operands must be 16-byte aligned");
  // CTA-pair kernel for wide dense layers: halves the L2->SMEM operand traffic per MMA (see conv_igemm2.cu)
  static const bool pair_enabled = [] {
    const char* e = getenv("PCV_IGEMM_2CTA");
    return !(e && e[0] == '0');
  }();
  // PCV_IGEMM2_MIN_COUT=16 sends narrow layers (MobileNetV2's 16/24/32-channel projections) to the pair kernel as a 64-wide
  // tile with zero-weight padding columns and clipped stores.  Measured slower than the single-CTA kernel's 32-wide tiles
  // (32->32 @112: 0.134 vs 0.116 ms: twice the TMEM reads and staging traffic for the padding), so the default stays 64.
  static const int pair_min_cout = [] {
    const char* e = getenv("PCV_IGEMM2_MIN_COUT");
    return e ? atoi(e) : 64;
  }();
  op->pair = pair_enabled && p.out_mode == 0 && d.Cout >= pair_min_cout && d.Cout % 8 == 0 && p.tiles_m >= 2;
  PCV_REQUIRE((gate == nullptr && dual == nullptr) || op->pair, "gated / dual-source conv: operands not aligned for the CTA-pair kernel");
  if (op->pair) {
    op->bn = grouped ? 64 : (d.Cout > 128 ? 256 : (d.Cout > 64 ? 128 : 64));
    igemm2_pick_smem(op->bn, p.num_kblocks, res != nullptr, taps, &p.stages, &p.ksub, &p.nstg);
    if (dual) {
      // measured on ResNet-50's four projection units (profiles/README.md, "dual-source ring sweep"): 2 k-blocks keep the
      // picker's 2x2/6; 6 k-blocks: 5x1/4 110 -> 96 us; 12 and 24 k-blocks: 3x2/2 82 -> 77 us and 81 -> 74 us
      if (p.num_kblocks >= 9) p.stages = 3, p.ksub = 2, p.nstg = 2;
      else if (p.num_kblocks >= 3) p.stages = 5, p.ksub = 1, p.nstg = 4;
      if (const char* e = getenv("PCV_IGEMM2_DUAL_CFG")) sscanf(e, "%d,%d,%d", &p.stages, &p.ksub, &p.nstg);
    }
  }
  p.tiles_n = ceil_div(d.Cout, op->bn);
  p.nsubs = 1;
  int b_box_rows = op->bn;
  if (op->pair) {
    // tile width = nsubs * 64 <= BN.  Candidates are scored by (waves of pair tiles over the machine) x (tile cost ~ width
    // + a fixed per-tile overhead): a layer whose tiles barely spill into another wave (stage 4 of ResNet-50: 49 M-pairs x
    // 2 N-tiles = 98 tiles on 74 CTA pairs, a second wave that is one third full) runs as 3 narrower N-tiles (147 tiles,
    // two FULL waves of 3/4-cost tiles); ties go to the candidate that pads ceil(Cout/64) the least, then to the widest.
    const int units = ceil_div(d.Cout, 64), maxns = op->bn / 64;
    const int pairs_avail = std::max(1, sm_count() / 2), pair_m = (p.tiles_m + 1) / 2;
    int best = maxns, bestpad = 1 << 30;
    double bestcost = 1e30;
    for (int ns = maxns; ns >= (op->bn == 256 ? 2 : maxns); --ns) {   // instantiated widths: 256 / 192 / 128, 128, 64
      const int pad = ceil_div(units, ns) * ns;
      const long long tiles = static_cast<long long>(pair_m) * ceil_div(units, ns);
      // (measured: 3x3 512->512 @7x7 0.065 -> 0.059 ms; short-K 1x1 layers are overhead-bound and gain nothing, so the
      // wave term only applies to k x k layers)
      // (the same term on the long-K 1x1 reductions of stage 4, 2048->512 @7x7 as 3 x 192-wide tiles: 40.3 -> 40.0 us, nothing)
      const double waves = taps > 1 ? static_cast<double>((tiles + pairs_avail - 1) / pairs_avail)
                                    : static_cast<double>(tiles) / pairs_avail;
      // per-tile cost ~ cycles of one K = 16 MMA step at this width (measured, shared-memory bound: N = 256 170 clk,
      // N = 192 130, N = 128 110 - narrower tiles do less work per operand byte) + a fixed share for the tile's hand-over
      const double mma_clk = ns >= 4 ? 170.0 : (ns == 3 ? 130.0 : 110.0);
      const double cost = waves * (mma_clk + 15.0);
      if (cost < bestcost * 0.97 || (cost <= bestcost * 1.03 && pad < bestpad)) {
        best = ns;
        bestpad = pad;
        bestcost = std::min(cost, bestcost);
      }
    }
    static const bool narrow = [] {
      const char* e = getenv("PCV_IGEMM2_NARROW");
      return !(e && e[0] == '0');
    }();
    // (the gated epilogue and the dual-source producer exist for full-width tiles)
    p.nsubs = (narrow && !grouped && gate == nullptr && dual == nullptr) ? best : maxns;
    if (gate != nullptr && dual != nullptr) p.nsubs = 2;   // two 128-column accumulators per TMEM buffer (conv_igemm2.cu)
    p.tiles_n = ceil_div(d.Cout, p.nsubs * 64);
    b_box_rows = p.nsubs * 32;   // each CTA of the pair loads half of the tile's weight rows
  }

  int rc;
  if (p.a_mode == 1) {
    rc = make_im2col_4d(&op->tmA, x, d, in_pitch);
  } else {
    rc = make_tiled_2d(&op->tmA, x, d.Cin, p.M, (uint64_t)in_pitch * 2, BLOCK_K, BLOCK_M, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (rc) return rc;
  const uint64_t kpad = (uint64_t)p.num_kblocks * BLOCK_K;   // (dual-source: both weight blocks, K-concatenated per output row)
  rc = make_tiled_2d(&op->tmB, w, kpad, d.Cout, kpad * 2, BLOCK_K, b_box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const int sub_cols = op->bn >= 64 ? 64 : op->bn;
  const CUtensorMapSwizzle oswz = sub_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  if (p.out_mode == 0) {
    rc = make_tiled_2d(&op->tmOut, y, d.Cout, p.M, (uint64_t)out_pitch * 2, sub_cols, BLOCK_M, oswz);
    if (rc) return rc;
    if (res) {
      rc = make_tiled_2d(&op->tmRes, res, d.Cout, p.M, (uint64_t)res_pitch * 2, sub_cols, BLOCK_M, oswz);
      if (rc) return rc;
    } else if (dual) {   // the second activation rides in the residual slot
      const pcv_conv_desc& d2 = *dual->d2;
      const int pitch2 = pitch_or(d2.in_pitch, d2.Cin);
      if (p.a_mode2 == 1) rc = make_im2col_4d(&op->tmRes, dual->x2, d2, pitch2);
      else rc = make_tiled_2d(&op->tmRes, dual->x2, d2.Cin, p.M, (uint64_t)pitch2 * 2, BLOCK_K, BLOCK_M, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    } else {
      op->tmRes = op->tmOut;
    }
  } else {
    op->tmOut = op->tmB;
    op->tmRes = op->tmB;
  }
  // Narrow, short-K layers (MobileNetV2 / V3 / EfficientNet projections and expansions to <= 32 channels): a 128 x 32
  // tile moves 16 KB and its per-tile chain (TMA -> MMA -> tcgen05.ld -> staging -> fence -> barrier -> TMA store) is
  // latency-, not bandwidth-bound (1330 clk per tile against 745 at the HBM roofline).  Two independent CTAs per SM
  // (3-stage ring each) overlap two such chains.
  static const bool twin_enabled = [] {
    const char* e = getenv("PCV_IGEMM_TWIN");
    return !(e && e[0] == '0');
  }();
  op->twin = twin_enabled && !op->pair && op->bn == 32 && p.num_kblocks <= 3 && p.tiles_m * p.tiles_n > sm_count();
  if (op->pair) op->grid = 2 * std::min(((p.tiles_m + 1) / 2) * p.tiles_n, sm_count() / 2);
  else op->grid = std::min(p.tiles_m * p.tiles_n, (op->twin ? 2 : 1) * sm_count());

  char nm[160];
  char cfg[48] = "";
  if (op->pair) {
    if (p.nsubs * 64 != op->bn) snprintf(cfg, sizeof cfg, " tw=%d st%dx%d/%d", p.nsubs * 64, p.stages, p.ksub, p.nstg);
    else snprintf(cfg, sizeof cfg, " st%dx%d/%d", p.stages, p.ksub, p.nstg);
  }
  char dl[48] = "";
  if (dual) snprintf(dl, sizeof dl, " +1x1 s%d %d@%dx%d", dual->d2->stride, dual->d2->Cin, dual->d2->H, dual->d2->W);
  snprintf(nm, sizeof nm, "conv_tc%s %dx%d s%d d%d g%d %d->%d @%dx%d bn=%d%s%s%s%s", op->pair ? "2" : "", d.kh, d.kw,
           d.stride, d.dil, d.groups, d.Cin, d.Cout, d.H, d.W, op->bn, cfg, gate ? " *gate" : "", res ? " +res" : "", dl);
  if (p.out_mode) strncat(nm, " direct", sizeof nm - strlen(nm) - 1);
  op->name = nm;
  const double e = 2.0;
  const double pin = (taps == 1 && d.stride > 1) ? (double)Ho * Wo : (double)d.H * d.W;
  op->flops = 2.0 * p.M * d.Cout * (d.Cin / d.groups) * taps;
  op->bytes = e * d.N * d.Cin * pin + ((d.flags & PCV_CONV_OUT_F32) ? 4.0 : e) * p.M * d.Cout +
              (res ? e * p.M * d.Cout : 0.0) + e * d.Cout * (d.Cin / d.groups) * taps + 4.0 * d.Cout +
              (gate ? 4.0 * d.N * d.Cout : 0.0);
  if (dual) {
    op->flops += 2.0 * p.M * d.Cout * dual->d2->Cin;
    op->bytes += e * p.M * dual->d2->Cin + e * d.Cout * dual->d2->Cin;
  }
  *out = op.release();
  return PCV_OK;
}

}  // namespace PCV_TIER
}  // namespace pcv
