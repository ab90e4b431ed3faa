// Bandwidth-bound 3x3 window ops on NHWC bf16 with TMA halo staging: depthwise convolution (+ folded BN bias,
// clamp activation, optional residual) and max pooling.
//
// Reference: dwconv3x3_block (pytorchcv/models/common/conv.py:476-508) inside LinearBottleneck
// (models/mobilenetv2.py:52-56) / DwsConvBlock (conv.py:546-608); nn.MaxPool2d(3, 2, 1) of ResInitBlock
// (models/resnet.py:255-258) and SEInitBlock (models/senet.py:154-157).
//
// Both ops move ~(1 + 1/S^2) tensors through HBM for a handful of FLOPs per byte, so the kernel is organised around
// the memory system, not the math:
//   * persistent CTAs; each tile = TH x TW output pixels x CB channels.  One elected thread fetches the tile's input
//     halo box ((TH-1)*S+KS rows x (TW-1)*S+KS columns x CB channels) with ONE 4-D TMA load into a ring of smem
//     stages (mbarrier complete_tx); image borders are zero-filled by the TMA unit, so the hot loop has no bounds
//     checks for the convolution and only predicate masks for the max.
//   * a thread owns 4 consecutive channels (8 bytes) of one output column and walks down the tile's rows: every
//     input row is read from smem once per thread (KS 8-byte LDS, conflict-free: consecutive lanes read consecutive
//     addresses) and contributes to the ceil(KS/S) output rows that are live in registers; weights stay in registers
//     as fp32 pairs and the multiply-accumulate is the packed fma.rn.f32x2 (two channels per instruction).
//   * results leave with coalesced 8-byte-per-lane global stores (CB*2 contiguous bytes per pixel).
#include <algorithm>
#include <cmath>

#include "ptx.cuh"
#include "runtime.h"

namespace pcv {
namespace PCV_TIER {

struct WinParams {
  int N, H, W, C, Ho, Wo;
  int pad;
  int out_pitch, res_pitch;
  int CB, TW, IW;
  int tiles_x, tiles_y, cblocks;
  long long num_tiles;
  int items;          // active threads per tile: TW * CB / 4
  int stages, stage_bytes, tx_bytes;   // smem stride of a stage (128 B multiple) / exact bytes of one TMA box
  float act_lo, act_hi;
  int act;                             // pcv_act; > ReLU6 (swish family) takes the fp32 path before packing
};

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// d = a * b + d on two packed fp32 lanes (Blackwell FFMA2); each lane rounds exactly like fmaf.
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  uint64_t dd = (static_cast<uint64_t>(__float_as_uint(d.y)) << 32) | __float_as_uint(d.x);
  const uint64_t aa = (static_cast<uint64_t>(__float_as_uint(a.y)) << 32) | __float_as_uint(a.x);
  const uint64_t bb = (static_cast<uint64_t>(__float_as_uint(b.y)) << 32) | __float_as_uint(b.x);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d.x = __uint_as_float(static_cast<uint32_t>(dd));
  d.y = __uint_as_float(static_cast<uint32_t>(dd >> 32));
}

__device__ __forceinline__ uint32_t hclamp2_u32(uint32_t v, uint32_t lo, uint32_t hi) {
  const e16x2 x = *reinterpret_cast<const e16x2*>(&v);
  const e16x2 r = __hmin2(__hmax2(x, *reinterpret_cast<const e16x2*>(&lo)),
                                   *reinterpret_cast<const e16x2*>(&hi));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  const e16x2 r = __hmax2(*reinterpret_cast<const e16x2*>(&a), *reinterpret_cast<const e16x2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// Tile coordinates as a mixed-radix counter (channel block fastest, then tile column, tile row, image) that is
// advanced by the CTA's stride with carries: no integer division in the persistent loop.
struct TileCoord {
  int cb, tx, ty, n;
};
__device__ __forceinline__ TileCoord tile_decode(const WinParams& p, uint32_t t) {
  TileCoord tc;
  tc.cb = t % p.cblocks; t /= p.cblocks;
  tc.tx = t % p.tiles_x; t /= p.tiles_x;
  tc.ty = t % p.tiles_y;
  tc.n = t / p.tiles_y;
  return tc;
}
__device__ __forceinline__ void tile_advance(const WinParams& p, TileCoord& tc, const TileCoord& step) {
  tc.cb += step.cb;
  if (tc.cb >= p.cblocks) { tc.cb -= p.cblocks; ++tc.tx; }
  tc.tx += step.tx;
  if (tc.tx >= p.tiles_x) { tc.tx -= p.tiles_x; ++tc.ty; }
  tc.ty += step.ty;
  if (tc.ty >= p.tiles_y) { tc.ty -= p.tiles_y; ++tc.n; }
  tc.n += step.n;
}

__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

// OP 0: depthwise conv, 1: max pool.  COLS = adjacent output columns per thread (their windows overlap, so the shared
// input columns are loaded and converted once).  RES = residual add in the epilogue.
template <int OP, int KS, int S, int TH, int COLS, bool RES>
__global__ void __launch_bounds__(KS == 5 ? 384 : 256, KS == 5 ? 1 : 2)   // 5x5: 25 taps x 4 channels of fp32 weights live in
                                                                         // registers (~165): one 384-thread CTA per SM
win_kernel(const __grid_constant__ CUtensorMap tmIn, const WinParams p, const float* __restrict__ w,
           const float* __restrict__ bias, const e16* __restrict__ res, e16* __restrict__ y) {
  constexpr int IH = (TH - 1) * S + KS;
  constexpr int NACC = (KS + S - 1) / S;
  constexpr int NIN = (COLS - 1) * S + KS;      // input columns touched by one thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  const int tid = threadIdx.x;
  const uint32_t num_tiles = static_cast<uint32_t>(p.num_tiles);

  const TileCoord step = tile_decode(p, gridDim.x);          // this CTA's stride as mixed-radix digits
  TileCoord tc = tile_decode(p, blockIdx.x);                 // tile being computed
  TileCoord lc = tc;                                         // tile being loaded (thread 0), `stages` tiles ahead
  uint32_t lt = blockIdx.x;

  if (tid == 0) {
    tma_prefetch_desc(&tmIn);
    for (int i = 0; i < p.stages; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();   // nothing of the previous kernel's output is read, and nothing it may still read is overwritten, before this
  if (tid == 0) {
    for (int i = 0; i < p.stages && lt < num_tiles; ++i) {
      mbar_arrive_expect_tx(&full[i], p.tx_bytes);
      tma_load_4d(&tmIn, &full[i], smem + i * p.stage_bytes, lc.cb * p.CB, lc.tx * p.TW * S - p.pad,
                  lc.ty * TH * S - p.pad, lc.n);
      lt += gridDim.x;
      tile_advance(p, lc, step);
    }
  }

  const int nv = p.CB >> 2;
  const bool active = tid < p.items;
  const int colg = active ? tid / nv : 0;                    // column group: output columns colg*COLS ..
  const int v = active ? tid - colg * nv : 0;
  const int col = colg * COLS;
  const int row_bytes = p.IW * p.CB * 2;
  const int px_bytes = p.CB * 2;
  const uint32_t thread_off = smem_u32(smem) + (col * S * p.CB + v * 4) * 2;
  const int H = p.H, W = p.W, Ho = p.Ho, Wo = p.Wo, pad = p.pad;
  const uint32_t y_row = static_cast<uint32_t>(Wo) * p.out_pitch, r_row = static_cast<uint32_t>(Wo) * p.res_pitch;
  const int out_pitch = p.out_pitch, res_pitch = p.res_pitch;

  float2 wr[OP == 0 ? KS * KS : 1][2];
  float2 b2[2];
  auto load_weights = [&](int c) {
    if (OP == 0) {
#pragma unroll
      for (int k = 0; k < KS * KS; ++k) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(w + static_cast<size_t>(k) * p.C + c));
        wr[k][0] = make_float2(q.x, q.y);
        wr[k][1] = make_float2(q.z, q.w);
      }
      const float4 q = __ldg(reinterpret_cast<const float4*>(bias + c));
      b2[0] = make_float2(q.x, q.y);
      b2[1] = make_float2(q.z, q.w);
    }
  };
  if (p.cblocks == 1) load_weights(v * 4);                   // one channel block: weights are loop-invariant
  const uint32_t act_lo2 = pack_e16x2(p.act_lo, p.act_lo), act_hi2 = pack_e16x2(p.act_hi, p.act_hi);

  int stage = 0;
  uint32_t phase = 0;
  for (uint32_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    const int c = tc.cb * p.CB + v * 4;
    if (p.cblocks != 1) load_weights(c);
    const int ho0 = tc.ty * TH;
    const int wo = tc.tx * p.TW + col;
    bool col_ok[COLS];
#pragma unroll
    for (int cc = 0; cc < COLS; ++cc) col_ok[cc] = active && (wo + cc) < Wo;
    // max pool: which input columns of this thread's windows fall inside the image
    bool in_ok[NIN];
#pragma unroll
    for (int j = 0; j < NIN; ++j) {
      const int wi = wo * S - pad + j;
      in_ok[j] = wi >= 0 && wi < W;
    }
    const int hi0 = ho0 * S - pad;
    // 32-bit element offsets (the host rejects tensors with >= 2^31 elements): one IMAD per store instead of a
    // 64-bit multiply chain
    const uint32_t pix0 = (static_cast<uint32_t>(tc.n) * Ho + ho0) * Wo + wo;
    const uint32_t yo = pix0 * out_pitch + c;
    const uint32_t ro = RES ? pix0 * res_pitch + c : 0u;
    const int rows_left = Ho - ho0;          // output rows of this tile that exist

    mbar_wait(&full[stage], phase);
    const uint32_t sbase = thread_off + stage * p.stage_bytes;

    float2 acc[NACC][COLS][2];
    uint2 mx[NACC][COLS];
#pragma unroll
    for (int ir = 0; ir < IH; ++ir) {
      uint2 raw[NIN];
#pragma unroll
      for (int j = 0; j < NIN; ++j) raw[j] = lds64(sbase + ir * row_bytes + j * px_bytes);
      float2 xv[NIN][2];
      if (OP == 0) {
#pragma unroll
        for (int j = 0; j < NIN; ++j) {
          xv[j][0] = make_float2(e16lo(raw[j].x), e16hi(raw[j].x));
          xv[j][1] = make_float2(e16lo(raw[j].y), e16hi(raw[j].y));
        }
      }
      const bool row_ok = (hi0 + ir) >= 0 && (hi0 + ir) < H;
#pragma unroll
      for (int fr = 0; fr < KS; ++fr) {
        if ((ir - fr) < 0 || (ir - fr) % S != 0 || (ir - fr) / S >= TH) continue;   // compile-time after unrolling
        const int ho = (ir - fr) / S;
        const int a = ho % NACC;
#pragma unroll
        for (int cc = 0; cc < COLS; ++cc) {
          if (OP == 0) {
            if (fr == 0) {
              acc[a][cc][0] = b2[0];
              acc[a][cc][1] = b2[1];
            }
#pragma unroll
            for (int fs = 0; fs < KS; ++fs) {
              ffma2(acc[a][cc][0], xv[cc * S + fs][0], wr[fr * KS + fs][0]);
              ffma2(acc[a][cc][1], xv[cc * S + fs][1], wr[fr * KS + fs][1]);
            }
          } else {
            if (fr == 0) mx[a][cc] = make_uint2(E16_NEG_INF2, E16_NEG_INF2);   // -inf, -inf
            if (row_ok) {
#pragma unroll
              for (int fs = 0; fs < KS; ++fs) {
                if (in_ok[cc * S + fs]) {
                  mx[a][cc].x = hmax2_u32(mx[a][cc].x, raw[cc * S + fs].x);
                  mx[a][cc].y = hmax2_u32(mx[a][cc].y, raw[cc * S + fs].y);
                }
              }
            }
          }
        }
        if (fr == KS - 1) {
          const bool row_live = ho < rows_left;
#pragma unroll
          for (int cc = 0; cc < COLS; ++cc) {
            const bool live = row_live && col_ok[cc];
            // one column per thread: a real branch around the whole epilogue measured faster (5.85 vs 5.0 TB/s on
            // the stride-2 layers); two columns: compute unconditionally and predicate only the memory operations
            if (COLS > 1 || live) {
              uint2 o;
              if (OP == 0) {
                float2 r0 = acc[a][cc][0], r1 = acc[a][cc][1];
                if (RES) {
                  uint2 rr = make_uint2(0u, 0u);
                  if (live) rr = __ldg(reinterpret_cast<const uint2*>(res + (ro + ho * r_row + cc * res_pitch)));
                  r0.x += e16lo(rr.x); r0.y += e16hi(rr.x); r1.x += e16lo(rr.y); r1.y += e16hi(rr.y);
                }
                if (p.act > PCV_ACT_RELU6) {   // swish / h-swish / (h-)sigmoid (EfficientNet, MobileNetV3): warp-uniform
                  float f[4] = {r0.x, r0.y, r1.x, r1.y};
                  fast_act_n(f, p.act);
                  r0 = make_float2(f[0], f[1]);
                  r1 = make_float2(f[2], f[3]);
                }
                // clamp AFTER rounding to bf16: the bounds (0, 6, +-inf) are bf16-exact and rounding is monotone,
                // so this equals round(clamp(x)) at half the instruction count (packed bf16x2 min / max)
                o.x = hclamp2_u32(pack_e16x2(r0.x, r0.y), act_lo2, act_hi2);
                o.y = hclamp2_u32(pack_e16x2(r1.x, r1.y), act_lo2, act_hi2);
              } else {
                o = mx[a][cc];
              }
              if (live) *reinterpret_cast<uint2*>(y + (yo + ho * y_row + cc * out_pitch)) = o;
            }
          }
        }
      }
    }

    __syncthreads();   // every thread is done reading this stage
    if (tid == 0 && lt < num_tiles) {
      mbar_arrive_expect_tx(&full[stage], p.tx_bytes);
      tma_load_4d(&tmIn, &full[stage], smem + stage * p.stage_bytes, lc.cb * p.CB, lc.tx * p.TW * S - p.pad,
                  lc.ty * TH * S - p.pad, lc.n);
      lt += gridDim.x;
      tile_advance(p, lc, step);
    }
    tile_advance(p, tc, step);
    if (++stage == p.stages) {
      stage = 0;
      phase ^= 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: tile geometry, tensor map, launch
// ---------------------------------------------------------------------------------------------------------------
struct WinCfg {
  int TH, threads, ctas_per_sm, smem_bytes, cols;
};

static bool pick_cfg(int C, int Ho, int Wo, int KS, int S, int cols, WinParams* p, WinCfg* cfg) {
  double best = -1.0;
  cfg->cols = cols;
  for (int CB = 8; CB <= std::min(C, 256); CB += 8) {
    if (C % CB != 0) continue;
    const int nv = CB / 4;
    const int max_threads = KS == 5 ? 384 : 256;   // __launch_bounds__ of win_kernel
    for (int TW = cols; TW <= round_up(Wo, cols) && TW / cols * nv <= max_threads; TW += cols) {
      const int IW = (TW - 1) * S + KS;
      if (IW > 256) break;
      for (int TH = 7; TH <= 8; ++TH) {
        const int IH = (TH - 1) * S + KS;
        const int stage = (IH * IW * CB * 2 + 127) & ~127;
        if (stage > 56 * 1024) continue;
        const int items = TW / cols * nv, threads = round_up(items, 32);
        const double e_w = static_cast<double>(Wo) / (ceil_div(Wo, TW) * TW);
        const double e_h = static_cast<double>(Ho) / (ceil_div(Ho, TH) * TH);
        const double e_t = static_cast<double>(items) / threads;
        const double halo = static_cast<double>(TH * S * TW * S) / (IH * IW);
        const double wide = std::min(1.0, CB * 2 / 64.0);          // >= 64 contiguous bytes per pixel and DRAM burst
        const double big = std::min(1.0, items * cols / 128.0);     // enough work per CTA to cover latency
        // 5x5: ~160 registers per thread cap an SM at ~400 resident threads, as k CTAs of <= 400/k threads each
        // (ncu: a single 128-thread CTA per SM left 6 % of the warp slots occupied)
        const double occ = KS == 5 ? std::min(1.0, std::min(3, 400 / threads) * threads / 384.0) : 1.0;
        const double score = e_w * e_h * e_t * std::sqrt(std::min(1.0, halo)) * wide * big * occ + 1e-6 * stage;
        if (score > best) {
          best = score;
          p->CB = CB; p->TW = TW; p->IW = IW; p->items = items; p->stage_bytes = stage; p->tx_bytes = IH * IW * CB * 2;
          cfg->TH = TH; cfg->threads = threads;
        }
      }
    }
  }
  if (best < 0) return false;
  if (KS == 5) {
    const int fit = std::max(1, std::min(3, 400 / cfg->threads));        // CTAs per SM the register file allows
    p->stages = std::max(2, std::min(4, (216 * 1024 / fit) / p->stage_bytes));
    cfg->smem_bytes = p->stages * p->stage_bytes + 8 * 8 + 128;
    cfg->ctas_per_sm = std::max(1, std::min(fit, (220 * 1024) / cfg->smem_bytes));
  } else {
    const int budget = 100 * 1024;   // per CTA; two CTAs per SM
    p->stages = std::max(2, std::min(4, budget / p->stage_bytes));
    cfg->smem_bytes = p->stages * p->stage_bytes + 8 * 8 + 128;
    cfg->ctas_per_sm = std::max(1, std::min(2, (220 * 1024) / cfg->smem_bytes));
  }
  p->tiles_x = ceil_div(Wo, p->TW);
  p->tiles_y = ceil_div(Ho, cfg->TH);
  p->cblocks = C / p->CB;
  return true;
}

static int make_win_map(CUtensorMap* tm, const void* x, int N, int H, int W, int C, int in_pitch, int CB, int IW,
                        int IH) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)W * in_pitch * 2, (cuuint64_t)H * W * in_pitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)IW, (cuuint32_t)IH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, TMAP_E16, 4, const_cast<void*>(x), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled (window op) failed (%d): C=%d W=%d H=%d N=%d box=%dx%dx%d", (int)r, C,
                W, H, N, CB, IW, IH);
  return PCV_OK;
}

struct WinOp : Op {
  CUtensorMap tm;
  WinParams p;
  WinCfg cfg;
  int op_kind, S, K = 3;
  const float *w, *bias;
  const e16* res;
  e16* y;

  template <int OP, int K_, int S_, int TH, int COLS, bool RES>
  cudaError_t go(cudaStream_t s) {
    static std::atomic<uint64_t> attr_done{0};   // per device (see runtime.h)
    if (cudaError_t e = set_max_smem_once(win_kernel<OP, K_, S_, TH, COLS, RES>, (K_ == 5 ? 224 : 112) * 1024, attr_done)) return e;
    const long long cap = static_cast<long long>(sm_count()) * cfg.ctas_per_sm;
    const int grid = static_cast<int>(std::min<long long>(p.num_tiles, cap));
    return launch_pdl(win_kernel<OP, K_, S_, TH, COLS, RES>, dim3(grid), dim3(cfg.threads), cfg.smem_bytes, s, tm, p, w, bias, res, y);
  }
  template <int OP, int K_, int S_, int COLS, bool RES>
  cudaError_t go_th(cudaStream_t s) {
    return cfg.TH == 7 ? go<OP, K_, S_, 7, COLS, RES>(s) : go<OP, K_, S_, 8, COLS, RES>(s);
  }
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    if (op_kind == 1) return S == 1 ? go_th<1, 3, 1, 1, false>(s) : go_th<1, 3, 2, 1, false>(s);
    if (K == 5) {   // dwconv5x5_block (conv.py:511-543): one output column per thread
      if (S == 1) return res ? go_th<0, 5, 1, 1, true>(s) : go_th<0, 5, 1, 1, false>(s);
      return res ? go_th<0, 5, 2, 1, true>(s) : go_th<0, 5, 2, 1, false>(s);
    }
    if (S == 1) return res ? go_th<0, 3, 1, 2, true>(s) : go_th<0, 3, 1, 2, false>(s);
    return res ? go_th<0, 3, 2, 1, true>(s) : go_th<0, 3, 2, 1, false>(s);
  }
};

// Returns PCV_OK and sets *out, or PCV_ERR_UNSUPPORTED (no message) when the shape is outside this kernel's domain
// and the caller should use its generic kernel.
int win_make(int op_kind, int N, int H, int W, int C, int k, int stride, int pad, int act, const void* x, int in_pitch,
             const float* w, const float* bias, const void* res, int res_pitch, void* y, int out_pitch, Op** out) {
  if ((k != 3 && !(k == 5 && op_kind == 0)) || (stride != 1 && stride != 2) || pad > k / 2 || C % 8 != 0 || in_pitch % 8 != 0 || out_pitch % 4 != 0 ||
      (res && res_pitch % 4 != 0) || act > PCV_ACT_HSIGMOID || reinterpret_cast<uintptr_t>(x) % 16 != 0 ||
      reinterpret_cast<uintptr_t>(y) % 8 != 0 || reinterpret_cast<uintptr_t>(res) % 8 != 0)
    return PCV_ERR_UNSUPPORTED;
  if (getenv("PCV_WIN_TMA") && getenv("PCV_WIN_TMA")[0] == '0') return PCV_ERR_UNSUPPORTED;
  auto op = std::make_unique<WinOp>();
  WinParams& p = op->p;
  p.N = N; p.H = H; p.W = W; p.C = C;
  p.Ho = (H + 2 * pad - k) / stride + 1;
  p.Wo = (W + 2 * pad - k) / stride + 1;
  if (p.Ho <= 0 || p.Wo <= 0) return PCV_ERR_UNSUPPORTED;
  p.pad = pad; p.out_pitch = out_pitch; p.res_pitch = res_pitch;
  p.act_lo = (act == PCV_ACT_RELU || act == PCV_ACT_RELU6) ? 0.f : -INFINITY;
  p.act_hi = act == PCV_ACT_RELU6 ? 6.f : INFINITY;
  p.act = act;
  const int cols = (op_kind == 0 && stride == 1 && k == 3) ? 2 : 1;   // must match the COLS template argument picked in launch()
  if (!pick_cfg(C, p.Ho, p.Wo, k, stride, cols, &p, &op->cfg)) return PCV_ERR_UNSUPPORTED;
  p.num_tiles = static_cast<long long>(N) * p.tiles_y * p.tiles_x * p.cblocks;
  if (p.num_tiles >= (1ll << 31)) return PCV_ERR_UNSUPPORTED;
  if (static_cast<long long>(N) * p.Ho * p.Wo * std::max(out_pitch, res_pitch) >= (1ll << 31)) return PCV_ERR_UNSUPPORTED;
  const int IH = (op->cfg.TH - 1) * stride + k;
  if (int rc = make_win_map(&op->tm, x, N, H, W, C, in_pitch, p.CB, p.IW, IH)) return rc;
  op->op_kind = op_kind; op->S = stride; op->K = k; op->w = w; op->bias = bias;
  op->res = reinterpret_cast<const e16*>(res);
  op->y = reinterpret_cast<e16*>(y);
  *out = op.release();
  return PCV_OK;
}

}  // namespace PCV_TIER
}  // namespace pcv
