// Whole inverted-residual block in ONE kernel (LinearBottleneck.forward, mobilenetv2.py:62-71; the expand -> depthwise ->
// project pattern of fbnet.py:77-87, spnasnet.py:72-82, proxylessnas.py:64-70):
//
//   y = act_pw( W_pw * act_dw( dw3x3( act_exp(W_exp * x + b_exp) ) + b_dw ) + b_pw  [+ residual] )
//
// The EXPANDED tensor - six times the block's input, the widest tensor of the network, written once and read once by the
// two-kernel plan (1x1 expansion, then conv_dwpw.cu) - never exists in HBM: the kernel reads the narrow input and writes
// the narrow output (MobileNetV2 bs256, 16 -> 96 -> 24 @112x112: 1.35 GB of traffic -> 0.14 GB).
//
// A tile is TH x 16 output pixels of one image (TH = 8 at stride 1, 4 at stride 2).  Per tile ONE 4-D TMA load brings the input
// halo box ((TH-1)S+3 rows x 15S+3 columns x Cin channels, out-of-image pixels zero-filled) into shared memory as a K-major
// SWIZZLE_128B MMA A operand with "row" = halo pixel.  Per 64-channel block of the expanded width:
//   1. the EXPANSION MMA warp multiplies the halo pixels (2 or 3 M-blocks of 128 rows) by that block's resident 64 x Cin weight
//      slice (ceil(Cin/16) K steps) into one of two TMEM accumulators;
//   2. a TEAM of four CUDA-core warps (two teams take alternate blocks) turns its accumulator into the activated 16-bit
//      expanded halo in the team's stage buffer (+ bias, clamp, ZERO for out-of-image pixels: the depthwise conv pads the
//      expanded tensor, not the input), 16-byte chunks XOR-swizzled by the halo column so that neither the row-per-thread
//      stores nor the stencil's loads conflict,
//   3. and runs the depthwise stencil of conv_dwpw.cu on it (fp32 FFMA2, weights in registers, tail blocks re-mapped onto
//      fewer channels), writing the activated result as the K-major A operand of the projection;
//   4. the PROJECTION MMA warp accumulates W_pw over the channel blocks in TMEM (double buffered across tiles) and four
//      epilogue warps add bias (+ the unit's identity), clamp and store the finished tile.
// The other team is half a block out of step, so TMEM round trips and barrier waits of one hide under the arithmetic of the
// other; the expansion MMA of block c+2 runs while block c is in the stencil.  Halo pixels are expanded once per tile that
// needs them (1.16x at stride 2): tensor work, which this CUDA-core-bound kernel has to spare.
//
// Measured (MobileNetV2 bs256, fp16, same box): stride-2 units 16->96->24 @112: 213 + 162 us (expansion, then pcv_dw_pw_fused)
// -> 327 us; 24->144->32 @56: 86 + 76 -> 134; 32->192->64 @28: 30 + 30 -> 51.  Stride-1 units are at parity (24->144->24 @56:
// 214 -> 227 us): the kernel is bound by its CUDA-core work (stencil + conversion of the expanded halo at ~0.3 IPC per warp),
// not by HBM, so the plan compiler records only stride-2 triples by default (plan.py).  A first version with the conversion on
// four dedicated warps and a stage-buffer ring was 1.7x slower than the two-kernel plan (TMEM latency exposed on one warp per
// scheduler).
//
// Domain: 1x1 stride-1 expansion with Cin <= 64 and a clamp-family activation, 3x3 depthwise (stride 1 or 2, pad 1), 1x1
// projection with Cout <= 128 (stride 1) / 64 (stride 2), maps >= 14 wide at the output.
#define PCV_MBAR_SUSPEND_NS 20000u   // idle role warps sleep in mbarrier.try_wait instead of re-polling (ptx.cuh)
#include "igemm_common.cuh"

namespace pcv {
namespace PCV_TIER {

constexpr int XD_TW = 16;
constexpr int XD_THREADS = 512;   // warps: 0 producer, 1 expansion MMA, 2 TMEM allocator, 3 projection MMA, 4-7 projection epilogue, 8-15 two stencil teams
constexpr int XD_NA = 4;
constexpr int XD_SMEM = 232448 - 1024;
constexpr int XD_WROW = 10 * BLOCK_K;   // floats of depthwise weights (9 taps + bias) per 64-channel block

struct XdParams {
  const float* w_dw;     // [9][C] fp32, BN folded
  const float* b_dw;     // [C]
  const float* b_exp;    // [>= C]
  const float* b_pw;     // [>= Cout]
  const e16* res;
  e16* out;
  int N, H, W, Cin, C, Ho, Wo, Cout;
  int kx;                // K = 16 steps of the expansion GEMM: ceil(Cin / 16)
  int nmma;              // projection MMA N: Cout rounded up to 16
  uint32_t idesc_pw;
  int out_pitch, res_pitch;
  int tiles_x, tiles_y, num_tiles, ncb;
  int stages, halo_bytes, xbufs, x_bytes, b_bytes;
  int na_shift;
  float ex_lo, ex_hi, dw_lo, dw_hi, pw_lo, pw_hi;
};

__device__ __forceinline__ void xd_tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void xd_ffma2(float2& d, const float2 a, const float2 b) {
  uint64_t dd = (static_cast<uint64_t>(__float_as_uint(d.y)) << 32) | __float_as_uint(d.x);
  const uint64_t aa = (static_cast<uint64_t>(__float_as_uint(a.y)) << 32) | __float_as_uint(a.x);
  const uint64_t bb = (static_cast<uint64_t>(__float_as_uint(b.y)) << 32) | __float_as_uint(b.x);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d.x = __uint_as_float(static_cast<uint32_t>(dd));
  d.y = __uint_as_float(static_cast<uint32_t>(dd >> 32));
}
__device__ __forceinline__ uint32_t xd_lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void xd_sts32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__host__ __device__ __forceinline__ int xd_ksteps(int C, int cb) {   // K = 16 steps of a (tail) channel block, as conv_dwpw.cu
  const int cvalid = C - cb * BLOCK_K;
  return cvalid > 32 ? 4 : (cvalid > 16 ? 2 : 1);
}

// The depthwise stencil of conv_dwpw.cu on a stage buffer whose 16-byte chunks are XOR-swizzled by the halo column: the thread's
// NJ input columns have loop-invariant chunk positions, so each is one base register and every load is [base_j + immediate].
template <int S, int TH, int NQ>
__device__ __forceinline__ void xd_stencil(const uint32_t stage_u32, const uint32_t abase, const int row0, const int cp,
                                           const float2 (&wr)[9], const float2 b2, const bool dw_relu, const uint32_t dw_hi2) {
  constexpr int IH = (TH - 1) * S + 3, IW = (XD_TW - 1) * S + 3;
  constexpr int NACC = (3 + S - 1) / S;
  constexpr int NJ = (NQ - 1) * S + 3;
  uint32_t sb[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const uint32_t ix = static_cast<uint32_t>(row0 * S + j);
    sb[j] = stage_u32 + ix * 128 + (((static_cast<uint32_t>(cp) >> 2) ^ (ix & 7u)) << 4) + (static_cast<uint32_t>(cp) & 3u) * 4;
  }
  uint32_t a_off[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const uint32_t r = static_cast<uint32_t>(row0 + q);
    a_off[q] = r * 128 + (((static_cast<uint32_t>(cp) >> 2) ^ (r & 7u)) << 4) + (static_cast<uint32_t>(cp) & 3u) * 4;
  }
  float2 acc[NACC][NQ];
  uint32_t raw[2][NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) raw[0][j] = xd_lds32(sb[j]);
#pragma unroll
  for (int ir = 0; ir < IH; ++ir) {
    if (ir + 1 < IH) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) raw[(ir + 1) & 1][j] = xd_lds32(sb[j] + (ir + 1) * IW * 128);
    }
    float2 xv[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) xv[j] = make_float2(e16lo(raw[ir & 1][j]), e16hi(raw[ir & 1][j]));
#pragma unroll
    for (int fr = 0; fr < 3; ++fr) {
      if ((ir - fr) < 0 || (ir - fr) % S != 0 || (ir - fr) / S >= TH) continue;   // compile-time after unrolling
      const int ho = (ir - fr) / S;
      const int a = ho % NACC;
      if (fr == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[a][q] = b2;
      }
#pragma unroll
      for (int fs = 0; fs < 3; ++fs) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) xd_ffma2(acc[a][q], xv[q * S + fs], wr[fr * 3 + fs]);
      }
      if (fr == 2) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const uint32_t o = dw_relu ? pack_relu_e16x2(acc[a][q].x, acc[a][q].y) : pack_e16x2(acc[a][q].x, acc[a][q].y);
          xd_sts32(abase + ho * (XD_TW * 128) + a_off[q], hmin2_e16(o, dw_hi2));
        }
      }
    }
  }
}

template <int S>
__global__ void __launch_bounds__(XD_THREADS, 1)
xdwpw_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWx,
             const __grid_constant__ CUtensorMap tmB, const XdParams p) {
  constexpr int TH = S == 1 ? 8 : 4;
  constexpr int IH = (TH - 1) * S + 3, IW = (XD_TW - 1) * S + 3;
  constexpr int NPX = IH * IW;                       // halo pixels: 180 (stride 1) / 297 (stride 2)
  constexpr int NMB = (NPX + BLOCK_M - 1) / BLOCK_M;  // expansion M-blocks: 2 / 3
  constexpr int EXP_COLS = NMB * 64;                  // TMEM columns of one expansion accumulator
  constexpr int PROJ0 = 2 * EXP_COLS;                 // projection accumulators: [PROJ0, 512), two of them
  constexpr int PROJ_COLS = (512 - PROJ0) / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                                            // xbufs x [NMB*128 rows x 128 B] input halo (expansion A operand)
  uint8_t* sStage = sX + p.xbufs * p.x_bytes;                    // stages x [NPX x 128 B] expanded halo
  uint8_t* sA = sStage + p.stages * p.halo_bytes;                // (1 << na_shift) x [128 x 128 B] projection A operand
  uint8_t* sWx = sA + (BLOCK_M * 128 << p.na_shift);             // ncb x [64 x 128 B] expansion weights
  uint8_t* sB = sWx + p.ncb * 8192;                              // ncb x [nmma x 128 B] projection weights
  float* sW = reinterpret_cast<float*>(sB + p.ncb * p.b_bytes);  // [ncb][10][64] depthwise weights + bias
  float* sBx = sW + p.ncb * XD_WROW;                             // [ncb * 64] expansion bias
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBx + p.ncb * BLOCK_K);
  uint64_t* a_full = bars;                        // [XD_NA] the 4 warps of a team
  uint64_t* a_empty = a_full + XD_NA;             // [XD_NA] projection MMA commit
  uint64_t* d_full = a_empty + XD_NA;             // [2] projection MMA commit
  uint64_t* d_empty = d_full + 2;                 // [2] 4 epilogue warps
  uint64_t* e_full = d_empty + 2;                 // [2] expansion MMA commit
  uint64_t* e_empty = e_full + 2;                 // [2] the 4 warps of the team that drained it
  uint64_t* x_full = e_empty + 2;                 // [2] input halo landed
  uint64_t* x_empty = x_full + 2;                 // [2] expansion MMA commit (last block of the tile)
  uint64_t* w_full = x_empty + 2;                 // resident weights landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int my_tiles = blockIdx.x < p.num_tiles ? (p.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWx);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < XD_NA; ++i) {
      mbar_init(&a_full[i], 4);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 1);
      mbar_init(&d_empty[i], 4);
      mbar_init(&e_full[i], 1);
      mbar_init(&e_empty[i], 4);
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
    }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  // depthwise weights + bias and the expansion bias of every channel block, zero beyond C (constants: read before the PDL wait)
  for (int i = threadIdx.x; i < p.ncb * XD_WROW; i += XD_THREADS) {
    const int cbk = i / BLOCK_K, c = (cbk / 10) * BLOCK_K + (i - cbk * BLOCK_K), k = cbk % 10;
    sW[i] = c < p.C ? __ldg(k < 9 ? p.w_dw + static_cast<size_t>(k) * p.C + c : p.b_dw + c) : 0.f;
  }
  for (int i = threadIdx.x; i < p.ncb * BLOCK_K; i += XD_THREADS) sBx[i] = i < p.C ? __ldg(p.b_exp + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();
  pdl_wait();

  auto tile_coords = [&](int t, int& n, int& ty, int& tx) {
    tx = t % p.tiles_x;
    const int r = t / p.tiles_x;
    ty = r % p.tiles_y;
    n = r / p.tiles_y;
  };

  if (warp == 0) {
    // ===================================== producer: resident weights, then one input halo per tile ====================
    if (my_tiles > 0 && elect_one()) {
      mbar_arrive_expect_tx(w_full, p.ncb * (8192 + p.nmma * 128));
      for (int cb = 0; cb < p.ncb; ++cb) {
        tma_load_2d(&tmWx, w_full, sWx + cb * 8192, 0, cb * BLOCK_K);
        tma_load_2d(&tmB, w_full, sB + cb * p.b_bytes, cb * BLOCK_K, 0);
      }
    }
    __syncwarp();
    for (int it = 0; it < my_tiles; ++it) {
      int n, ty, tx;
      tile_coords(blockIdx.x + it * gridDim.x, n, ty, tx);
      const int xb = it % p.xbufs;
      mbar_wait(&x_empty[xb], ((it / p.xbufs) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&x_full[xb], NPX * 128);
        xd_tma_load_4d(&tmX, &x_full[xb], sX + xb * p.x_bytes, 0, tx * XD_TW * S - 1, ty * TH * S - 1, n);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================== expansion MMA: halo pixels x W_exp slice -> TMEM ============================
    constexpr uint32_t IDESC_X = make_idesc_e16(BLOCK_M, 64);
    int g = 0;
    if (my_tiles > 0) mbar_wait(w_full, 0);
    for (int it = 0; it < my_tiles; ++it) {
      const int xb = it % p.xbufs;
      mbar_wait(&x_full[xb], (it / p.xbufs) & 1);
      const uint32_t x_lo = smem_desc_lo(smem_u32(sX + xb * p.x_bytes));
      for (int cb = 0; cb < p.ncb; ++cb, ++g) {
        const int eb = g & 1;
        mbar_wait(&e_empty[eb], ((g >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t w_lo = smem_desc_lo(smem_u32(sWx + cb * 8192));
        const uint32_t d0 = tmem_base + eb * EXP_COLS;
        if (elect_one()) {
#pragma unroll
          for (int mb = 0; mb < NMB; ++mb) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < p.kx) umma_bf16_lohi(d0 + mb * 64, x_lo + mb * (BLOCK_M * 128 >> 4) + 2 * k, w_lo + 2 * k, IDESC_X, k != 0 ? 1u : 0u);
          }
          umma_commit(&e_full[eb]);
          if (cb == p.ncb - 1) umma_commit(&x_empty[xb]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // ===================================== projection MMA: stencil output x W_pw slice, accumulated over the blocks ====
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA));
    int g = 0;
    if (my_tiles > 0) mbar_wait(w_full, 0);
    for (int it = 0; it < my_tiles; ++it) {
      const int buf = it & 1;
      mbar_wait(&d_empty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + PROJ0 + buf * PROJ_COLS;
      for (int cb = 0; cb < p.ncb; ++cb, ++g) {
        const int ab = g & ((1 << p.na_shift) - 1);
        mbar_wait(&a_full[ab], (g >> p.na_shift) & 1);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + ab * (BLOCK_M * 128 >> 4);
        const uint32_t b_lo = smem_desc_lo(smem_u32(sB + cb * p.b_bytes));
        const int ks = xd_ksteps(p.C, cb);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            if (k < ks) umma_bf16_lohi(d_tmem, a_lo + 2 * k, b_lo + 2 * k, p.idesc_pw, (cb | k) != 0 ? 1u : 0u);
          umma_commit(&a_empty[ab]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&d_full[buf]);
      __syncwarp();
    }
  } else if (warp >= 8) {
    // ===================================== expansion epilogue + depthwise stencil on CUDA cores ========================
    // Two TEAMS of four warps take alternate channel blocks (team = g & 1 = the expansion accumulator they read).  A team first
    // turns its accumulator into the activated 16-bit expanded halo in ITS stage buffer (a thread owns one halo pixel per
    // M-block: + bias, clamp, zero outside the image, 16-byte chunks XOR-swizzled by the halo column), then runs the stencil
    // of conv_dwpw.cu on it; the other team is half a block out of step, so one team's TMEM round trips and barrier waits hide
    // under the other's arithmetic.
    const int team = (warp - 8) >> 2;
    const int quad = (warp - 8) & 3;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t dw_hi2 = pack_e16x2(p.dw_hi, p.dw_hi);
    const bool dw_relu = p.dw_lo == 0.f;
    const bool ex_relu = p.ex_lo == 0.f, ex_cap = p.ex_hi != INFINITY;
    const uint32_t ex_hi2 = pack_e16x2(p.ex_hi, p.ex_hi);
    const uint32_t stage_u32 = smem_u32(sStage + team * p.halo_bytes);
    const int total_g = my_tiles * p.ncb;
    float2 wr[9], b2;
    // channel block and tile coordinates advance incrementally (g += 2; a tile step is a mixed-radix add of gridDim.x): the
    // divisions they replace were ~15 % of these warps' instructions per block
    const int dtx = static_cast<int>(gridDim.x) % p.tiles_x, dq = static_cast<int>(gridDim.x) / p.tiles_x;
    const int dty = dq % p.tiles_y, dn = dq / p.tiles_y;
    int cb = team % p.ncb, n, ty, tx;
    tile_coords(blockIdx.x + (team / p.ncb) * gridDim.x, n, ty, tx);
    for (int g = team; g < total_g; g += 2) {
      const int ab = g & ((1 << p.na_shift) - 1);
      const int gy0 = ty * TH * S - 1, gx0 = tx * XD_TW * S - 1;
      // ---- expansion epilogue: TMEM -> stage buffer (every warp of the team is past its previous stencil: named barrier)
      mbar_wait(&e_full[team], (g >> 1) & 1);
      tc_fence_after();
      named_bar_sync(1 + team, 128);
      const float* bx = sBx + cb * BLOCK_K;
#pragma unroll 1
      for (int mb = 0; mb < NMB; ++mb) {
        if (mb * BLOCK_M + quad * 32 >= NPX) break;        // warp-uniform: this warp's rows of the M-block are padding
        const int px = mb * BLOCK_M + quad * 32 + lane;     // halo pixel of this thread's accumulator row
        const int iy = px / IW, ix = px - iy * IW;
        const bool in_tile = px < NPX;
        const bool valid = in_tile && (gy0 + iy) >= 0 && (gy0 + iy) < p.H && (gx0 + ix) >= 0 && (gx0 + ix) < p.W;
        const uint32_t vmask = valid ? 0xFFFFFFFFu : 0u;
        const bool all_valid = __all_sync(0xffffffffu, valid || !in_tile);   // interior warp: no masking instructions
        const uint32_t row_u32 = stage_u32 + px * 128;
        const uint32_t sw = static_cast<uint32_t>(ix) & 7u;
        uint32_t acc0[32], acc1[32];
        tmem_ld_32x32(tmem_base + lane_sel + team * EXP_COLS + mb * 64, acc0);
        tmem_ld_32x32(tmem_base + lane_sel + team * EXP_COLS + mb * 64 + 32, acc1);
        tmem_ld_wait_regs(acc0);
        tmem_ld_wait_regs(acc1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = *reinterpret_cast<const float4*>(bx + h * 32 + 4 * i);
            const float2 s0 = add2(make_float2(__uint_as_float(h ? acc1[4 * i + 0] : acc0[4 * i + 0]),
                                               __uint_as_float(h ? acc1[4 * i + 1] : acc0[4 * i + 1])), make_float2(b4.x, b4.y));
            const float2 s1 = add2(make_float2(__uint_as_float(h ? acc1[4 * i + 2] : acc0[4 * i + 2]),
                                               __uint_as_float(h ? acc1[4 * i + 3] : acc0[4 * i + 3])), make_float2(b4.z, b4.w));
            v[4 * i + 0] = s0.x; v[4 * i + 1] = s0.y; v[4 * i + 2] = s1.x; v[4 * i + 3] = s1.y;
          }
          uint32_t o[16];
          clamp_pack32(v, o, ex_relu, ex_cap, ex_hi2);
          if (!all_valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] &= vmask;
          }
          if (in_tile) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              sts128(row_u32 + (((static_cast<uint32_t>(h * 4 + c)) ^ sw) << 4), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
          }
        }
      }
      tc_fence_before();
      named_bar_sync(1 + team, 128);     // the whole expanded halo is in the stage buffer; the accumulator is free
      if (lane == 0) mbar_arrive(&e_empty[team]);
      // ---- depthwise stencil -> A operand of the projection
      const int ks = xd_ksteps(p.C, cb);
      const int cp = ks == 4 ? lane : (ks == 2 ? (lane & 15) : (lane & 7));
      const int row0 = quad * 4 + (ks == 4 ? 0 : (ks == 2 ? (lane >> 4) * 2 : (lane >> 3)));
      {
        const float2* wsm = reinterpret_cast<const float2*>(sW + cb * XD_WROW) + cp;
#pragma unroll
        for (int k = 0; k < 9; ++k) wr[k] = wsm[k * (BLOCK_K / 2)];
        b2 = wsm[9 * (BLOCK_K / 2)];
      }
      mbar_wait(&a_empty[ab], ((g >> p.na_shift) & 1) ^ 1);
      const uint32_t abase = smem_u32(sA) + ab * (BLOCK_M * 128);
      if (ks == 4) xd_stencil<S, TH, 4>(stage_u32, abase, row0, cp, wr, b2, dw_relu, dw_hi2);
      else if (ks == 2) xd_stencil<S, TH, 2>(stage_u32, abase, row0, cp, wr, b2, dw_relu, dw_hi2);
      else xd_stencil<S, TH, 1>(stage_u32, abase, row0, cp, wr, b2, dw_relu, dw_hi2);
      fence_proxy_async_smem();   // the A block is read by tcgen05.mma (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[ab]);
      cb += 2;
      while (cb >= p.ncb) {   // next tile(s) of this CTA
        cb -= p.ncb;
        tx += dtx;
        int carry = 0;
        if (tx >= p.tiles_x) { tx -= p.tiles_x; carry = 1; }
        ty += dty + carry;
        if (ty >= p.tiles_y) { ty -= p.tiles_y; n += 1; }
        n += dn;
      }
    }
  } else if (warp >= 4) {
    // ===================================== projection epilogue: bias (+ identity), clamp, 16-bit, direct stores =========
    const int q4 = warp & 3;
    const uint32_t lane_sel = static_cast<uint32_t>(q4 * 32) << 16;
    const uint32_t lo2 = pack_e16x2(p.pw_lo, p.pw_lo), hi2 = pack_e16x2(p.pw_hi, p.pw_hi);
    const int prow = q4 * 32 + lane;                      // accumulator row == output pixel of the tile
    const int ty_l = prow / XD_TW, tx_l = prow - ty_l * XD_TW;
    for (int it = 0; it < my_tiles; ++it) {
      const int buf = it & 1;
      int n, ty, tx;
      tile_coords(blockIdx.x + it * gridDim.x, n, ty, tx);
      const int oy = ty * TH + ty_l, ox = tx * XD_TW + tx_l;
      const bool ok = ty_l < TH && oy < p.Ho && ox < p.Wo;
      const size_t pix = (static_cast<size_t>(n) * p.Ho + oy) * p.Wo + ox;
      mbar_wait(&d_full[buf], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j * 32 < p.nmma; ++j) {
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + lane_sel + PROJ0 + buf * PROJ_COLS + j * 32, acc);
        const int ncol = min(32, p.Cout - j * 32);
        uint4 r4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          r4[c] = make_uint4(0u, 0u, 0u, 0u);
          if (p.res != nullptr && ok && 8 * c < ncol)
            r4[c] = __ldg(reinterpret_cast<const uint4*>(p.res + pix * p.res_pitch + j * 32) + c);
        }
        tmem_ld_wait_regs(acc);
        float v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = (4 * i < ncol) ? __ldg(reinterpret_cast<const float4*>(p.b_pw + j * 32) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 * i + 0] = __uint_as_float(acc[4 * i + 0]) + b4.x;
          v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + b4.y;
          v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + b4.z;
          v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + b4.w;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          v[8 * c + 0] += e16lo(r4[c].x); v[8 * c + 1] += e16hi(r4[c].x);
          v[8 * c + 2] += e16lo(r4[c].y); v[8 * c + 3] += e16hi(r4[c].y);
          v[8 * c + 4] += e16lo(r4[c].z); v[8 * c + 5] += e16hi(r4[c].z);
          v[8 * c + 6] += e16lo(r4[c].w); v[8 * c + 7] += e16hi(r4[c].w);
        }
        if (ok) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + pix * p.out_pitch + j * 32);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (8 * c < ncol) {
              uint4 o;
              o.x = hmin2_e16(hmax2_e16(pack_e16x2(v[8 * c + 0], v[8 * c + 1]), lo2), hi2);
              o.y = hmin2_e16(hmax2_e16(pack_e16x2(v[8 * c + 2], v[8 * c + 3]), lo2), hi2);
              o.z = hmin2_e16(hmax2_e16(pack_e16x2(v[8 * c + 4], v[8 * c + 5]), lo2), hi2);
              o.w = hmin2_e16(hmax2_e16(pack_e16x2(v[8 * c + 6], v[8 * c + 7]), lo2), hi2);
              dst[c] = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct XdOp : Op {
  CUtensorMap tmX, tmWx, tmB;
  XdParams p;
  int stride, grid, smem_bytes;
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    static std::atomic<uint64_t> done1{0}, done2{0};
    if (stride == 1) {
      if (cudaError_t e = set_max_smem_once(xdwpw_kernel<1>, XD_SMEM, done1)) return e;
      return launch_pdl(xdwpw_kernel<1>, dim3(grid), dim3(XD_THREADS), smem_bytes, s, tmX, tmWx, tmB, p);
    }
    if (cudaError_t e = set_max_smem_once(xdwpw_kernel<2>, XD_SMEM, done2)) return e;
    return launch_pdl(xdwpw_kernel<2>, dim3(grid), dim3(XD_THREADS), smem_bytes, s, tmX, tmWx, tmB, p);
  }
};

struct XdGeom {
  int S, TH, IH, IW, npx, nmb, ncb, nmma, halo_bytes, x_bytes, b_bytes, stages, xbufs, na_shift, smem, Ho, Wo;
};

static bool xd_clampish(int act) { return act == PCV_ACT_NONE || act == PCV_ACT_RELU || act == PCV_ACT_RELU6; }

static bool xd_geom(const pcv_conv_desc& ex, const pcv_conv_desc& dw, const pcv_conv_desc& pw, XdGeom* g) {
  if (ex.kh != 1 || ex.kw != 1 || ex.stride != 1 || ex.pad != 0 || ex.groups != 1 || ex.Cin > 64 || ex.Cin % 8 != 0 ||
      ex.Cout != dw.Cin || !xd_clampish(ex.act) || ex.flags != 0 || ex.in_row_pitch != 0 || ex.N != dw.N || ex.H != dw.H ||
      ex.W != dw.W || pitch_or(ex.in_pitch, ex.Cin) % 8 != 0)
    return false;
  const bool depthwise = dw.groups == dw.Cin && dw.Cin == dw.Cout;
  if (!depthwise || dw.kh != 3 || dw.kw != 3 || dw.pad != 1 || dw.dil != 1 || (dw.stride != 1 && dw.stride != 2) ||
      dw.Cin % 8 != 0 || !xd_clampish(dw.act) || dw.flags != 0 || dw.in_row_pitch != 0)
    return false;
  if (pw.kh != 1 || pw.kw != 1 || pw.stride != 1 || pw.pad != 0 || pw.groups != 1 || pw.Cin != dw.Cout || pw.Cout % 8 != 0 ||
      !xd_clampish(pw.act) || pw.flags != 0 || pw.in_row_pitch != 0)
    return false;
  g->S = dw.stride;
  g->TH = g->S == 1 ? 8 : 4;
  g->Ho = conv_out(dw.H, 3, dw.stride, 1, 1);
  g->Wo = conv_out(dw.W, 3, dw.stride, 1, 1);
  if (pw.N != dw.N || pw.H != g->Ho || pw.W != g->Wo || g->Wo < 14 || g->Ho < g->TH) return false;
  if (pitch_or(pw.out_pitch, pw.Cout) % 8 || pitch_or(pw.res_pitch, pw.Cout) % 8) return false;
  g->IH = (g->TH - 1) * g->S + 3;
  g->IW = (XD_TW - 1) * g->S + 3;
  g->npx = g->IH * g->IW;
  g->nmb = ceil_div(g->npx, BLOCK_M);
  g->ncb = ceil_div(dw.Cin, BLOCK_K);
  g->nmma = round_up(pw.Cout, 16);
  if (g->nmma > (512 - 2 * g->nmb * 64) / 2) return false;   // two projection accumulators next to two expansion accumulators
  if (((g->ncb - 1) * BLOCK_K + xd_ksteps(dw.Cin, g->ncb - 1) * 16) * 2 > dw.Cin * 3) return false;   // mostly padding
  g->halo_bytes = round_up(g->npx * 128, 1024);
  g->x_bytes = g->nmb * BLOCK_M * 128;
  g->b_bytes = round_up(g->nmma * 128, 1024);
  const int fixed = 1024 + g->ncb * (8192 + g->b_bytes + XD_WROW * 4 + BLOCK_K * 4) + 512;
  // one stage buffer per stencil team; preference: two input-halo buffers, then four A operand buffers
  for (int xbufs = 2; xbufs >= 1; --xbufs)
    for (int na_shift = 2; na_shift >= 1; --na_shift) {
      const int smem = fixed + xbufs * g->x_bytes + 2 * g->halo_bytes + (BLOCK_M * 128 << na_shift);
      if (smem <= XD_SMEM) {
        g->stages = 2; g->xbufs = xbufs; g->na_shift = na_shift; g->smem = smem;
        return true;
      }
    }
  return false;
}

int xdwpw_ok(const pcv_conv_desc& ex, const pcv_conv_desc& dw, const pcv_conv_desc& pw) {
  XdGeom g;
  return xd_geom(ex, dw, pw, &g) ? 1 : 0;
}

int xdwpw_make(const pcv_conv_desc& ex, const pcv_conv_desc& dw, const pcv_conv_desc& pw, const void* x, const void* w_ex,
               const float* b_ex, const float* w_dw, const float* b_dw, const void* w_pw, const float* b_pw, const void* res,
               void* y, Op** out) {
  XdGeom g;
  if (!xd_geom(ex, dw, pw, &g)) return fail(PCV_ERR_UNSUPPORTED, "expand -> depthwise -> project triple outside the fused kernel's domain");
  PCV_REQUIRE(x && w_ex && b_ex && w_dw && b_dw && w_pw && b_pw && y, "NULL tensor pointer");
  for (const void* ptr : {x, w_ex, w_pw, res, static_cast<const void*>(y), static_cast<const void*>(w_dw), static_cast<const void*>(b_dw),
                          static_cast<const void*>(b_ex), static_cast<const void*>(b_pw)})
    PCV_REQUIRE(reinterpret_cast<uintptr_t>(ptr) % 16 == 0, "fused expand -> dw -> pw operands must be 16-byte aligned");
  const int in_pitch = pitch_or(ex.in_pitch, ex.Cin);
  auto op = std::make_unique<XdOp>();
  XdParams& p = op->p;
  p.w_dw = w_dw; p.b_dw = b_dw; p.b_exp = b_ex; p.b_pw = b_pw;
  p.res = reinterpret_cast<const e16*>(res);
  p.out = reinterpret_cast<e16*>(y);
  p.N = dw.N; p.H = dw.H; p.W = dw.W; p.Cin = ex.Cin; p.C = dw.Cin; p.Ho = g.Ho; p.Wo = g.Wo; p.Cout = pw.Cout;
  p.kx = ceil_div(ex.Cin, 16);
  p.nmma = g.nmma;
  p.idesc_pw = make_idesc_e16(BLOCK_M, g.nmma);
  p.out_pitch = pitch_or(pw.out_pitch, pw.Cout);
  p.res_pitch = pitch_or(pw.res_pitch, pw.Cout);
  p.tiles_x = ceil_div(g.Wo, XD_TW);
  p.tiles_y = ceil_div(g.Ho, g.TH);
  const long long tiles = static_cast<long long>(dw.N) * p.tiles_x * p.tiles_y;
  PCV_REQUIRE(tiles < (1ll << 30), "too many tiles");
  p.num_tiles = static_cast<int>(tiles);
  p.ncb = g.ncb;
  p.stages = g.stages; p.halo_bytes = g.halo_bytes; p.xbufs = g.xbufs; p.x_bytes = g.x_bytes; p.b_bytes = g.b_bytes;
  p.na_shift = g.na_shift;
  auto lohi = [](int act, float* lo, float* hi) {
    *lo = (act == PCV_ACT_RELU || act == PCV_ACT_RELU6) ? 0.f : -INFINITY;
    *hi = act == PCV_ACT_RELU6 ? 6.f : INFINITY;
  };
  lohi(ex.act, &p.ex_lo, &p.ex_hi);
  lohi(dw.act, &p.dw_lo, &p.dw_hi);
  lohi(pw.act, &p.pw_lo, &p.pw_hi);
  op->stride = g.S;
  op->smem_bytes = g.smem;
  op->grid = std::min(p.num_tiles, sm_count());
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  {
    // input halo: 64-channel box (channels beyond Cin are out of bounds: zero-filled, never fetched) -> 128-byte K-major rows
    cuuint64_t dims[4] = {(cuuint64_t)ex.Cin, (cuuint64_t)dw.W, (cuuint64_t)dw.H, (cuuint64_t)dw.N};
    cuuint64_t strides[3] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)dw.W * in_pitch * 2, (cuuint64_t)dw.H * dw.W * in_pitch * 2};
    cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)g.IW, (cuuint32_t)g.IH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&op->tmX, TMAP_E16, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled (expand -> dw -> pw input halo) failed (%d)", (int)r);
  }
  // expansion weights: packed [C, 64] K-major (igemm_pack of the 1x1 expansion; Cin <= 64 is one k-block)
  if (int rc = make_tiled_2d(&op->tmWx, w_ex, BLOCK_K, dw.Cin, BLOCK_K * 2, BLOCK_K, BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  const uint64_t kpad = (uint64_t)g.ncb * BLOCK_K;
  if (int rc = make_tiled_2d(&op->tmB, w_pw, kpad, pw.Cout, kpad * 2, BLOCK_K, g.nmma, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  char nm[200];
  snprintf(nm, sizeof nm, "conv_xdwpw fused 1x1 %d->%d -> dw3x3 s%d -> 1x1 %d->%d @%dx%d%s st=%d xb=%d na=%d", ex.Cin, dw.Cin, g.S,
           dw.Cin, pw.Cout, dw.H, dw.W, res ? " +res" : "", g.stages, g.xbufs, 1 << g.na_shift);
  op->name = nm;
  const double Mi = static_cast<double>(dw.N) * dw.H * dw.W, Mo = static_cast<double>(dw.N) * g.Ho * g.Wo;
  op->flops = 2.0 * Mi * ex.Cin * dw.Cin + 2.0 * Mo * dw.Cin * 9 + 2.0 * Mo * pw.Cout * dw.Cin;
  // fused group: x in, y out (+ identity), the three weight sets; neither the expanded nor the depthwise tensor
  // touches HBM
  op->bytes = 2.0 * Mi * ex.Cin + 2.0 * Mo * pw.Cout * (res ? 2.0 : 1.0) + 2.0 * ex.Cin * dw.Cin + 4.0 * dw.Cin * 11 +
              2.0 * dw.Cin * pw.Cout + 4.0 * pw.Cout;
  *out = op.release();
  return PCV_OK;
}

}  // namespace PCV_TIER
}  // namespace pcv
