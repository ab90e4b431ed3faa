// Fused depthwise 3x3 -> pointwise 1x1 (DwsConvBlock.forward, conv.py:605-608; LinearBottleneck conv2 -> conv3 + identity,
// mobilenetv2.py:52-71): the depthwise result - the widest tensor of an inverted-residual block - never goes to HBM.
//
//   y = act_pw( W_pw * act_dw(dw3x3(x) + b_dw) + b_pw  [+ residual] )
//
// A tile is 8 x 16 output pixels (= the 128 rows of one tcgen05 MMA) of one image.  Per 64-channel block of the depthwise
// width the producer warp fetches the input halo box ((8-1)S+3 rows x (16-1)S+3 columns x 64 channels, ONE 4-D TMA load,
// borders zero-filled by the TMA unit) together with that block's slice of the pointwise weights; eight CUDA-core warps run
// the depthwise stencil (fp32 FFMA2, weights in registers, a thread owns 2 channels of two adjacent output columns and
// walks down the rows) and write the activated 16-bit result straight into shared memory IN THE K-MAJOR SWIZZLE_128B LAYOUT OF AN MMA
// A OPERAND; the MMA warp multiplies it by the weight slice and accumulates over the channel blocks in TMEM; four epilogue
// warps add bias (+ the unit's identity), clamp and store.  Depthwise compute of block c+1 overlaps the MMA of block c
// (two A buffers) and the epilogue of tile t overlaps tile t+1 (two TMEM accumulators).
//
// A TAIL channel block with <= 32 (<= 16) valid channels runs as a 32- (16-) channel block: the stencil warps re-map their
// lanes onto fewer channels and more columns (dp_stencil<S, 2 / 1>) and the MMA issues 2 (1) K = 16 steps, so C = 96 / 144 pay
// for 96 / 144 channels of stencil work instead of 128 / 192 (MobileNetV2 bs256: 146.7 k -> 148.9 k img/s) and C = 32 fuses.
//
// Saved per block: one write + one read of the depthwise tensor (MobileNetV2 bs256: 2.4 GB of the step's 7.3 GB) and one
// kernel launch.  Domain: 3x3 depthwise, stride 1 or 2, pad 1, clamp-family activations, Cout <= 256, maps >= 14 wide.
#define PCV_MBAR_SUSPEND_NS 20000u   // idle role warps sleep in mbarrier.try_wait instead of re-polling (ptx.cuh)
#include "igemm_common.cuh"

namespace pcv {
namespace PCV_TIER {

constexpr int DP_TH = 8, DP_TW = 16;
constexpr int DP_THREADS = 512;   // warps: 0 producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4-7 epilogue, 8-15 depthwise
constexpr int DP_MAX_STAGES = 6;
constexpr int DP_NA = 4;           // A operand buffers: two per stencil team (one each when shared memory is short)
constexpr int DP_SMEM = 232448 - 1024;   // dynamic shared memory budget (the rest: the static debug timeline)
constexpr int DP_WROW = 10 * BLOCK_K;   // floats of depthwise weights (9 taps + bias) per 64-channel block

struct DwPwParams {
  const float* w_dw;     // [9][C] fp32, BN folded
  const float* b_dw;     // [C]
  const float* b_pw;     // [>= Cout]
  const e16* res;
  e16* out;
  int N, H, W, C, Ho, Wo, Cout;
  int nmma;              // MMA N: Cout rounded up to 16
  uint32_t idesc;
  int out_pitch, res_pitch;
  int tiles_x, tiles_y, num_tiles, ncb;
  int stages, stage_bytes, halo_bytes, halo_tx, b_tx, b_bytes;
  int na_shift;          // log2(A operand buffers): 2, or 1 when the halos leave no room for four
  int b_res;             // the whole pointwise weight matrix stays in shared memory (else one slice rides with each halo)
  float dw_lo, dw_hi, pw_lo, pw_hi;
  int dbg;               // PCV_DP_DBG timing experiments: 1 no stencil, 2 no halo loads, 4 no epilogue traffic, 8 no MMA
};

__device__ __forceinline__ void dp_tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void dp_ffma2(float2& d, const float2 a, const float2 b) {
  uint64_t dd = (static_cast<uint64_t>(__float_as_uint(d.y)) << 32) | __float_as_uint(d.x);
  const uint64_t aa = (static_cast<uint64_t>(__float_as_uint(a.y)) << 32) | __float_as_uint(a.x);
  const uint64_t bb = (static_cast<uint64_t>(__float_as_uint(b.y)) << 32) | __float_as_uint(b.x);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d.x = __uint_as_float(static_cast<uint32_t>(dd));
  d.y = __uint_as_float(static_cast<uint32_t>(dd >> 32));
}
__device__ __forceinline__ uint2 dp_lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dp_sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint32_t dp_lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dp_sts32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t dp_hclamp2(uint32_t v, uint32_t lo, uint32_t hi) {
  return hmin2_e16(hmax2_e16(v, lo), hi);
}

// The depthwise stencil of one (tile, 64-channel block) for the NQ adjacent output columns row0 .. row0 + NQ - 1 of the
// tile that this thread owns, channels 2 cp and 2 cp + 1 of the block: walks down the IH input rows with the next row's loads in
// flight, 3 S + 3 ... (NQ - 1) S + 3 input columns per row, NQ x NACC independent FFMA2 chains, and writes the activated 16-bit
// results into rows (ho * 16 + row0 + q) of the A operand (16-byte chunks XOR-swizzled by row & 7, independent of ho).
template <int S, int NQ>
__device__ __forceinline__ void dp_stencil(const uint32_t stage_u32, const uint32_t abase, const int row0, const int cp,
                                           const float2 (&wr)[9], const float2 b2, const bool dw_relu, const uint32_t dw_hi2) {
  constexpr int IH = (DP_TH - 1) * S + 3, IW = (DP_TW - 1) * S + 3;
  constexpr int NACC = (3 + S - 1) / S;
  constexpr int NJ = (NQ - 1) * S + 3;
  const uint32_t sbase = stage_u32 + (row0 * S * 64 + cp * 2) * 2;
  uint32_t a_off[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const uint32_t r = static_cast<uint32_t>(row0 + q);
    a_off[q] = r * 128 + (((static_cast<uint32_t>(cp) >> 2) ^ (r & 7u)) << 4) + (static_cast<uint32_t>(cp) & 3u) * 4;
  }
  float2 acc[NACC][NQ];
  uint32_t raw[2][NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) raw[0][j] = dp_lds32(sbase + j * 128);
#pragma unroll
  for (int ir = 0; ir < IH; ++ir) {
    if (ir + 1 < IH) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) raw[(ir + 1) & 1][j] = dp_lds32(sbase + ((ir + 1) * IW + j) * 128);
    }
    float2 xv[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) xv[j] = make_float2(e16lo(raw[ir & 1][j]), e16hi(raw[ir & 1][j]));
#pragma unroll
    for (int fr = 0; fr < 3; ++fr) {
      if ((ir - fr) < 0 || (ir - fr) % S != 0 || (ir - fr) / S >= DP_TH) continue;   // compile-time after unrolling
      const int ho = (ir - fr) / S;
      const int a = ho % NACC;
      if (fr == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[a][q] = b2;
      }
#pragma unroll
      for (int fs = 0; fs < 3; ++fs) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) dp_ffma2(acc[a][q], xv[q * S + fs], wr[fr * 3 + fs]);
      }
      if (fr == 2) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const uint32_t o = dw_relu ? pack_relu_e16x2(acc[a][q].x, acc[a][q].y) : pack_e16x2(acc[a][q].x, acc[a][q].y);
          dp_sts32(abase + ho * (DP_TW * 128) + a_off[q], hmin2_e16(o, dw_hi2));
        }
      }
    }
  }
}

// K = 16 MMA steps a channel block needs: a TAIL block with <= 32 (<= 16) valid channels is computed and multiplied as a
// 32- (16-) channel block - the stencil warps re-map their lanes onto fewer channels and more columns (dp_stencil<S, 2 / 1>).
__host__ __device__ __forceinline__ int dp_ksteps(int C, int cb) {
  const int cvalid = C - cb * BLOCK_K;
  return cvalid > 32 ? 4 : (cvalid > 16 ? 2 : 1);
}

template <int S>
__global__ void __launch_bounds__(DP_THREADS, 1)
dwpw_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmB, const DwPwParams p) {
  constexpr int IH = (DP_TH - 1) * S + 3, IW = (DP_TW - 1) * S + 3;
  constexpr int NACC = (3 + S - 1) / S;
  __shared__ long long tlog[64];   // PCV_DP_DBG & 16: clock64 timeline of iterations 24..31 of CTA 0
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sStage = smem;                                        // stages x [halo | B slice (1024-aligned)]
  uint8_t* sA = sStage + p.stages * p.stage_bytes;               // DP_NA x [128 x 128 B] A operand
  uint8_t* sB = sA + (BLOCK_M * 128 << p.na_shift);                          // resident pointwise weights: ncb x [nmma x 128 B]
  float* sW = reinterpret_cast<float*>(sB + (p.b_res ? p.ncb * p.b_bytes : 0));   // [ncb][10][64] depthwise weights + bias
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + p.ncb * DP_WROW);
  uint64_t* full = bars;                          // [stages] halo + B landed
  uint64_t* empty = full + DP_MAX_STAGES;         // [stages] 8 depthwise warps + 1 MMA commit
  uint64_t* a_full = empty + DP_MAX_STAGES;       // [DP_NA] the 4 warps of a team
  uint64_t* a_empty = a_full + DP_NA;             // [DP_NA] MMA commit
  uint64_t* d_full = a_empty + DP_NA;             // [2] MMA commit
  uint64_t* d_empty = d_full + 2;                 // [2] 4 epilogue warps
  uint64_t* b_full = d_empty + 2;                 // resident pointwise weights landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int my_tiles = blockIdx.x < p.num_tiles ? (p.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], p.b_res ? 4 : 5);
    }
    for (int i = 0; i < DP_NA; ++i) {
      mbar_init(&a_full[i], 4);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 1);
      mbar_init(&d_empty[i], 4);
    }
    mbar_init(b_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  // depthwise weights + bias of every channel block, zero beyond C (constants: may be read before the PDL wait)
  for (int i = threadIdx.x; i < p.ncb * DP_WROW; i += DP_THREADS) {
    const int cbk = i / BLOCK_K, c = (cbk / 10) * BLOCK_K + (i - cbk * BLOCK_K), k = cbk % 10;
    sW[i] = c < p.C ? __ldg(k < 9 ? p.w_dw + static_cast<size_t>(k) * p.C + c : p.b_dw + c) : 0.f;
  }
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
    // ===================================== producer: one (halo box, weight slice) per tile and channel block ===========
    int g = 0;
    if (p.b_res && my_tiles > 0 && elect_one()) {
      mbar_arrive_expect_tx(b_full, p.ncb * p.b_tx);
      for (int cb = 0; cb < p.ncb; ++cb) tma_load_2d(&tmB, b_full, sB + cb * p.b_bytes, cb * BLOCK_K, 0);
    }
    __syncwarp();
    for (int it = 0; it < my_tiles; ++it) {
      int n, ty, tx;
      tile_coords(blockIdx.x + it * gridDim.x, n, ty, tx);
      for (int cb = 0; cb < p.ncb; ++cb, ++g) {
        const int st = g % p.stages;
        mbar_wait(&empty[st], ((g / p.stages) & 1) ^ 1);
        if (elect_one()) {
          uint8_t* dst = sStage + st * p.stage_bytes;
          mbar_arrive_expect_tx(&full[st], ((p.dbg & 2) ? 0 : p.halo_tx) + (p.b_res ? 0 : p.b_tx));
          if (!(p.dbg & 2)) dp_tma_load_4d(&tmIn, &full[st], dst, cb * BLOCK_K, tx * DP_TW * S - 1, ty * DP_TH * S - 1, n);
          if (!p.b_res) tma_load_2d(&tmB, &full[st], dst + p.halo_bytes, cb * BLOCK_K, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ====================================================================
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA));
    int g = 0;
    if (p.b_res && my_tiles > 0) mbar_wait(b_full, 0);
    for (int it = 0; it < my_tiles; ++it) {
      const int buf = it & 1;
      mbar_wait(&d_empty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * 256;
      for (int cb = 0; cb < p.ncb; ++cb, ++g) {
        const int st = g % p.stages, ab = g & ((1 << p.na_shift) - 1);
        const bool tl = (p.dbg & 16) && blockIdx.x == 0 && lane == 0 && g >= 24 && g < 32;
        long long tq0 = 0;
        if (tl) tq0 = clock64();
        if (!p.b_res) mbar_wait(&full[st], (g / p.stages) & 1);   // the weight slice shares the stage's barrier with the halo
        mbar_wait(&a_full[ab], (g >> p.na_shift) & 1);
        if (tl) {
          tlog[48 + (g - 24) * 2] = tq0;
          tlog[48 + (g - 24) * 2 + 1] = clock64();
        }
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + ab * (BLOCK_M * 128 >> 4);
        const uint32_t b_lo = smem_desc_lo(smem_u32(p.b_res ? sB + cb * p.b_bytes : sStage + st * p.stage_bytes + p.halo_bytes));
        const int ks = dp_ksteps(p.C, cb);
        if (elect_one()) {
          if (!(p.dbg & 8))
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            if (k < ks) umma_bf16_lohi(d_tmem, a_lo + 2 * k, b_lo + 2 * k, p.idesc, (cb | k) != 0 ? 1u : 0u);
          umma_commit(&a_empty[ab]);
          if (!p.b_res) umma_commit(&empty[st]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&d_full[buf]);
      __syncwarp();
    }
  } else if (warp >= 8) {
    // ===================================== depthwise stencil on CUDA cores ==============================================
    // Two TEAMS of four warps take alternate channel blocks (team = g & 1) and run out of step with each other, so one
    // team's barrier round trips and weight reloads hide under the other's arithmetic; each team owns two of the four A
    // buffers.  Inside a team a warp owns four adjacent output columns x all 64 channels and a thread two channels of
    // them, walking down the rows: 3 S + 3 input columns per row feed the four outputs, the next row's loads are issued
    // before this row's arithmetic, and the columns (times the rows in flight) are independent FFMA2 chains.  Every
    // shared-memory access of a warp is one contiguous 128-byte line.
    const int team = (warp - 8) >> 2;
    const int quad = (warp - 8) & 3;     // output columns 4 quad .. 4 quad + 3 of the tile
    const uint32_t dw_hi2 = pack_e16x2(p.dw_hi, p.dw_hi);
    const bool dw_relu = p.dw_lo == 0.f;
    const int total_g = my_tiles * p.ncb;
    float2 wr[9], b2;
    // ring position / phase and channel block advance incrementally (g += 2): a division per block and warp was 8 % of the
    // stencil warps' instructions (ncu source page, profiles/README.md)
    int st = team % p.stages, st_phase = (team / p.stages) & 1, cb = team % p.ncb;
    for (int g = team; g < total_g; g += 2) {
      const int ab = g & ((1 << p.na_shift) - 1);
      // lane -> (channel pair, column group): a full block gives a thread 2 channels of the warp's 4 columns; a tail block with
      // <= 32 (<= 16) valid channels gives it 2 (1) of those columns, so that no lane works on padding channels
      const int ks = dp_ksteps(p.C, cb);
      const int cp = ks == 4 ? lane : (ks == 2 ? (lane & 15) : (lane & 7));     // channels 2 cp, 2 cp + 1 of the block
      const int row0 = quad * 4 + (ks == 4 ? 0 : (ks == 2 ? (lane >> 4) * 2 : (lane >> 3)));
      // this block's depthwise weights: 9 taps x 2 channels (zero beyond C: the TMA zero-fills those channels too)
      if (g == team || p.ncb > 1) {
        const float2* wsm = reinterpret_cast<const float2*>(sW + cb * DP_WROW) + cp;
#pragma unroll
        for (int k = 0; k < 9; ++k) wr[k] = wsm[k * (BLOCK_K / 2)];
        b2 = wsm[9 * (BLOCK_K / 2)];
      }
      const bool tl = (p.dbg & 16) && blockIdx.x == 0 && quad == 0 && lane == 0 && g >= 24 && g < 32;
      long long tq0 = 0, tq1 = 0, tq2 = 0, tq3 = 0;
      if (tl) tq0 = clock64();
      mbar_wait(&full[st], st_phase);
      if (tl) tq1 = clock64();
      mbar_wait(&a_empty[ab], ((g >> p.na_shift) & 1) ^ 1);
      if (tl) tq2 = clock64();
      const uint32_t stage_u32 = smem_u32(sStage + st * p.stage_bytes);
      const uint32_t abase = smem_u32(sA) + ab * (BLOCK_M * 128);
      if (!(p.dbg & 1)) {
        if (ks == 4) dp_stencil<S, 4>(stage_u32, abase, row0, cp, wr, b2, dw_relu, dw_hi2);
        else if (ks == 2) dp_stencil<S, 2>(stage_u32, abase, row0, cp, wr, b2, dw_relu, dw_hi2);
        else dp_stencil<S, 1>(stage_u32, abase, row0, cp, wr, b2, dw_relu, dw_hi2);
      }
      if (tl) tq3 = clock64();
      fence_proxy_async_smem();   // the A block is read by tcgen05.mma (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&a_full[ab]);
        mbar_arrive(&empty[st]);
      }
      if (tl) {
        long long* L = tlog + (g - 24) * 6;
        L[0] = tq0; L[1] = tq1; L[2] = tq2; L[3] = tq3; L[4] = clock64();
      }
      st += 2;
      if (st >= p.stages) { st -= p.stages; st_phase ^= 1; }
      cb += 2;
      if (cb >= p.ncb) cb -= p.ncb;
      if (cb >= p.ncb) cb -= p.ncb;
    }
  } else if (warp >= 4) {
    // ===================================== epilogue: bias (+ identity), clamp, 16-bit, direct stores ====================
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const int ty_l = row / DP_TW, tx_l = row - ty_l * DP_TW;
    const uint32_t lo2 = pack_e16x2(p.pw_lo, p.pw_lo), hi2 = pack_e16x2(p.pw_hi, p.pw_hi);
    for (int it = 0; it < my_tiles; ++it) {
      const int buf = it & 1;
      int n, ty, tx;
      tile_coords(blockIdx.x + it * gridDim.x, n, ty, tx);
      const int oy = ty * DP_TH + ty_l, ox = tx * DP_TW + tx_l;
      const bool ok = oy < p.Ho && ox < p.Wo;
      const size_t pix = (static_cast<size_t>(n) * p.Ho + oy) * p.Wo + ox;
      mbar_wait(&d_full[buf], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j * 32 < ((p.dbg & 4) ? 0 : p.nmma); ++j) {
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + buf * 256 + j * 32, acc);
        const int ncol = min(32, p.Cout - j * 32);          // columns of this chunk that exist (multiple of 8)
        float4 b4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          b4[i] = (4 * i < ncol) ? __ldg(reinterpret_cast<const float4*>(p.b_pw + j * 32) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
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
          v[4 * i + 0] = __uint_as_float(acc[4 * i + 0]) + b4[i].x;
          v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + b4[i].y;
          v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + b4[i].z;
          v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + b4[i].w;
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
              o.x = dp_hclamp2(pack_e16x2(v[8 * c + 0], v[8 * c + 1]), lo2, hi2);
              o.y = dp_hclamp2(pack_e16x2(v[8 * c + 2], v[8 * c + 3]), lo2, hi2);
              o.z = dp_hclamp2(pack_e16x2(v[8 * c + 4], v[8 * c + 5]), lo2, hi2);
              o.w = dp_hclamp2(pack_e16x2(v[8 * c + 6], v[8 * c + 7]), lo2, hi2);
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
  if ((p.dbg & 16) && blockIdx.x == 0 && threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i)
      printf("g=%d dw: top %lld full+%lld aempty+%lld stencil+%lld arrive+%lld | mma: top %lld afull+%lld\n", 24 + i,
             tlog[i * 6] - tlog[0], tlog[i * 6 + 1] - tlog[i * 6], tlog[i * 6 + 2] - tlog[i * 6 + 1],
             tlog[i * 6 + 3] - tlog[i * 6 + 2], tlog[i * 6 + 4] - tlog[i * 6 + 3], tlog[48 + 2 * i] - tlog[0],
             tlog[48 + 2 * i + 1] - tlog[48 + 2 * i]);
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct DwPwOp : Op {
  CUtensorMap tmIn, tmB;
  DwPwParams p;
  int stride, grid, smem_bytes;
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    static std::atomic<uint64_t> done1{0}, done2{0};
    if (stride == 1) {
      if (cudaError_t e = set_max_smem_once(dwpw_kernel<1>, DP_SMEM, done1)) return e;
      return launch_pdl(dwpw_kernel<1>, dim3(grid), dim3(DP_THREADS), smem_bytes, s, tmIn, tmB, p);
    }
    if (cudaError_t e = set_max_smem_once(dwpw_kernel<2>, DP_SMEM, done2)) return e;
    return launch_pdl(dwpw_kernel<2>, dim3(grid), dim3(DP_THREADS), smem_bytes, s, tmIn, tmB, p);
  }
};

struct DwPwGeom {
  int S, IH, IW, ncb, nmma, halo_bytes, b_bytes, stage_bytes, stages, smem, Ho, Wo, b_res, na_shift;
};

static bool clampish(int act) { return act == PCV_ACT_NONE || act == PCV_ACT_RELU || act == PCV_ACT_RELU6; }

static bool dwpw_geom(const pcv_conv_desc& dw, const pcv_conv_desc& pw, DwPwGeom* g) {
  const bool depthwise = dw.groups == dw.Cin && dw.Cin == dw.Cout;
  if (!depthwise || dw.kh != 3 || dw.kw != 3 || dw.pad != 1 || dw.dil != 1 || (dw.stride != 1 && dw.stride != 2) ||
      dw.Cin % 8 != 0 || !clampish(dw.act) || dw.flags != 0 || dw.in_row_pitch != 0)
    return false;
  if (pw.kh != 1 || pw.kw != 1 || pw.stride != 1 || pw.pad != 0 || pw.groups != 1 || pw.Cin != dw.Cout || pw.Cout % 8 != 0 ||
      pw.Cout > 256 || !clampish(pw.act) || pw.flags != 0 || pw.in_row_pitch != 0)
    return false;
  g->S = dw.stride;
  g->Ho = conv_out(dw.H, 3, dw.stride, 1, 1);
  g->Wo = conv_out(dw.W, 3, dw.stride, 1, 1);
  if (pw.N != dw.N || pw.H != g->Ho || pw.W != g->Wo || g->Wo < 14 || g->Ho < 8) return false;
  if (pitch_or(dw.in_pitch, dw.Cin) % 8 || pitch_or(pw.out_pitch, pw.Cout) % 8 || pitch_or(pw.res_pitch, pw.Cout) % 8) return false;
  g->IH = (DP_TH - 1) * g->S + 3;
  g->IW = (DP_TW - 1) * g->S + 3;
  g->ncb = ceil_div(dw.Cin, BLOCK_K);
  g->nmma = round_up(pw.Cout, 16);
  g->halo_bytes = round_up(g->IH * g->IW * 128, 1024);
  g->b_bytes = round_up(g->nmma * 128, 1024);
  // a channel block that is mostly padding wastes the stencil's arithmetic; tail blocks run as 32- / 16-channel blocks
  // (dp_ksteps), so what counts is the padding that is left after that (C = 24: 32 channels of work for 24)
  if (((g->ncb - 1) * BLOCK_K + dp_ksteps(dw.Cin, g->ncb - 1) * 16) * 2 > dw.Cin * 3) return false;
  for (g->na_shift = 2; g->na_shift >= 1; --g->na_shift) {
    const int fixed = 1024 + (BLOCK_M * 128 << g->na_shift) + g->ncb * DP_WROW * 4 + 512;
    // the pointwise weights either stay resident (a stage is a bare halo, released by the stencil warps alone) or ride
    // slice by slice with the halos (the stage is held until its MMA retires): resident when that costs no pipeline depth
    const int st_res = std::min(DP_MAX_STAGES, (DP_SMEM - fixed - g->ncb * g->b_bytes) / g->halo_bytes);
    const int st_rid = std::min(DP_MAX_STAGES, (DP_SMEM - fixed) / (g->halo_bytes + g->b_bytes));
    g->b_res = (st_res >= 2 && (st_res >= st_rid || st_res >= 4)) ? 1 : 0;
    g->stages = g->b_res ? st_res : st_rid;
    g->stage_bytes = g->halo_bytes + (g->b_res ? 0 : g->b_bytes);
    g->smem = fixed + g->stages * g->stage_bytes + (g->b_res ? g->ncb * g->b_bytes : 0);
    if (g->stages >= 2) return true;
  }
  return false;
}

int dwpw_ok(const pcv_conv_desc& dw, const pcv_conv_desc& pw) {
  DwPwGeom g;
  return dwpw_geom(dw, pw, &g) ? 1 : 0;
}

int dwpw_make(const pcv_conv_desc& dw, const pcv_conv_desc& pw, const void* x, const float* w_dw, const float* b_dw,
              const void* w_pw, const float* b_pw, const void* res, void* y, Op** out) {
  DwPwGeom g;
  if (!dwpw_geom(dw, pw, &g)) return fail(PCV_ERR_UNSUPPORTED, "depthwise -> pointwise pair outside the fused kernel's domain");
  PCV_REQUIRE(x && w_dw && b_dw && w_pw && b_pw && y, "NULL tensor pointer");
  for (const void* ptr : {x, w_pw, res, static_cast<const void*>(y), static_cast<const void*>(w_dw), static_cast<const void*>(b_dw)})
    PCV_REQUIRE(reinterpret_cast<uintptr_t>(ptr) % 16 == 0, "fused dw -> pw operands must be 16-byte aligned");
  const int in_pitch = pitch_or(dw.in_pitch, dw.Cin);
  auto op = std::make_unique<DwPwOp>();
  DwPwParams& p = op->p;
  p.w_dw = w_dw; p.b_dw = b_dw; p.b_pw = b_pw;
  p.res = reinterpret_cast<const e16*>(res);
  p.out = reinterpret_cast<e16*>(y);
  p.N = dw.N; p.H = dw.H; p.W = dw.W; p.C = dw.Cin; p.Ho = g.Ho; p.Wo = g.Wo; p.Cout = pw.Cout;
  p.nmma = g.nmma;
  p.idesc = make_idesc_e16(BLOCK_M, g.nmma);
  p.out_pitch = pitch_or(pw.out_pitch, pw.Cout);
  p.res_pitch = pitch_or(pw.res_pitch, pw.Cout);
  p.tiles_x = ceil_div(g.Wo, DP_TW);
  p.tiles_y = ceil_div(g.Ho, DP_TH);
  const long long tiles = static_cast<long long>(dw.N) * p.tiles_x * p.tiles_y;
  PCV_REQUIRE(tiles < (1ll << 30), "too many tiles");
  p.num_tiles = static_cast<int>(tiles);
  p.ncb = g.ncb;
  p.stages = g.stages; p.stage_bytes = g.stage_bytes; p.halo_bytes = g.halo_bytes;
  p.halo_tx = g.IH * g.IW * 128;
  p.b_tx = g.nmma * 128;
  p.b_bytes = g.b_bytes; p.b_res = g.b_res; p.na_shift = g.na_shift;
  auto lohi = [](int act, float* lo, float* hi) {
    *lo = (act == PCV_ACT_RELU || act == PCV_ACT_RELU6) ? 0.f : -INFINITY;
    *hi = act == PCV_ACT_RELU6 ? 6.f : INFINITY;
  };
  lohi(dw.act, &p.dw_lo, &p.dw_hi);
  lohi(pw.act, &p.pw_lo, &p.pw_hi);
  { const char* e = getenv("PCV_DP_DBG"); p.dbg = e ? atoi(e) : 0; }
  op->stride = g.S;
  op->smem_bytes = g.smem;
  op->grid = std::min(p.num_tiles, sm_count());
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  {
    cuuint64_t dims[4] = {(cuuint64_t)dw.Cin, (cuuint64_t)dw.W, (cuuint64_t)dw.H, (cuuint64_t)dw.N};
    cuuint64_t strides[3] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)dw.W * in_pitch * 2, (cuuint64_t)dw.H * dw.W * in_pitch * 2};
    cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)g.IW, (cuuint32_t)g.IH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&op->tmIn, TMAP_E16, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PCV_ERR_CUDA, "cuTensorMapEncodeTiled (dw -> pw halo) failed (%d)", (int)r);
  }
  const uint64_t kpad = (uint64_t)g.ncb * BLOCK_K;
  if (int rc = make_tiled_2d(&op->tmB, w_pw, kpad, pw.Cout, kpad * 2, BLOCK_K, g.nmma, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  char nm[176];
  snprintf(nm, sizeof nm, "conv_dwpw fused dw3x3 s%d C=%d -> 1x1 %d->%d @%dx%d%s st=%d", g.S, dw.Cin, dw.Cin, pw.Cout, dw.H, dw.W,
           res ? " +res" : "", g.stages);
  if (g.b_res) strncat(nm, " Bres", sizeof nm - strlen(nm) - 1);
  if (g.na_shift == 1) strncat(nm, " na2", sizeof nm - strlen(nm) - 1);
  op->name = nm;
  const double Mo = static_cast<double>(dw.N) * g.Ho * g.Wo;
  op->flops = 2.0 * Mo * dw.Cin * 9 + 2.0 * Mo * pw.Cout * dw.Cin;
  // fused group: x in, y out (+ identity), both weights; the depthwise tensor never touches HBM
  op->bytes = 2.0 * dw.N * dw.Cin * dw.H * dw.W + 2.0 * Mo * pw.Cout * (res ? 2.0 : 1.0) + 4.0 * dw.Cin * 10 +
              2.0 * dw.Cin * pw.Cout + 4.0 * pw.Cout;
  *out = op.release();
  return PCV_OK;
}

}  // namespace PCV_TIER
}  // namespace pcv
