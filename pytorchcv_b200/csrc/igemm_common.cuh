// Definitions shared by the 1-CTA (conv_igemm.cu) and CTA-pair (conv_igemm2.cu) implicit-GEMM kernels.
#pragma once
#include "ptx.cuh"
#include "runtime.h"

namespace pcv {
namespace PCV_TIER {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int NUM_THREADS = 256;
constexpr int EPI_THREADS = 128;

struct IgemmParams {
  const float* bias;          // [round_up(Cout,128)] folded BN bias
  void* out;                  // direct-store modes only
  const e16* res;   // direct-store modes only
  int M, Cout;
  int out_pitch, res_pitch;
  int HoWo, Wo;
  int stride, pad, dil, kw;
  int cblocks;                // 64-channel blocks per filter tap
  int num_kblocks;            // taps * cblocks
  int tiles_m, tiles_n;
  int act, has_res;
  float act_lo, act_hi;        // ReLU / ReLU6 / none as a clamp; other activations take the slow path
  float act_a;                 // PCV_ACT_LEAKY_RELU: negative slope
  const float* gate;           // PCV_CONV_SE_GATE (pair kernel): fp32 [images][Cout], multiplies (acc + bias) before the residual
  int n_img;                   // images (rows of `gate`)
  const float* bias2;          // gated dual-source conv: the shortcut conv's own folded bias (added outside the gate)
  int kb_split;                // dual-source 1x1 (pcv_conv1x1_dual): k-blocks >= kb_split read the SECOND activation (the tensor
  int a_mode2, stride2;        // map in the residual slot; its 1x1 conv may be strided), weights are K-concatenated
  int a_mode;                 // 0: 2-D tiled [Cin, M]; 1: im2col 4-D
  int out_mode;               // 0: TMA bf16 store; 1: direct bf16; 2: direct fp32
  int grouped;                // 1: A channel window = n_tile*g_in_span (block-diagonal weights)
  int g_in_span;              // grouped: input channels read by one 64-wide output tile = 64 * (Cin/g) / (Cout/g) <= 64
  int dbg;                    // PCV_IGEMM_DBG: bit0 skip operand TMA, bit1 skip MMA issue (throughput experiments)
  int stages, ksub, nstg;     // CTA-pair kernel: ring depth / 64-K sub-blocks per stage / staging slots (per layer)
  int nsubs;                  // CTA-pair kernel: 64-column sub-tiles per tile (tile width = nsubs*64 <= BN), per layer
};

// conv_igemm2.cu: launch the cta_group::2 kernel (BN = 128 or 256) on `grid` CTAs (a multiple of 2)
cudaError_t launch_igemm2(int bn, int grid, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                          const CUtensorMap& tmRes, const IgemmParams& p, cudaStream_t s);

// conv_igemm3.cu: 3x3 stride-1 halo kernel; PCV_ERR_UNSUPPORTED (no message) when the layer is outside its domain
int igemm3_try_make(const pcv_conv_desc& d, const void* x, const void* w, const float* bias, const void* res, void* y,
                    Op** out);

// conv_igemm3s.cu: space-to-depth stem (PCV_CONV_IN_OVERLAP view of 16-channel s2d pixels) with a 32-byte-row halo tile
int stem_halo_try_make(const pcv_conv_desc& d, const void* x, const void* w, const float* bias, const void* res,
                       void* y, Op** out);

// conv_igemm.cu helpers shared with conv_f32x3.cu
int pick_bn(const pcv_conv_desc& d, int tiles_m);
int make_tiled_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                  uint32_t box_inner, uint32_t box_rows, CUtensorMapSwizzle swz);
int make_im2col_4d(CUtensorMap* tm, const void* base, const pcv_conv_desc& d, int in_pitch);

void igemm2_pick_smem(int bn, int num_kblocks, bool has_res, int taps, int* stages, int* ksub, int* nstg);

}  // namespace PCV_TIER
}  // namespace pcv
