// CUDA-core direct convolution: the true-fp32 tier (max rel err <= 1e-4 rules out TF32/bf16 tensor cores)
// and the generic fallback / on-GPU cross-check for shapes the tcgen05 path does not take (Cin % 8 != 0, odd groups).
// Same fused semantics as conv_igemm.cu: y = act(conv(x, w') + b' [+ residual]) with BN folded into w', b'
// (reference: pytorchcv/models/common/conv.py:278-286, models/resnet.py:221-229).
//
// Implicit GEMM on FFMA: CTA tile = 64 output pixels x 64 output channels of one group, K chunk = 16 input channels
// of one filter tap, 256 threads x (4 pixels x 4 channels) register tile.  Activations NHWC (any channel pitch),
// weights packed [group][tap][Cin/g][Cout/g] fp32 so both smem fills are contiguous.
#include "ptx.cuh"
#include "runtime.h"

namespace pcv {

struct SimtParams {
  int N, H, W, Ho, Wo, M;
  int cin_g, cout_g, groups;
  int kh, kw, stride, pad, dil;
  int in_pitch, out_pitch, res_pitch;
  int act, out_f32;
  float act_a;   // PCV_ACT_LEAKY_RELU: negative slope
};

__device__ __forceinline__ float simt_act(float v, int act, float a) {
  switch (act) {
    case PCV_ACT_LEAKY_RELU: return v >= 0.f ? v : v * a;
    case PCV_ACT_RELU: return fmaxf(v, 0.f);
    case PCV_ACT_RELU6: return fminf(fmaxf(v, 0.f), 6.f);
    case PCV_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case PCV_ACT_SWISH: return v / (1.f + expf(-v));
    case PCV_ACT_HSWISH: return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f;
    case PCV_ACT_HSIGMOID: return fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f;
    default: return v;
  }
}

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }
template <>
__device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// TILE = 64: the throughput shape (4 x 4 register tile per thread).  TILE = 32 (2 x 2 per thread): four times the CTAs
// for layers whose 64 x 64 grid cannot fill the machine (ResNet-18 at batch 8: 56 CTAs walked K = 4608 on the 7 x 7 maps).
template <typename T, int TILE>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const SimtParams p, const T* __restrict__ x, const float* __restrict__ w,
                 const float* __restrict__ bias, const T* __restrict__ res, void* __restrict__ y) {
  constexpr int TM = TILE, TN = TILE, TK = 16;
  constexpr int R = TILE / 16;        // register tile edge, also elements per thread of each smem fill
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const int g = blockIdx.z;

  // fill roles
  const int a_pix = tid / (16 / R);         // 0..TM-1
  const int a_c = (tid % (16 / R)) * R;     // first of this thread's R channels of the K chunk
  const int b_k = tid >> 4;                 // 0..15
  const int b_n = (tid & 15) * R;           // 0..TN-R
  // compute roles
  const int ty = tid >> 4;                  // pixel group
  const int tx = tid & 15;                  // channel group

  const int am = m0 + a_pix;
  const bool a_valid = am < p.M;
  int a_img = 0, a_h0 = 0, a_w0 = 0;
  if (a_valid) {
    a_img = am / (p.Ho * p.Wo);
    const int r = am - a_img * p.Ho * p.Wo;
    const int ho = r / p.Wo;
    a_h0 = ho * p.stride - p.pad;
    a_w0 = (r - ho * p.Wo) * p.stride - p.pad;
  }

  float acc[R][R];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.f;

  const int taps = p.kh * p.kw;
  for (int tap = 0; tap < taps; ++tap) {
    const int fr = tap / p.kw, fs = tap - fr * p.kw;
    const int hi = a_h0 + fr * p.dil, wi = a_w0 + fs * p.dil;
    const bool pix_ok = a_valid && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
    const T* xp = x + (static_cast<size_t>(a_img) * p.H * p.W + static_cast<size_t>(hi) * p.W + wi) * p.in_pitch +
                  static_cast<size_t>(g) * p.cin_g;
    const float* wp = w + (static_cast<size_t>(g) * taps + tap) * p.cin_g * p.cout_g;
    for (int c0 = 0; c0 < p.cin_g; c0 += TK) {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int c = c0 + a_c + i;
        As[a_c + i][a_pix] = (pix_ok && c < p.cin_g) ? to_f<T>(xp[c]) : 0.f;
      }
      {
        const int k = c0 + b_k;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int n = n0 + b_n + j;
          Bs[b_k][b_n + j] = (k < p.cin_g && n < p.cout_g) ? wp[static_cast<size_t>(k) * p.cout_g + n] : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        float av[R], bv[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
          av[i] = As[k][ty * R + i];
          bv[i] = Bs[k][tx * R + i];
        }
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
          for (int j = 0; j < R; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int m = m0 + ty * R + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int n = n0 + tx * R + j;
      if (n >= p.cout_g) continue;
      const int co = g * p.cout_g + n;
      float v = acc[i][j] + bias[co];
      if (res) v += to_f<T>(res[static_cast<size_t>(m) * p.res_pitch + co]);
      v = simt_act(v, p.act, p.act_a);
      const size_t o = static_cast<size_t>(m) * p.out_pitch + co;
      if (p.out_f32 || sizeof(T) == 4) {
        reinterpret_cast<float*>(y)[o] = v;
      } else {
        reinterpret_cast<T*>(y)[o] = from_f<T>(v);
      }
    }
  }
}

// w [Cout, Cin/g, kh, kw] -> [g][tap][Cin/g][Cout/g], BN folded; values rounded to bf16 when the tier is bf16 so
// this path reproduces the tensor-core path's operand precision.
__global__ void simt_pack_kernel(const float* __restrict__ w, const float* __restrict__ conv_bias,
                                 const float* __restrict__ g, const float* __restrict__ b,
                                 const float* __restrict__ mean, const float* __restrict__ var, float eps, int Cout,
                                 int cin_g, int cout_g, int taps, int round16, float* __restrict__ wp,
                                 float* __restrict__ bias_out) {
  const size_t total = static_cast<size_t>(Cout) * cin_g * taps;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // idx enumerates the destination
    size_t r = idx;
    const int n = static_cast<int>(r % cout_g); r /= cout_g;
    const int c = static_cast<int>(r % cin_g); r /= cin_g;
    const int tap = static_cast<int>(r % taps); r /= taps;
    const int grp = static_cast<int>(r);
    const int o = grp * cout_g + n;
    const float scale = g ? g[o] / sqrtf(var[o] + eps) : 1.f;
    float v = w[(static_cast<size_t>(o) * cin_g + c) * taps + tap] * scale;
    if (round16 == 1) v = __bfloat162float(__float2bfloat16(v));
    else if (round16 == 2) v = __half2float(__float2half_rn(v));
    wp[idx] = v;
  }
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < Cout; o += gridDim.x * blockDim.x) {
    const float cb = conv_bias ? conv_bias[o] : 0.f;
    bias_out[o] = g ? (cb - mean[o]) * (g[o] / sqrtf(var[o] + eps)) + b[o] : cb;
  }
}

int simt_packed_bytes(const pcv_conv_desc& d, int dtype, size_t* w_bytes, size_t* b_bytes) {
  *w_bytes = static_cast<size_t>(d.Cout) * (d.Cin / d.groups) * d.kh * d.kw * 4;
  *b_bytes = static_cast<size_t>(d.Cout) * 4;
  return PCV_OK;
}

int simt_pack(const pcv_conv_desc& d, int dtype, const float* w, const float* conv_bias, const float* g,
              const float* b, const float* m, const float* v, float eps, void* w_packed, float* bias_out,
              cudaStream_t s) {
  const size_t total = static_cast<size_t>(d.Cout) * (d.Cin / d.groups) * d.kh * d.kw;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 4096));
  simt_pack_kernel<<<blocks, 256, 0, s>>>(w, conv_bias, g, b, m, v, eps, d.Cout, d.Cin / d.groups, d.Cout / d.groups,
                                          d.kh * d.kw, dtype == PCV_BF16 ? 1 : (dtype == PCV_F16 ? 2 : 0), reinterpret_cast<float*>(w_packed), bias_out);
  g_launches++;
  PCV_CHECK_CUDA(cudaGetLastError());
  return PCV_OK;
}

struct SimtOp : Op {
  SimtParams p;
  int dtype;
  const void* x;
  const float* w;
  const float* bias;
  const void* res;
  void* y;
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    const long long ctas64 = static_cast<long long>(ceil_div(p.M, 64)) * ceil_div(p.cout_g, 64) * p.groups;
    const bool small = ctas64 < 2ll * sm_count() && ceil_div(p.M, 32) <= 2147483647 / 2;   // cannot fill the machine
    if (small) {
      dim3 grid(ceil_div(p.M, 32), ceil_div(p.cout_g, 32), p.groups);
      if (dtype == PCV_F32)
        conv_simt_kernel<float, 32><<<grid, 256, 0, s>>>(p, (const float*)x, w, bias, (const float*)res, y);
      else if (dtype == PCV_F16)
        conv_simt_kernel<__half, 32><<<grid, 256, 0, s>>>(p, (const __half*)x, w, bias,
                                                                 (const __half*)res, y);
    else
        conv_simt_kernel<__nv_bfloat16, 32><<<grid, 256, 0, s>>>(p, (const __nv_bfloat16*)x, w, bias,
                                                                 (const __nv_bfloat16*)res, y);
      return cudaGetLastError();
    }
    dim3 grid(ceil_div(p.M, 64), ceil_div(p.cout_g, 64), p.groups);
    if (dtype == PCV_F32)
      conv_simt_kernel<float, 64><<<grid, 256, 0, s>>>(p, (const float*)x, w, bias, (const float*)res, y);
    else if (dtype == PCV_F16)
      conv_simt_kernel<__half, 64><<<grid, 256, 0, s>>>(p, (const __half*)x, w, bias,
                                                               (const __half*)res, y);
    else
      conv_simt_kernel<__nv_bfloat16, 64><<<grid, 256, 0, s>>>(p, (const __nv_bfloat16*)x, w, bias,
                                                               (const __nv_bfloat16*)res, y);
    return cudaGetLastError();
  }
};

int simt_make(const pcv_conv_desc& d, int dtype, const void* x, const void* w, const float* bias, const void* res,
              void* y, Op** out) {
  auto op = std::make_unique<SimtOp>();
  SimtParams& p = op->p;
  p.N = d.N; p.H = d.H; p.W = d.W;
  p.Ho = conv_out(d.H, d.kh, d.stride, d.pad, d.dil);
  p.Wo = conv_out(d.W, d.kw, d.stride, d.pad, d.dil);
  PCV_REQUIRE(p.Ho > 0 && p.Wo > 0, "conv output is empty (H=%d W=%d k=%d)", d.H, d.W, d.kh);
  p.M = d.N * p.Ho * p.Wo;
  p.cin_g = d.Cin / d.groups; p.cout_g = d.Cout / d.groups; p.groups = d.groups;
  p.kh = d.kh; p.kw = d.kw; p.stride = d.stride; p.pad = d.pad; p.dil = d.dil;
  p.in_pitch = pitch_or(d.in_pitch, d.Cin);
  p.out_pitch = pitch_or(d.out_pitch, d.Cout);
  p.res_pitch = pitch_or(d.res_pitch, d.Cout);
  p.act = d.act;
  p.act_a = d.act_param;
  p.out_f32 = (d.flags & PCV_CONV_OUT_F32) ? 1 : 0;
  PCV_REQUIRE(ceil_div(p.cout_g, 32) <= 65535 && p.groups <= 65535, "grid too large for the CUDA-core conv");
  op->dtype = dtype; op->x = x; op->w = reinterpret_cast<const float*>(w); op->bias = bias; op->res = res; op->y = y;
  char nm[160];
  snprintf(nm, sizeof nm, "conv_simt_%s %dx%d s%d d%d g%d %d->%d @%dx%d%s", dtype_name(dtype), d.kh,
           d.kw, d.stride, d.dil, d.groups, d.Cin, d.Cout, d.H, d.W, res ? " +res" : "");
  op->name = nm;
  const double e = esize(dtype);
  const int taps = d.kh * d.kw;
  const double pin = (taps == 1 && d.stride > 1) ? (double)p.Ho * p.Wo : (double)d.H * d.W;
  op->flops = 2.0 * p.M * d.Cout * p.cin_g * taps;
  op->bytes = e * d.N * d.Cin * pin + (p.out_f32 ? 4.0 : e) * p.M * d.Cout + (res ? e * p.M * d.Cout : 0.0) +
              4.0 * d.Cout * p.cin_g * taps + 4.0 * d.Cout;
  *out = op.release();
  return PCV_OK;
}

}  // namespace pcv
