// Depthwise k x k convolution (+ folded BN bias, activation, optional residual), NHWC, bandwidth-bound.
// Reference: dwconv_block / dwconv3x3_block / dwconv5x5_block (pytorchcv/models/common/conv.py:437-543), used by
// LinearBottleneck (models/mobilenetv2.py:52-56) and DwsConvBlock (conv.py:546-608).
//
// Arithmetic intensity is ~3.4 FLOP/B, so the design goal is HBM efficiency: every thread owns 8 consecutive
// channels (one 16-byte bf16 vector) of a strip of TW consecutive output pixels along W, so each input vector is
// loaded once per strip row and reused from registers across the strip; consecutive threads own consecutive channel
// vectors, which makes every global access a fully coalesced 16-byte-per-lane transaction.  Weights are fp32
// [tap][C] (BN scale folded in) and stay L1/L2 resident.
#include "ptx.cuh"
#include "runtime.h"

namespace pcv {

struct DwParams {
  int N, H, W, C, Ho, Wo;
  int k, stride, pad, dil;
  int in_pitch, out_pitch, res_pitch;
  int act;
  int strips;  // ceil(Wo / TW)
};

__device__ __forceinline__ float dw_act(float v, int act) {
  switch (act) {
    case PCV_ACT_RELU: return fmaxf(v, 0.f);
    case PCV_ACT_RELU6: return fminf(fmaxf(v, 0.f), 6.f);
    case PCV_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    case PCV_ACT_SWISH: return v / (1.f + __expf(-v));
    case PCV_ACT_HSWISH: return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
    case PCV_ACT_HSIGMOID: return fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
    default: return v;
  }
}

template <typename T>
struct Vec8;
template <>
struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
    f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 v;
    v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
    v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = v;
  }
};
template <>
struct Vec8<__half> {
  static __device__ __forceinline__ void load(const __half* p, float (&f)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    f[0] = f16lo(v.x); f[1] = f16hi(v.x); f[2] = f16lo(v.y); f[3] = f16hi(v.y);
    f[4] = f16lo(v.z); f[5] = f16hi(v.z); f[6] = f16lo(v.w); f[7] = f16hi(v.w);
  }
  static __device__ __forceinline__ void store(__half* p, const float (&f)[8]) {
    uint4 v;
    v.x = pack_f16x2(f[0], f[1]); v.y = pack_f16x2(f[2], f[3]);
    v.z = pack_f16x2(f[4], f[5]); v.w = pack_f16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = v;
  }
};
template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// Specialised strip kernel: dilation 1, compile-time kernel size and stride.
template <typename T, int KS, int S, int TW>
__global__ void __launch_bounds__(256)
dwconv_strip_kernel(const DwParams p, const T* __restrict__ x, const float* __restrict__ w,
                    const float* __restrict__ bias, const T* __restrict__ res, T* __restrict__ y) {
  constexpr int SPAN = (TW - 1) * S + KS;  // input columns touched by one strip
  const int cvecs = p.C >> 3;
  const long long total = static_cast<long long>(p.N) * p.Ho * p.strips * cvecs;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = idx;
    const int cv = static_cast<int>(r % cvecs); r /= cvecs;
    const int strip = static_cast<int>(r % p.strips); r /= p.strips;
    const int ho = static_cast<int>(r % p.Ho);
    const int n = static_cast<int>(r / p.Ho);
    const int c = cv << 3;
    const int wo0 = strip * TW;
    const int wi0 = wo0 * S - p.pad;
    const int hi0 = ho * S - p.pad;

    float acc[TW][8];
    {
      float b[8];
      Vec8<float>::load(bias + c, b);
#pragma unroll
      for (int t = 0; t < TW; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[t][e] = b[e];
    }
#pragma unroll
    for (int fr = 0; fr < KS; ++fr) {
      const int hi = hi0 + fr;
      if (hi < 0 || hi >= p.H) continue;
      float wr[KS][8];
#pragma unroll
      for (int fs = 0; fs < KS; ++fs) Vec8<float>::load(w + static_cast<size_t>(fr * KS + fs) * p.C + c, wr[fs]);
      const T* row = x + (static_cast<size_t>(n) * p.H + hi) * p.W * p.in_pitch + c;
#pragma unroll
      for (int col = 0; col < SPAN; ++col) {
        const int wi = wi0 + col;
        if (wi < 0 || wi >= p.W) continue;
        float v[8];
        Vec8<T>::load(row + static_cast<size_t>(wi) * p.in_pitch, v);
#pragma unroll
        for (int t = 0; t < TW; ++t) {
          const int fs = col - t * S;  // compile-time after unrolling
          if (fs >= 0 && fs < KS) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[t][e] = fmaf(v[e], wr[fs][e], acc[t][e]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < TW; ++t) {
      const int wo = wo0 + t;
      if (wo >= p.Wo) break;
      const size_t pix = (static_cast<size_t>(n) * p.Ho + ho) * p.Wo + wo;
      if (res) {
        float rv[8];
        Vec8<T>::load(res + pix * p.res_pitch + c, rv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[t][e] += rv[e];
      }
      if (sizeof(T) == 2) {
        fast_act_n(acc[t], p.act);          // bf16 tier: MUFU.TANH based swish / sigmoid (ptx.cuh)
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[t][e] = dw_act(acc[t][e], p.act);
      }
      Vec8<T>::store(y + pix * p.out_pitch + c, acc[t]);
    }
  }
}

// Generic kernel: any kernel size / stride / dilation, one output pixel x 8 channels per thread.
template <typename T>
__global__ void __launch_bounds__(256)
dwconv_generic_kernel(const DwParams p, const T* __restrict__ x, const float* __restrict__ w,
                      const float* __restrict__ bias, const T* __restrict__ res, T* __restrict__ y) {
  const int cvecs = p.C >> 3;
  const long long total = static_cast<long long>(p.N) * p.Ho * p.Wo * cvecs;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = idx;
    const int cv = static_cast<int>(r % cvecs); r /= cvecs;
    const int wo = static_cast<int>(r % p.Wo); r /= p.Wo;
    const int ho = static_cast<int>(r % p.Ho);
    const int n = static_cast<int>(r / p.Ho);
    const int c = cv << 3;
    float acc[8];
    Vec8<float>::load(bias + c, acc);
    for (int fr = 0; fr < p.k; ++fr) {
      const int hi = ho * p.stride - p.pad + fr * p.dil;
      if (hi < 0 || hi >= p.H) continue;
      for (int fs = 0; fs < p.k; ++fs) {
        const int wi = wo * p.stride - p.pad + fs * p.dil;
        if (wi < 0 || wi >= p.W) continue;
        float v[8], wv[8];
        Vec8<T>::load(x + ((static_cast<size_t>(n) * p.H + hi) * p.W + wi) * p.in_pitch + c, v);
        Vec8<float>::load(w + static_cast<size_t>(fr * p.k + fs) * p.C + c, wv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(v[e], wv[e], acc[e]);
      }
    }
    const size_t pix = (static_cast<size_t>(n) * p.Ho + ho) * p.Wo + wo;
    if (res) {
      float rv[8];
      Vec8<T>::load(res + pix * p.res_pitch + c, rv);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += rv[e];
    }
    if (sizeof(T) == 2) {
      fast_act_n(acc, p.act);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = dw_act(acc[e], p.act);
    }
    Vec8<T>::store(y + pix * p.out_pitch + c, acc);
  }
}

// w [C, 1, k, k] -> fp32 [tap][C] with BN scale folded.
__global__ void dw_pack_kernel(const float* __restrict__ w, const float* __restrict__ conv_bias,
                               const float* __restrict__ g, const float* __restrict__ b,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps, int C,
                               int taps, int round16, float* __restrict__ wp, float* __restrict__ bias_out) {
  const int total = C * taps;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = idx % C, tap = idx / C;
    const float scale = g ? g[c] / sqrtf(var[c] + eps) : 1.f;
    float v = w[c * taps + tap] * scale;
    if (round16 == 1) v = __bfloat162float(__float2bfloat16(v));   // weights carry the tier's storage precision
    else if (round16 == 2) v = __half2float(__float2half_rn(v));
    wp[idx] = v;
  }
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    const float cb = conv_bias ? conv_bias[c] : 0.f;
    bias_out[c] = g ? (cb - mean[c]) * (g[c] / sqrtf(var[c] + eps)) + b[c] : cb;
  }
}

int dw_packed_bytes(const pcv_conv_desc& d, int dtype, size_t* w_bytes, size_t* b_bytes) {
  *w_bytes = static_cast<size_t>(d.Cout) * d.kh * d.kw * 4;
  *b_bytes = static_cast<size_t>(d.Cout) * 4;
  return PCV_OK;
}

int dw_pack(const pcv_conv_desc& d, int dtype, const float* w, const float* conv_bias, const float* g, const float* b,
            const float* m, const float* v, float eps, void* w_packed, float* bias_out, cudaStream_t s) {
  const int total = d.Cout * d.kh * d.kw;
  dw_pack_kernel<<<ceil_div(total, 256), 256, 0, s>>>(w, conv_bias, g, b, m, v, eps, d.Cout, d.kh * d.kw,
                                                      dtype == PCV_BF16 ? 1 : (dtype == PCV_F16 ? 2 : 0), reinterpret_cast<float*>(w_packed), bias_out);
  g_launches++;
  PCV_CHECK_CUDA(cudaGetLastError());
  return PCV_OK;
}

struct DwOp : Op {
  DwParams p;
  int dtype;
  const void* x;
  const float* w;
  const float* bias;
  const void* res;
  void* y;

  template <typename T>
  cudaError_t run(cudaStream_t s) {
    constexpr int TW = 4;
    const T* xx = reinterpret_cast<const T*>(x);
    const T* rr = reinterpret_cast<const T*>(res);
    T* yy = reinterpret_cast<T*>(y);
    const bool strip_ok = p.dil == 1 && (p.k == 3 || p.k == 5) && (p.stride == 1 || p.stride == 2);
    DwParams q = p;
    q.strips = ceil_div(p.Wo, TW);
    const long long items = strip_ok ? static_cast<long long>(p.N) * p.Ho * q.strips * (p.C >> 3)
                                     : static_cast<long long>(p.N) * p.Ho * p.Wo * (p.C >> 3);
    const int grid = static_cast<int>(std::min<long long>((items + 255) / 256, static_cast<long long>(sm_count()) * 32));
    if (!strip_ok) {
      dwconv_generic_kernel<T><<<grid, 256, 0, s>>>(q, xx, w, bias, rr, yy);
    } else if (p.k == 3 && p.stride == 1) {
      dwconv_strip_kernel<T, 3, 1, TW><<<grid, 256, 0, s>>>(q, xx, w, bias, rr, yy);
    } else if (p.k == 3) {
      dwconv_strip_kernel<T, 3, 2, TW><<<grid, 256, 0, s>>>(q, xx, w, bias, rr, yy);
    } else if (p.stride == 1) {
      dwconv_strip_kernel<T, 5, 1, TW><<<grid, 256, 0, s>>>(q, xx, w, bias, rr, yy);
    } else {
      dwconv_strip_kernel<T, 5, 2, TW><<<grid, 256, 0, s>>>(q, xx, w, bias, rr, yy);
    }
    return cudaGetLastError();
  }
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    return dtype == PCV_F32 ? run<float>(s) : (dtype == PCV_F16 ? run<__half>(s) : run<__nv_bfloat16>(s));
  }
};

int dw_make(const pcv_conv_desc& d, int dtype, const void* x, const void* w, const float* bias, const void* res,
            void* y, Op** out) {
  PCV_REQUIRE(d.kh == d.kw, "depthwise conv needs a square kernel");
  char nm[128];
  const double e = esize(dtype);
  const int Ho_ = conv_out(d.H, d.kh, d.stride, d.pad, d.dil), Wo_ = conv_out(d.W, d.kw, d.stride, d.pad, d.dil);
  const double M = static_cast<double>(d.N) * Ho_ * Wo_;
  const double flops = 2.0 * M * d.Cout * d.kh * d.kw;
  const double bytes = e * d.N * d.Cin * d.H * d.W + e * M * d.Cout * (res ? 2.0 : 1.0) + 4.0 * d.Cout * d.kh * d.kw + 4.0 * d.Cout;
  if (is16(dtype) && d.dil == 1) {
    // TMA halo-staged kernel (window_tma.cu) for the 3x3 stride-1/2 layers of the MobileNet family
    Op* wop = nullptr;
    const int rc = (dtype == PCV_F16 ? hf::win_make : bf::win_make)(0, d.N, d.H, d.W, d.Cout, d.kh, d.stride, d.pad, d.act, x, pitch_or(d.in_pitch, d.Cin),
                            reinterpret_cast<const float*>(w), bias, res, pitch_or(d.res_pitch, d.Cout), y,
                            pitch_or(d.out_pitch, d.Cout), &wop);
    if (rc == PCV_OK) {
      snprintf(nm, sizeof nm, "dwconv_tma_%s %dx%d s%d d%d C=%d @%dx%d%s", dtype_name(dtype), d.kh, d.kw, d.stride, d.dil,
               d.Cout, d.H, d.W, res ? " +res" : "");
      wop->name = nm; wop->flops = flops; wop->bytes = bytes;
      *out = wop;
      return PCV_OK;
    }
    if (rc != PCV_ERR_UNSUPPORTED) return rc;
  }
  auto op = std::make_unique<DwOp>();
  DwParams& p = op->p;
  p.N = d.N; p.H = d.H; p.W = d.W; p.C = d.Cout;
  p.Ho = conv_out(d.H, d.kh, d.stride, d.pad, d.dil);
  p.Wo = conv_out(d.W, d.kw, d.stride, d.pad, d.dil);
  PCV_REQUIRE(p.Ho > 0 && p.Wo > 0, "conv output is empty");
  p.k = d.kh; p.stride = d.stride; p.pad = d.pad; p.dil = d.dil;
  p.in_pitch = pitch_or(d.in_pitch, d.Cin);
  p.out_pitch = pitch_or(d.out_pitch, d.Cout);
  p.res_pitch = pitch_or(d.res_pitch, d.Cout);
  PCV_REQUIRE(p.C % 8 == 0 && p.in_pitch % 8 == 0 && p.out_pitch % 8 == 0 && (!res || p.res_pitch % 8 == 0),
              "depthwise conv needs channels and pitches that are multiples of 8");
  p.act = d.act;
  p.strips = 0;
  op->dtype = dtype; op->x = x; op->w = reinterpret_cast<const float*>(w); op->bias = bias; op->res = res; op->y = y;
  snprintf(nm, sizeof nm, "dwconv_%s %dx%d s%d d%d C=%d @%dx%d%s", dtype_name(dtype), d.kh, d.kw,
           d.stride, d.dil, d.Cout, d.H, d.W, res ? " +res" : "");
  op->name = nm;
  op->flops = flops;
  op->bytes = bytes;
  *out = op.release();
  return PCV_OK;
}

}  // namespace pcv
