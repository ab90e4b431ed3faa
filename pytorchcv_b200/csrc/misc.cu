// Bandwidth-bound helpers of the eval path: pooling, SE squeeze/excite/scale, residual add, layout edges, bilinear.
// All NHWC with 8-channel (16-byte bf16) vectors per lane; consecutive lanes own consecutive channel vectors.
#include "ptx.cuh"
#include "runtime.h"

namespace pcv {

__device__ __forceinline__ float misc_act(float v, int act) {
  switch (act) {
    case PCV_ACT_RELU: return fmaxf(v, 0.f);
    case PCV_ACT_RELU6: return fminf(fmaxf(v, 0.f), 6.f);
    case PCV_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case PCV_ACT_SWISH: return v / (1.f + expf(-v));
    case PCV_ACT_HSWISH: return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f;
    case PCV_ACT_HSIGMOID: return fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f;
    case PCV_ACT_CLAMP01: return fminf(fmaxf(v, 0.f), 1.f);
    default: return v;
  }
}

template <typename T>
struct V8;
template <>
struct V8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
    f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 v;
    v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
    v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = v;
  }
  static __device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
};
template <>
struct V8<__half> {
  static __device__ __forceinline__ void load(const __half* p, float (&f)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    f[0] = f16lo(v.x); f[1] = f16hi(v.x); f[2] = f16lo(v.y); f[3] = f16hi(v.y);
    f[4] = f16lo(v.z); f[5] = f16hi(v.z); f[6] = f16lo(v.w); f[7] = f16hi(v.w);
  }
  static __device__ __forceinline__ void store(__half* p, const float (&f)[8]) {
    uint4 v;
    v.x = pack_f16x2(f[0], f[1]); v.y = pack_f16x2(f[2], f[3]);
    v.z = pack_f16x2(f[4], f[5]); v.w = pack_f16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = v;
  }
  static __device__ __forceinline__ float ld1(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ void st1(__half* p, float v) { *p = __float2half_rn(v); }
};
template <>
struct V8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0];
    const float4 b = reinterpret_cast<const float4*>(p)[1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  static __device__ __forceinline__ float ld1(const float* p) { return *p; }
  static __device__ __forceinline__ void st1(float* p, float v) { *p = v; }
};

static inline int grid_for(long long items, int block = 256) {
  return static_cast<int>(std::min<long long>((items + block - 1) / block, static_cast<long long>(sm_count()) * 32));
}

// ---------------------------------------------------------------------------------------------------------------
// max pool (nn.MaxPool2d: -inf padding, floor mode)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_kernel(int N, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad, const T* __restrict__ x,
               int in_pitch, T* __restrict__ y, int out_pitch) {
  const int cvecs = C >> 3;
  const long long total = static_cast<long long>(N) * Ho * Wo * cvecs;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = idx;
    const int cv = static_cast<int>(r % cvecs); r /= cvecs;
    const int wo = static_cast<int>(r % Wo); r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
    for (int fr = 0; fr < k; ++fr) {
      const int hi = ho * stride - pad + fr;
      if (hi < 0 || hi >= H) continue;
      for (int fs = 0; fs < k; ++fs) {
        const int wi = wo * stride - pad + fs;
        if (wi < 0 || wi >= W) continue;
        float v[8];
        V8<T>::load(x + ((static_cast<size_t>(n) * H + hi) * W + wi) * in_pitch + (cv << 3), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
      }
    }
    V8<T>::store(y + ((static_cast<size_t>(n) * Ho + ho) * Wo + wo) * out_pitch + (cv << 3), m);
  }
}

struct MaxPoolOp : Op {
  int dtype, N, H, W, C, Ho, Wo, k, stride, pad, in_pitch, out_pitch;
  const void* x;
  void* y;
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    const int grid = grid_for(static_cast<long long>(N) * Ho * Wo * (C >> 3));
    if (dtype == PCV_F32)
      maxpool_kernel<float><<<grid, 256, 0, s>>>(N, H, W, C, Ho, Wo, k, stride, pad, (const float*)x, in_pitch,
                                                 (float*)y, out_pitch);
    else if (dtype == PCV_F16)
      maxpool_kernel<__half><<<grid, 256, 0, s>>>(N, H, W, C, Ho, Wo, k, stride, pad, (const __half*)x,
                                                         in_pitch, (__half*)y, out_pitch);
    else
      maxpool_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(N, H, W, C, Ho, Wo, k, stride, pad, (const __nv_bfloat16*)x,
                                                         in_pitch, (__nv_bfloat16*)y, out_pitch);
    return cudaGetLastError();
  }
};

// ---------------------------------------------------------------------------------------------------------------
// global average pool / SE squeeze: one CTA per (image, channel slab).  The slab width is chosen on the host so that
// the grid has >= ~4 CTAs per SM; inside the CTA `cv` lanes own consecutive 8-channel vectors (16-byte coalesced
// loads, slab*2 contiguous bytes per pixel) and 256/cv pixel lanes stride over the pixels with 4 loads in flight per
// thread; fp32 accumulation, fixed-order smem tree reduce (deterministic).
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gap_kernel(int HW, int C, int slab, const T* __restrict__ x, int in_pitch, void* __restrict__ out, int out_f32) {
  __shared__ float red[256 * 8];
  const int n = blockIdx.y;
  const int c0 = blockIdx.x * slab;
  const int cv = slab >> 3;           // channel-vector lanes (power of two, <= 64)
  const int PL = 256 / cv;            // pixel lanes
  const int cvi = threadIdx.x & (cv - 1);
  const int pl = threadIdx.x / cv;
  const int c = c0 + cvi * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (c < C) {
    const T* base = x + static_cast<size_t>(n) * HW * in_pitch + c;
    int p = pl;
    for (; p + 3 * PL < HW; p += 4 * PL) {
      float v0[8], v1[8], v2[8], v3[8];
      V8<T>::load(base + static_cast<size_t>(p) * in_pitch, v0);
      V8<T>::load(base + static_cast<size_t>(p + PL) * in_pitch, v1);
      V8<T>::load(base + static_cast<size_t>(p + 2 * PL) * in_pitch, v2);
      V8<T>::load(base + static_cast<size_t>(p + 3 * PL) * in_pitch, v3);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += (v0[e] + v1[e]) + (v2[e] + v3[e]);
    }
    for (; p < HW; p += PL) {
      float v[8];
      V8<T>::load(base + static_cast<size_t>(p) * in_pitch, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
  }
  // red[pl][cvi][e], padded by one float per 8 to dodge bank conflicts on the column sums below
#pragma unroll
  for (int e = 0; e < 8; ++e) red[(pl * cv + cvi) * 8 + e] = acc[e];
  __syncthreads();
  for (int ch = threadIdx.x; ch < slab; ch += 256) {
    float s = 0.f;
    for (int q = 0; q < PL; ++q) s += red[q * slab + ch];
    if (c0 + ch < C) {
      const float mean = s / static_cast<float>(HW);
      const size_t o = static_cast<size_t>(n) * C + c0 + ch;
      if (out_f32) reinterpret_cast<float*>(out)[o] = mean;
      else V8<T>::st1(reinterpret_cast<T*>(out) + o, mean);   // the tier's own storage type
    }
  }
}

struct GapOp : Op {
  int dtype, N, HW, C, in_pitch, out_f32;
  const void* x;
  void* out;
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    int slab = 512;
    while (slab > 64 && (slab > C || static_cast<long long>(N) * ceil_div(C, slab) < 4ll * sm_count())) slab >>= 1;
    dim3 grid(ceil_div(C, slab), N);
    if (dtype == PCV_F32) gap_kernel<float><<<grid, 256, 0, s>>>(HW, C, slab, (const float*)x, in_pitch, out, out_f32);
    else if (dtype == PCV_F16) gap_kernel<__half><<<grid, 256, 0, s>>>(HW, C, slab, (const __half*)x, in_pitch, out, out_f32);
    else gap_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(HW, C, slab, (const __nv_bfloat16*)x, in_pitch, out, out_f32);
    return cudaGetLastError();
  }
};

// ---------------------------------------------------------------------------------------------------------------
// nn.AdaptiveAvgPool2d(k) (PyramidPoolingBranch, pspnet.py:71-75): bin (by, bx) averages rows
// [floor(by*H/k), ceil((by+1)*H/k)) x the same in W (torch's adaptive pooling rule); fp32 accumulate.
// CTA = one bin of one image x a slab of 256 channels: 32 lanes of 8-channel vectors x 8 pixel lanes, 4 loads in flight.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
adaptive_avgpool_kernel(int H, int W, int C, int OH, int OW, const T* __restrict__ x, int in_pitch, T* __restrict__ y) {
  __shared__ float red[8][256];
  const int bin = blockIdx.x, n = blockIdx.z;
  const int by = bin / OW, bx = bin - by * OW;
  const int h0 = (by * H) / OH, h1 = ((by + 1) * H + OH - 1) / OH;
  const int w0 = (bx * W) / OW, w1 = ((bx + 1) * W + OW - 1) / OW;
  const int cvi = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.y * 256 + cvi * 8;
  const int bw = w1 - w0, npx = (h1 - h0) * bw;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (c < C) {
    const T* base = x + (static_cast<size_t>(n) * H * W) * in_pitch + c;
    for (int p = pl; p < npx; p += 8) {
      const int r = p / bw, q = p - r * bw;
      float v[8];
      V8<T>::load(base + (static_cast<size_t>(h0 + r) * W + w0 + q) * in_pitch, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[pl][cvi * 8 + e] = acc[e];
  __syncthreads();
  const int ch = threadIdx.x;
  if (blockIdx.y * 256 + ch < C) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += red[q][ch];
    const float mean = s / static_cast<float>(npx);
    const size_t o = ((static_cast<size_t>(n) * OH + by) * OW + bx) * C + blockIdx.y * 256 + ch;
    V8<T>::st1(y + o, mean);
  }
}

struct AdaptivePoolOp : Op {
  int dtype, N, H, W, C, OH, OW, in_pitch;
  const void* x;
  void* y;
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    dim3 grid(OH * OW, ceil_div(C, 256), N);
    if (dtype == PCV_F32)
      adaptive_avgpool_kernel<float><<<grid, 256, 0, s>>>(H, W, C, OH, OW, (const float*)x, in_pitch, (float*)y);
    else if (dtype == PCV_F16)
      adaptive_avgpool_kernel<__half><<<grid, 256, 0, s>>>(H, W, C, OH, OW, (const __half*)x, in_pitch, (__half*)y);
    else
      adaptive_avgpool_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(H, W, C, OH, OW, (const __nv_bfloat16*)x, in_pitch, (__nv_bfloat16*)y);
    return cudaGetLastError();
  }
};

// ---------------------------------------------------------------------------------------------------------------
// skinny fp32 FC: y[n, j] = act(b[j] + sum_c x[n, c] * W[j, c]).  CTA = 8 images x 8 outputs (one output per warp),
// lanes stride over c so W rows are read coalesced and reused across the 8 images.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fc_f32_kernel(int N, int C, int J, const float* __restrict__ x, const float* __restrict__ W,
              const float* __restrict__ b, int act, float* __restrict__ y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  const int n0 = blockIdx.y * 8;
  if (j >= J) return;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const float* wr = W + static_cast<size_t>(j) * C;
  for (int c = lane; c < C; c += 32) {
    const float wv = __ldg(wr + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + i;
      if (n < N) acc[i] = fmaf(wv, __ldg(x + static_cast<size_t>(n) * C + c), acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  }
  if (lane == 0) {
    const float bj = b ? b[j] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + i;
      if (n < N) y[static_cast<size_t>(n) * J + j] = misc_act(acc[i] + bj, act);
    }
  }
}

// SE excite as two well-parallelised skinny-FC kernels (the per-image-group single-CTA version was latency-bound:
// 160 us at C = 2048 with only 64 CTAs in flight).  Both kernels tile (IMGS images) x (a slice of the outputs) per CTA
// so that ~500 CTAs are resident at C = 2048, keep the IMGS input rows in shared memory and read each weight row
// exactly once per CTA with float4 loads.
//   FC1: warp-per-hidden-unit (2 units per warp in flight), lanes stride over C, shuffle reduction.
//   FC2: thread-per-output-channel, the hidden vectors broadcast from smem, coalesced gate stores.
template <int IMGS, int JPW>
__global__ void __launch_bounds__(256)
se_fc1_kernel(int N, int C, int Cmid, const float* __restrict__ pooled, const float* __restrict__ w1,
              const float* __restrict__ b1, int mid_act, float* __restrict__ mid) {
  extern __shared__ float se_smem[];   // [IMGS][C]
  const int n0 = blockIdx.x * IMGS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c4n = C >> 2;
  for (int i = threadIdx.x; i < IMGS * c4n; i += 256) {
    const int img = i / c4n, c4 = i - img * c4n;
    const int n = min(n0 + img, N - 1);
    reinterpret_cast<float4*>(se_smem)[i] = __ldg(reinterpret_cast<const float4*>(pooled + static_cast<size_t>(n) * C) + c4);
  }
  __syncthreads();
  const int j0 = (blockIdx.y * 8 + warp) * JPW;
  float acc[JPW][IMGS];
#pragma unroll
  for (int u = 0; u < JPW; ++u)
#pragma unroll
    for (int i = 0; i < IMGS; ++i) acc[u][i] = 0.f;
#pragma unroll 2
  for (int c4 = lane; c4 < c4n; c4 += 32) {
    float4 wv[JPW];
#pragma unroll
    for (int u = 0; u < JPW; ++u)
      wv[u] = __ldg(reinterpret_cast<const float4*>(w1 + static_cast<size_t>(min(j0 + u, Cmid - 1)) * C) + c4);
#pragma unroll
    for (int i = 0; i < IMGS; ++i) {
      const float4 xv = reinterpret_cast<const float4*>(se_smem + i * C)[c4];
#pragma unroll
      for (int u = 0; u < JPW; ++u)
        acc[u][i] = fmaf(wv[u].x, xv.x, fmaf(wv[u].y, xv.y, fmaf(wv[u].z, xv.z, fmaf(wv[u].w, xv.w, acc[u][i]))));
    }
  }
#pragma unroll
  for (int u = 0; u < JPW; ++u)
#pragma unroll
    for (int i = 0; i < IMGS; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[u][i] += __shfl_xor_sync(0xffffffffu, acc[u][i], o);
    }
  if (lane == 0) {
#pragma unroll
    for (int u = 0; u < JPW; ++u) {
      const int j = j0 + u;
      if (j < Cmid) {
        const float bj = b1 ? b1[j] : 0.f;
#pragma unroll
        for (int i = 0; i < IMGS; ++i)
          if (n0 + i < N) mid[static_cast<size_t>(n0 + i) * Cmid + j] = misc_act(acc[u][i] + bj, mid_act);
      }
    }
  }
}

template <int IMGS>
__global__ void __launch_bounds__(256)
se_fc2_kernel(int N, int C, int Cmid, const float* __restrict__ mid, const float* __restrict__ w2,
              const float* __restrict__ b2, int out_act, float* __restrict__ gate) {
  extern __shared__ float se_smem[];   // [IMGS][Cmid]
  const int n0 = blockIdx.x * IMGS;
  const int m4n = Cmid >> 2;
  for (int i = threadIdx.x; i < IMGS * m4n; i += 256) {
    const int img = i / m4n, m4 = i - img * m4n;
    const int n = min(n0 + img, N - 1);
    reinterpret_cast<float4*>(se_smem)[i] = __ldg(reinterpret_cast<const float4*>(mid + static_cast<size_t>(n) * Cmid) + m4);
  }
  __syncthreads();
  const int o = blockIdx.y * 256 + threadIdx.x;
  if (o >= C) return;
  float acc[IMGS];
  const float bo = b2 ? b2[o] : 0.f;
#pragma unroll
  for (int i = 0; i < IMGS; ++i) acc[i] = bo;
  const float4* wr = reinterpret_cast<const float4*>(w2 + static_cast<size_t>(o) * Cmid);
#pragma unroll 8
  for (int m4 = 0; m4 < m4n; ++m4) {
    const float4 wv = __ldg(wr + m4);
#pragma unroll
    for (int i = 0; i < IMGS; ++i) {
      const float4 mv = reinterpret_cast<const float4*>(se_smem + i * Cmid)[m4];
      acc[i] = fmaf(wv.x, mv.x, fmaf(wv.y, mv.y, fmaf(wv.z, mv.z, fmaf(wv.w, mv.w, acc[i]))));
    }
  }
#pragma unroll
  for (int i = 0; i < IMGS; ++i)
    if (n0 + i < N) gate[static_cast<size_t>(n0 + i) * C + o] = misc_act(acc[i], out_act);
}

struct SeExciteOp : Op {
  int N, C, Cmid, mid_act, out_act;
  int Cin = 0;   // width of `pooled` / rows of w1 (== C for a plain SEBlock; the conv3-folded form pools conv3's INPUT)
  const float *pooled, *w1, *b1, *w2, *b2;
  float* gate;
  float* mid;  // scratch lives at gate + N*C (caller sizes gate as N*(C+Cmid))
  cudaError_t launch(cudaStream_t s) override {
    constexpr int IMGS = 4, JPW = 2;
    g_launches += 2;
    const int Ci = Cin > 0 ? Cin : C;
    const size_t smem1 = static_cast<size_t>(IMGS) * Ci * sizeof(float), smem2 = static_cast<size_t>(IMGS) * Cmid * sizeof(float);
    const bool aligned = ((reinterpret_cast<uintptr_t>(pooled) | reinterpret_cast<uintptr_t>(w1) |
                           reinterpret_cast<uintptr_t>(w2) | reinterpret_cast<uintptr_t>(mid)) & 15) == 0;
    if (C % 4 == 0 && Ci % 4 == 0 && Cmid % 4 == 0 && smem1 <= 200 * 1024 && aligned && N <= 65535 * IMGS) {
      if (smem1 > 48 * 1024) {
        static std::atomic<uint64_t> attr_done{0};   // per device (see runtime.h)
        if (cudaError_t e = set_max_smem_once(se_fc1_kernel<IMGS, JPW>, 200 * 1024, attr_done)) return e;
      }
      se_fc1_kernel<IMGS, JPW><<<dim3(ceil_div(N, IMGS), ceil_div(Cmid, 8 * JPW)), 256, smem1, s>>>(N, Ci, Cmid, pooled, w1, b1, mid_act, mid);
      se_fc2_kernel<IMGS><<<dim3(ceil_div(N, IMGS), ceil_div(C, 256)), 256, smem2, s>>>(N, C, Cmid, mid, w2, b2, out_act, gate);
      return cudaGetLastError();
    }
    fc_f32_kernel<<<dim3(ceil_div(Cmid, 8), ceil_div(N, 8)), 256, 0, s>>>(N, Ci, Cmid, pooled, w1, b1, mid_act, mid);
    fc_f32_kernel<<<dim3(ceil_div(C, 8), ceil_div(N, 8)), 256, 0, s>>>(N, Cmid, C, mid, w2, b2, out_act, gate);
    return cudaGetLastError();
  }
};

// ---------------------------------------------------------------------------------------------------------------
// SE scale (+ identity, + activation);  plain residual add
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
se_scale_kernel(int HW, int C, long long total_vecs, const T* __restrict__ x, const float* __restrict__ gate,
                const T* __restrict__ idn, int act, T* __restrict__ y) {
  const int cvecs = C >> 3;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total_vecs;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvecs);
    const long long pix = idx / cvecs;
    const int n = static_cast<int>(pix / HW);
    float v[8], g[8];
    V8<T>::load(x + idx * 8, v);
    V8<float>::load(gate + static_cast<size_t>(n) * C + cv * 8, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= g[e];
    if (idn) {
      float r[8];
      V8<T>::load(idn + idx * 8, r);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += r[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = misc_act(v[e], act);
    V8<T>::store(y + idx * 8, v);
  }
}

struct SeScaleOp : Op {
  int dtype, N, HW, C, act;
  const void *x, *idn;
  const float* gate;
  void* y;
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    const long long vecs = static_cast<long long>(N) * HW * (C >> 3);
    const int grid = grid_for(vecs);
    if (dtype == PCV_F32)
      se_scale_kernel<float><<<grid, 256, 0, s>>>(HW, C, vecs, (const float*)x, gate, (const float*)idn, act, (float*)y);
    else if (dtype == PCV_F16)
      se_scale_kernel<__half><<<grid, 256, 0, s>>>(HW, C, vecs, (const __half*)x, gate,
                                                          (const __half*)idn, act, (__half*)y);
    else
      se_scale_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(HW, C, vecs, (const __nv_bfloat16*)x, gate,
                                                          (const __nv_bfloat16*)idn, act, (__nv_bfloat16*)y);
    return cudaGetLastError();
  }
};

template <typename T>
__global__ void __launch_bounds__(256)
add_act_kernel(long long vecs, const T* __restrict__ a, const T* __restrict__ b, int act, T* __restrict__ y) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < vecs;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    float u[8], v[8];
    V8<T>::load(a + idx * 8, u);
    V8<T>::load(b + idx * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) u[e] = misc_act(u[e] + v[e], act);
    V8<T>::store(y + idx * 8, u);
  }
}

struct AddActOp : Op {
  int dtype, act;
  long long vecs;
  const void *a, *b;
  void* y;
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    const int grid = grid_for(vecs);
    if (dtype == PCV_F32) add_act_kernel<float><<<grid, 256, 0, s>>>(vecs, (const float*)a, (const float*)b, act, (float*)y);
    else if (dtype == PCV_F16) add_act_kernel<__half><<<grid, 256, 0, s>>>(vecs, (const __half*)a, (const __half*)b, act,
                                                             (__half*)y);
    else add_act_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(vecs, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, act,
                                                             (__nv_bfloat16*)y);
    return cudaGetLastError();
  }
};

// ---------------------------------------------------------------------------------------------------------------
// layout edges: NCHW fp32 <-> NHWC T
// ---------------------------------------------------------------------------------------------------------------
// Image element types the network edge accepts (include/pcv_b200.h pcv_image_type): the reference's fp32 NCHW tensor, or
// a 16-bit / 8-bit copy of it (half / a quarter of the host->device bytes).  value = float(x) * scale[c] + bias[c].
template <typename TI>
struct Img;
template <>
struct Img<float> {
  static __device__ __forceinline__ float ld1(const float* p) { return *p; }
  static __device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
  static __device__ __forceinline__ float4 ld4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
};
template <>
struct Img<__nv_bfloat16> {
  static __device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ float2 ld2(const __nv_bfloat16* p) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
    return make_float2(bf16lo(v), bf16hi(v));
  }
  static __device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    const uint2 v = __ldcs(reinterpret_cast<const uint2*>(p));
    return make_float4(bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y));
  }
};
template <>
struct Img<__half> {
  static __device__ __forceinline__ float ld1(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ float2 ld2(const __half* p) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
    return make_float2(f16lo(v), f16hi(v));
  }
  static __device__ __forceinline__ float4 ld4(const __half* p) {
    const uint2 v = __ldcs(reinterpret_cast<const uint2*>(p));
    return make_float4(f16lo(v.x), f16hi(v.x), f16lo(v.y), f16hi(v.y));
  }
};
template <>
struct Img<uint8_t> {
  static __device__ __forceinline__ float ld1(const uint8_t* p) { return static_cast<float>(*p); }
  static __device__ __forceinline__ float2 ld2(const uint8_t* p) {
    const uint16_t v = *reinterpret_cast<const uint16_t*>(p);
    return make_float2(static_cast<float>(v & 0xFF), static_cast<float>(v >> 8));
  }
  static __device__ __forceinline__ float4 ld4(const uint8_t* p) {
    const uint32_t v = __ldcs(reinterpret_cast<const uint32_t*>(p));
    return make_float4(static_cast<float>(v & 0xFF), static_cast<float>((v >> 8) & 0xFF),
                       static_cast<float>((v >> 16) & 0xFF), static_cast<float>(v >> 24));
  }
};
struct ImgAffine {   // per-channel value = x * scale + bias for the first 4 channels (identity beyond)
  float scale[4], bias[4];
};
static inline size_t img_esize(int img_type) { return img_type == PCV_IMG_F32 ? 4 : (img_type == PCV_IMG_U8 ? 1 : 2); }
static inline const char* img_name(int img_type) {
  return img_type == PCV_IMG_F32 ? "f32" : (img_type == PCV_IMG_U8 ? "u8" : (img_type == PCV_IMG_F16 ? "f16" : "bf16"));
}
static inline ImgAffine make_affine(int C, const float* scale_host, const float* bias_host) {
  ImgAffine af;
  for (int c = 0; c < 4; ++c) {
    af.scale[c] = (scale_host && c < C) ? scale_host[c] : 1.f;
    af.bias[c] = (bias_host && c < C) ? bias_host[c] : 0.f;
  }
  return af;
}


template <typename TI, typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(int N, int C, int HW, int c_pitch, const TI* __restrict__ x, T* __restrict__ y, ImgAffine af) {
  const int cvecs = c_pitch >> 3;
  const long long total = static_cast<long long>(N) * cvecs * HW;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(idx % HW);
    const long long r = idx / HW;
    const int cv = static_cast<int>(r % cvecs);
    const int n = static_cast<int>(r / cvecs);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cv * 8 + e;
      float t = c < C ? Img<TI>::ld1(x + (static_cast<size_t>(n) * C + c) * HW + p) : 0.f;
      if (c < 4 && c < C) t = fmaf(t, af.scale[c], af.bias[c]);
      v[e] = t;
    }
    V8<T>::store(y + (static_cast<size_t>(n) * HW + p) * c_pitch + cv * 8, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(int N, int C, int HW, int c_pitch, const T* __restrict__ x, float* __restrict__ y) {
  const long long total = static_cast<long long>(N) * C * HW;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(idx % HW);
    const long long r = idx / HW;
    const int c = static_cast<int>(r % C);
    const int n = static_cast<int>(r / C);
    y[idx] = V8<T>::ld1(x + (static_cast<size_t>(n) * HW + p) * c_pitch + c);
  }
}

// nn.ZeroPad2d((left, right, top, bottom)) on NHWC (conv.py:245-249: ConvBlock with a 4-tuple padding; EfficientNet tf_mode's
// F.pad(calc_tf_padding(...)), efficientnet.py:27-55,108-109,189-190,236-237): one pass, 8 channels per thread
template <typename T>
__global__ void __launch_bounds__(256)
zero_pad_kernel(int N, int H, int W, int C, int in_pitch, int pl, int pt, int Ho, int Wo, int out_pitch,
                const T* __restrict__ x, T* __restrict__ y) {
  const int cv = C >> 3;
  const long long total = static_cast<long long>(N) * Ho * Wo * cv;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = idx;
    const int c8 = static_cast<int>(r % cv) << 3; r /= cv;
    const int wo = static_cast<int>(r % Wo); r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    const int h = ho - pt, w = wo - pl;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (h >= 0 && h < H && w >= 0 && w < W) V8<T>::load(x + ((static_cast<size_t>(n) * H + h) * W + w) * in_pitch + c8, v);
    V8<T>::store(y + ((static_cast<size_t>(n) * Ho + ho) * Wo + wo) * out_pitch + c8, v);
  }
}

struct ZeroPadOp : Op {
  int dtype, N, H, W, C, in_pitch, pl, pt, Ho, Wo, out_pitch;
  const void* x;
  void* y;
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    const int grid = grid_for(static_cast<long long>(N) * Ho * Wo * (C >> 3));
    if (dtype == PCV_F32)
      zero_pad_kernel<float><<<grid, 256, 0, s>>>(N, H, W, C, in_pitch, pl, pt, Ho, Wo, out_pitch, (const float*)x, (float*)y);
    else if (dtype == PCV_F16)
      zero_pad_kernel<__half><<<grid, 256, 0, s>>>(N, H, W, C, in_pitch, pl, pt, Ho, Wo, out_pitch, (const __half*)x, (__half*)y);
    else
      zero_pad_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(N, H, W, C, in_pitch, pl, pt, Ho, Wo, out_pitch, (const __nv_bfloat16*)x,
                                                          (__nv_bfloat16*)y);
    return cudaGetLastError();
  }
};

struct LayoutOp : Op {
  int dtype, N, C, HW, c_pitch, to_nhwc, img_type = PCV_IMG_F32;
  ImgAffine af;
  const void* x;
  void* y;
  template <typename TI>
  void ingest(int grid, cudaStream_t s) {
    if (dtype == PCV_F32) nchw_to_nhwc_kernel<TI, float><<<grid, 256, 0, s>>>(N, C, HW, c_pitch, (const TI*)x, (float*)y, af);
    else if (dtype == PCV_F16) nchw_to_nhwc_kernel<TI, __half><<<grid, 256, 0, s>>>(N, C, HW, c_pitch, (const TI*)x, (__half*)y, af);
    else nchw_to_nhwc_kernel<TI, __nv_bfloat16><<<grid, 256, 0, s>>>(N, C, HW, c_pitch, (const TI*)x, (__nv_bfloat16*)y, af);
  }
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    if (to_nhwc) {
      const int grid = grid_for(static_cast<long long>(N) * (c_pitch >> 3) * HW);
      switch (img_type) {
        case PCV_IMG_BF16: ingest<__nv_bfloat16>(grid, s); break;
        case PCV_IMG_F16: ingest<__half>(grid, s); break;
        case PCV_IMG_U8: ingest<uint8_t>(grid, s); break;
        default: ingest<float>(grid, s); break;
      }
    } else {
      const int grid = grid_for(static_cast<long long>(N) * C * HW);
      if (dtype == PCV_F32) nhwc_to_nchw_kernel<float><<<grid, 256, 0, s>>>(N, C, HW, c_pitch, (const float*)x, (float*)y);
      else if (dtype == PCV_F16) nhwc_to_nchw_kernel<__half><<<grid, 256, 0, s>>>(N, C, HW, c_pitch, (const __half*)x, (float*)y);
      else nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(N, C, HW, c_pitch, (const __nv_bfloat16*)x, (float*)y);
    }
    return cudaGetLastError();
  }
};

// ---------------------------------------------------------------------------------------------------------------
// bilinear, align_corners=True:  src = dst * (in-1)/(out-1)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
bilinear_nchw_kernel(int N, int Hin, int Win, int C, const T* __restrict__ x, int in_pitch, int Hout, int Wout,
                     float* __restrict__ y, float sh, float sw) {
  // thread = (n, c, oh, ow), ow-fastest: coalesced fp32 plane writes; the 4 taps of neighbouring lanes share lines
  const long long total = static_cast<long long>(N) * C * Hout * Wout;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ow = static_cast<int>(idx % Wout);
    long long r = idx / Wout;
    const int oh = static_cast<int>(r % Hout); r /= Hout;
    const int c = static_cast<int>(r % C);
    const int n = static_cast<int>(r / C);
    const float fy = oh * sh, fx = ow * sw;
    int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    y0 = min(y0, Hin - 1); x0 = min(x0, Win - 1);
    const int y1 = min(y0 + 1, Hin - 1), x1 = min(x0 + 1, Win - 1);
    const float ly = fy - y0, lx = fx - x0;
    const T* base = x + static_cast<size_t>(n) * Hin * Win * in_pitch + c;
    const float v00 = V8<T>::ld1(base + (static_cast<size_t>(y0) * Win + x0) * in_pitch);
    const float v01 = V8<T>::ld1(base + (static_cast<size_t>(y0) * Win + x1) * in_pitch);
    const float v10 = V8<T>::ld1(base + (static_cast<size_t>(y1) * Win + x0) * in_pitch);
    const float v11 = V8<T>::ld1(base + (static_cast<size_t>(y1) * Win + x1) * in_pitch);
    y[idx] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}

// Row-staged variant for the network-edge upsample (DeepLabv3FinalBlock, deeplabv3.py:53): one CTA per (image,
// output row).  The VERTICAL interpolation happens while staging: smem holds one fp32 row v[c][x] = lerp(row y0, row y1).
// A thread then owns four adjacent output columns (their source columns x0 / weights are computed once, and at
// upsampling ratios >= 2 they touch at most three adjacent source pixels a, a+1, a+2) and walks over the channels: three
// smem reads, four selects and four FMAs per float4 of the fp32 NCHW output, streaming stores.  (The first version
// recomputed both interpolations per output element: ~120 instructions per float4, issue-bound at 0.145 ms per
// 16 x 21 x 480 x 480 map; the map itself is 310 MB of HBM writes.)
template <typename T>
__global__ void __launch_bounds__(256)
bilinear_nchw_rows_kernel(int Hin, int Win, int C, const T* __restrict__ x, int in_pitch, int Hout, int Wout,
                          float* __restrict__ y, float sh, float sw) {
  extern __shared__ float bl_smem[];   // [C][Win + 2] (two pad columns so a+1, a+2 never leave the row)
  const int oh = blockIdx.x, n = blockIdx.y;
  const float fy = oh * sh;
  const int y0 = min(static_cast<int>(fy), Hin - 1);
  const int y1 = min(y0 + 1, Hin - 1);
  const float ly = fy - y0;
  const int WP = Win + 2;
  const T* row0 = x + (static_cast<size_t>(n) * Hin + y0) * Win * in_pitch;
  const T* row1 = x + (static_cast<size_t>(n) * Hin + y1) * Win * in_pitch;
  for (int xx = threadIdx.x / 32; xx < Win; xx += 8) {        // a warp per source pixel, lanes over its channels
    for (int c = threadIdx.x & 31; c < C; c += 32) {
      const float a = V8<T>::ld1(row0 + static_cast<size_t>(xx) * in_pitch + c);
      const float b = V8<T>::ld1(row1 + static_cast<size_t>(xx) * in_pitch + c);
      const float v = (1.f - ly) * a + ly * b;
      bl_smem[c * WP + xx] = v;
      if (xx == Win - 1) bl_smem[c * WP + Win] = bl_smem[c * WP + Win + 1] = v;   // clamp-to-edge pads
    }
  }
  __syncthreads();
  const int w4 = Wout >> 2;
  const int csplit = max(1, 256 / w4);                       // channel slices worked on concurrently
  for (int i = threadIdx.x; i < w4 * csplit; i += 256) {
    const int cs = i / w4, ow0 = (i - cs * w4) << 2;
    int a = min(static_cast<int>(ow0 * sw), Win - 1);
    float lx[4];
    bool hi[4], hi2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float fx = (ow0 + e) * sw;
      const int x0 = min(static_cast<int>(fx), Win - 1);
      lx[e] = fx - x0;
      hi[e] = x0 >= a + 1;      // x0 in {a, a+1} (sw <= 0.5); kept exact for x0 == a + 2 by hi2
      hi2[e] = x0 >= a + 2;
    }
    const bool wide = hi2[3];                                // only possible when sw > 1/3: take the generic path below
    float* dst = y + ((static_cast<size_t>(n) * C + cs) * Hout + oh) * Wout + ow0;
    const size_t cstep = static_cast<size_t>(csplit) * Hout * Wout;
    for (int c = cs; c < C; c += csplit, dst += cstep) {
      const float* r = bl_smem + c * WP;
      float o[4];
      if (!wide) {
        const float v0 = r[a], v1 = r[a + 1], v2 = r[a + 2];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = hi[e] ? v1 : v0, p1 = hi[e] ? v2 : v1;
          o[e] = fmaf(lx[e], p1 - p0, p0);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float fx = (ow0 + e) * sw;
          const int x0 = min(static_cast<int>(fx), Win - 1);
          o[e] = fmaf(lx[e], r[x0 + 1] - r[x0], r[x0]);
        }
      }
      __stcs(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bilinear_nhwc_kernel(int N, int Hin, int Win, int C, const T* __restrict__ x, int in_pitch, int Hout, int Wout,
                     T* __restrict__ y, int out_pitch, float sh, float sw) {
  const int cvecs = C >> 3;
  const long long total = static_cast<long long>(N) * Hout * Wout * cvecs;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvecs);
    long long r = idx / cvecs;
    const int ow = static_cast<int>(r % Wout); r /= Wout;
    const int oh = static_cast<int>(r % Hout);
    const int n = static_cast<int>(r / Hout);
    const float fy = oh * sh, fx = ow * sw;
    int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    y0 = min(y0, Hin - 1); x0 = min(x0, Win - 1);
    const int y1 = min(y0 + 1, Hin - 1), x1 = min(x0 + 1, Win - 1);
    const float ly = fy - y0, lx = fx - x0;
    const T* base = x + static_cast<size_t>(n) * Hin * Win * in_pitch + cv * 8;
    float a[8], b[8], c2[8], d[8], o[8];
    V8<T>::load(base + (static_cast<size_t>(y0) * Win + x0) * in_pitch, a);
    V8<T>::load(base + (static_cast<size_t>(y0) * Win + x1) * in_pitch, b);
    V8<T>::load(base + (static_cast<size_t>(y1) * Win + x0) * in_pitch, c2);
    V8<T>::load(base + (static_cast<size_t>(y1) * Win + x1) * in_pitch, d);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      o[e] = (1.f - ly) * ((1.f - lx) * a[e] + lx * b[e]) + ly * ((1.f - lx) * c2[e] + lx * d[e]);
    V8<T>::store(y + ((static_cast<size_t>(n) * Hout + oh) * Wout + ow) * out_pitch + cv * 8, o);
  }
}

struct BilinearOp : Op {
  int dtype, N, Hin, Win, C, in_pitch, Hout, Wout, out_pitch, nchw;
  const void* x;
  void* y;
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    const float sh = Hout > 1 ? static_cast<float>(Hin - 1) / static_cast<float>(Hout - 1) : 0.f;
    const float sw = Wout > 1 ? static_cast<float>(Win - 1) / static_cast<float>(Wout - 1) : 0.f;
    if (nchw && Wout % 4 == 0 && N <= 65535 && static_cast<size_t>(C) * (Win + 2) * 4 <= 48 * 1024 &&
        reinterpret_cast<uintptr_t>(y) % 16 == 0) {
      const size_t smem = static_cast<size_t>(C) * (Win + 2) * sizeof(float);
      if (dtype == PCV_F32)
        bilinear_nchw_rows_kernel<float><<<dim3(Hout, N), 256, smem, s>>>(Hin, Win, C, (const float*)x, in_pitch, Hout, Wout, (float*)y, sh, sw);
      else if (dtype == PCV_F16)
        bilinear_nchw_rows_kernel<__half><<<dim3(Hout, N), 256, smem, s>>>(Hin, Win, C, (const __half*)x, in_pitch, Hout, Wout, (float*)y, sh, sw);
    else
        bilinear_nchw_rows_kernel<__nv_bfloat16><<<dim3(Hout, N), 256, smem, s>>>(Hin, Win, C, (const __nv_bfloat16*)x, in_pitch, Hout, Wout, (float*)y, sh, sw);
    } else if (nchw) {
      const int grid = grid_for(static_cast<long long>(N) * C * Hout * Wout);
      if (dtype == PCV_F32)
        bilinear_nchw_kernel<float><<<grid, 256, 0, s>>>(N, Hin, Win, C, (const float*)x, in_pitch, Hout, Wout, (float*)y, sh, sw);
      else if (dtype == PCV_F16)
        bilinear_nchw_kernel<__half><<<grid, 256, 0, s>>>(N, Hin, Win, C, (const __half*)x, in_pitch, Hout,
                                                                 Wout, (float*)y, sh, sw);
    else
        bilinear_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(N, Hin, Win, C, (const __nv_bfloat16*)x, in_pitch, Hout,
                                                                 Wout, (float*)y, sh, sw);
    } else {
      const int grid = grid_for(static_cast<long long>(N) * Hout * Wout * (C >> 3));
      if (dtype == PCV_F32)
        bilinear_nhwc_kernel<float><<<grid, 256, 0, s>>>(N, Hin, Win, C, (const float*)x, in_pitch, Hout, Wout, (float*)y,
                                                         out_pitch, sh, sw);
      else if (dtype == PCV_F16)
        bilinear_nhwc_kernel<__half><<<grid, 256, 0, s>>>(N, Hin, Win, C, (const __half*)x, in_pitch, Hout,
                                                                 Wout, (__half*)y, out_pitch, sh, sw);
    else
        bilinear_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(N, Hin, Win, C, (const __nv_bfloat16*)x, in_pitch, Hout,
                                                                 Wout, (__nv_bfloat16*)y, out_pitch, sh, sw);
    }
    return cudaGetLastError();
  }
};

// y = act(x * scale[c] + shift[c]), negative side times slope[c]: the stand-alone BN -> ReLU pre-activation of PreConvBlock /
// PreResActivation (conv.py:717-731, preresnet.py:203-221), nn.PReLU and LeakyReLU (activ.py:84-120).  One pass at the copy
// roofline: a thread owns 8 channels of a pixel (one 16-byte load / store in the 16-bit tiers), the per-channel vectors
// come through the read-only path (C * 12 bytes, L1-resident).
template <typename T>
__global__ void __launch_bounds__(256)
channel_affine_act_kernel(long long pixels, int C, int in_pitch, int out_pitch, const T* __restrict__ x,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ slope, int act, T* __restrict__ y) {
  const int cv = C >> 3;
  const long long total = pixels * cv;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(idx % cv) << 3;
    const long long p = idx / cv;
    float v[8];
    V8<T>::load(x + p * in_pitch + c8, v);
    if (scale) {
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c8)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c8 + 4));
      v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w; v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
    }
    if (shift) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + c8)), b1 = __ldg(reinterpret_cast<const float4*>(shift + c8 + 4));
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (act != PCV_ACT_NONE) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = misc_act(v[e], act);
    }
    if (slope) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(slope + c8)), a1 = __ldg(reinterpret_cast<const float4*>(slope + c8 + 4));
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = v[e] >= 0.f ? v[e] : v[e] * a[e];
    }
    V8<T>::store(y + p * out_pitch + c8, v);
  }
}

struct ChannelAffineOp : Op {
  int dtype, C, in_pitch, out_pitch, act;
  long long pixels;
  const void* x;
  const float *scale, *shift, *slope;
  void* y;
  template <typename T>
  void run(int grid, cudaStream_t s) {
    channel_affine_act_kernel<T><<<grid, 256, 0, s>>>(pixels, C, in_pitch, out_pitch, (const T*)x, scale, shift, slope, act, (T*)y);
  }
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    const int grid = grid_for(pixels * (C >> 3));
    if (dtype == PCV_F32) run<float>(grid, s);
    else if (dtype == PCV_F16) run<__half>(grid, s);
    else run<__nv_bfloat16>(grid, s);
    return cudaGetLastError();
  }
};

}  // namespace pcv

using namespace pcv;

static const char* dn(int dtype) { return dtype_name(dtype); }
#define PCV_DTYPE_OK(dt) PCV_REQUIRE((dt) == PCV_BF16 || (dt) == PCV_F32 || (dt) == PCV_F16, "unknown dtype %d", (dt))

extern "C" {

int pcv_maxpool2d(pcv_plan* plan, int dtype, int N, int H, int W, int C, int k, int stride, int pad, const void* x,
                  int in_pitch, void* y, int out_pitch, pcv_stream stream) {
  PCV_DTYPE_OK(dtype);
  PCV_REQUIRE(x && y, "NULL tensor pointer");
  PCV_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0 && pad >= 0 && 2 * pad <= k, "bad maxpool dims");
  in_pitch = pitch_or(in_pitch, C);
  out_pitch = pitch_or(out_pitch, C);
  PCV_REQUIRE(C % 8 == 0 && in_pitch % 8 == 0 && out_pitch % 8 == 0, "maxpool needs channel counts/pitches % 8 == 0");
  if (is16(dtype)) {
    Op* wop = nullptr;
    const int rc = (dtype == PCV_F16 ? hf::win_make : bf::win_make)(1, N, H, W, C, k, stride, pad, PCV_ACT_NONE, x, in_pitch, nullptr, nullptr, nullptr, 0, y,
                            out_pitch, &wop);
    if (rc == PCV_OK) {
      char nm[96];
      snprintf(nm, sizeof nm, "maxpool_tma_%s %dx%d s%d C=%d @%dx%d", dn(dtype), k, k, stride, C, H, W);
      wop->name = nm;
      const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
      wop->bytes = 2.0 * static_cast<double>(N) * C * (static_cast<double>(H) * W + static_cast<double>(Ho) * Wo);
      return submit(plan, wop, static_cast<cudaStream_t>(stream));
    }
    if (rc != PCV_ERR_UNSUPPORTED) return rc;
  }
  auto op = std::make_unique<MaxPoolOp>();
  op->dtype = dtype; op->N = N; op->H = H; op->W = W; op->C = C; op->k = k; op->stride = stride; op->pad = pad;
  op->Ho = (H + 2 * pad - k) / stride + 1;
  op->Wo = (W + 2 * pad - k) / stride + 1;
  PCV_REQUIRE(op->Ho > 0 && op->Wo > 0, "maxpool output is empty");
  op->in_pitch = in_pitch; op->out_pitch = out_pitch; op->x = x; op->y = y;
  char nm[96];
  snprintf(nm, sizeof nm, "maxpool_%s %dx%d s%d C=%d @%dx%d", dn(dtype), k, k, stride, C, H, W);
  op->name = nm;
  op->bytes = esize(dtype) * static_cast<double>(N) * C * (static_cast<double>(H) * W + static_cast<double>(op->Ho) * op->Wo);
  return submit(plan, op.release(), static_cast<cudaStream_t>(stream));
}

int pcv_global_avgpool(pcv_plan* plan, int dtype, int N, int HW, int C, const void* x, int in_pitch, void* pooled,
                       int out_dtype, pcv_stream stream) {
  PCV_DTYPE_OK(dtype);
  PCV_DTYPE_OK(out_dtype);
  PCV_REQUIRE(x && pooled, "NULL tensor pointer");
  PCV_REQUIRE(N > 0 && HW > 0 && C > 0 && N <= 65535, "bad avgpool dims");
  in_pitch = pitch_or(in_pitch, C);
  PCV_REQUIRE(C % 8 == 0 && in_pitch % 8 == 0, "avgpool needs channel count/pitch % 8 == 0");
  auto op = std::make_unique<GapOp>();
  op->dtype = dtype; op->N = N; op->HW = HW; op->C = C; op->in_pitch = in_pitch; op->out_f32 = out_dtype == PCV_F32;
  op->x = x; op->out = pooled;
  char nm[96];
  snprintf(nm, sizeof nm, "gavgpool_%s C=%d HW=%d", dn(dtype), C, HW);
  op->name = nm;
  op->bytes = esize(dtype) * static_cast<double>(N) * C * HW + esize(out_dtype) * static_cast<double>(N) * C;
  return submit(plan, op.release(), static_cast<cudaStream_t>(stream));
}

int pcv_adaptive_avgpool(pcv_plan* plan, int dtype, int N, int H, int W, int C, const void* x, int in_pitch, int out_h,
                         int out_w, void* y, pcv_stream stream) {
  PCV_DTYPE_OK(dtype);
  PCV_REQUIRE(x && y, "NULL tensor pointer");
  PCV_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && N <= 65535, "bad adaptive avgpool dims");
  PCV_REQUIRE(out_h > 0 && out_w > 0 && out_h <= H && out_w <= W && out_h * out_w <= 65535, "adaptive avgpool output %dx%d must fit the %dx%d map", out_h, out_w, H, W);
  in_pitch = pitch_or(in_pitch, C);
  PCV_REQUIRE(C % 8 == 0 && in_pitch % 8 == 0, "adaptive avgpool needs channel count/pitch % 8 == 0");
  auto op = std::make_unique<AdaptivePoolOp>();
  op->dtype = dtype; op->N = N; op->H = H; op->W = W; op->C = C; op->OH = out_h; op->OW = out_w; op->in_pitch = in_pitch;
  op->x = x; op->y = y;
  char nm[96];
  snprintf(nm, sizeof nm, "adaptive_avgpool_%s C=%d %dx%d->%dx%d", dn(dtype), C, H, W, out_h, out_w);
  op->name = nm;
  // overlapping bins re-read the shared rows / columns (L2 hits); algorithmic bytes = the map once + the bins
  op->bytes = esize(dtype) * static_cast<double>(N) * C * (static_cast<double>(H) * W + out_h * out_w);
  return submit(plan, op.release(), static_cast<cudaStream_t>(stream));
}

int pcv_se_excite_ex(pcv_plan* plan, int N, int Cin, int Cmid, int C, const float* pooled, const float* w1, const float* b1,
                     const float* w2, const float* b2, int mid_act, int out_act, float* gate, pcv_stream stream) {
  PCV_REQUIRE(pooled && w1 && w2 && gate, "NULL tensor pointer");
  PCV_REQUIRE(N > 0 && C > 0 && Cmid > 0 && Cin > 0, "bad SE dims");
  auto op = std::make_unique<SeExciteOp>();
  op->N = N; op->C = C; op->Cin = Cin; op->Cmid = Cmid; op->mid_act = mid_act; op->out_act = out_act;
  op->pooled = pooled; op->w1 = w1; op->b1 = b1; op->w2 = w2; op->b2 = b2; op->gate = gate;
  op->mid = gate + static_cast<size_t>(N) * C;
  op->launches = 2;
  char nm[96];
  if (Cin == C) snprintf(nm, sizeof nm, "se_excite C=%d mid=%d", C, Cmid);
  else snprintf(nm, sizeof nm, "se_excite in=%d mid=%d C=%d (conv3 folded into fc1)", Cin, Cmid, C);
  op->name = nm;
  op->flops = 2.0 * N * Cmid * (Cin + C);
  op->bytes = 4.0 * Cmid * (Cin + C) + 4.0 * N * (Cin + C);
  return submit(plan, op.release(), static_cast<cudaStream_t>(stream));
}

int pcv_se_excite(pcv_plan* plan, int N, int C, int Cmid, const float* pooled, const float* w1, const float* b1,
                  const float* w2, const float* b2, int mid_act, int out_act, float* gate, pcv_stream stream) {
  return pcv_se_excite_ex(plan, N, C, Cmid, C, pooled, w1, b1, w2, b2, mid_act, out_act, gate, stream);
}

int pcv_se_scale_add_act(pcv_plan* plan, int dtype, int N, int HW, int C, const void* x, const float* gate,
                         const void* identity, int act, void* y, pcv_stream stream) {
  PCV_DTYPE_OK(dtype);
  PCV_REQUIRE(x && gate && y, "NULL tensor pointer");
  PCV_REQUIRE(N > 0 && HW > 0 && C > 0 && C % 8 == 0, "SE scale needs C % 8 == 0");
  auto op = std::make_unique<SeScaleOp>();
  op->dtype = dtype; op->N = N; op->HW = HW; op->C = C; op->act = act; op->x = x; op->gate = gate; op->idn = identity;
  op->y = y;
  char nm[96];
  snprintf(nm, sizeof nm, "se_scale_%s C=%d HW=%d%s", dn(dtype), C, HW, identity ? " +id" : "");
  op->name = nm;
  op->bytes = esize(dtype) * static_cast<double>(N) * HW * C * (identity ? 3.0 : 2.0);
  return submit(plan, op.release(), static_cast<cudaStream_t>(stream));
}

int pcv_add_act(pcv_plan* plan, int dtype, size_t count, const void* a, const void* b, int act, void* y,
                pcv_stream stream) {
  PCV_DTYPE_OK(dtype);
  PCV_REQUIRE(a && b && y, "NULL tensor pointer");
  PCV_REQUIRE(count > 0 && count % 8 == 0, "add_act needs count % 8 == 0");
  auto op = std::make_unique<AddActOp>();
  op->dtype = dtype; op->act = act; op->vecs = static_cast<long long>(count / 8); op->a = a; op->b = b; op->y = y;
  op->name = std::string("add_act_") + dn(dtype);
  op->bytes = 3.0 * esize(dtype) * static_cast<double>(count);
  return submit(plan, op.release(), static_cast<cudaStream_t>(stream));
}

int pcv_channel_affine_act(pcv_plan* plan, int dtype, size_t pixels, int C, const void* x, int in_pitch, const float* scale,
                           const float* shift, const float* slope, int act, void* y, int out_pitch, pcv_stream stream) {
  PCV_DTYPE_OK(dtype);
  PCV_REQUIRE(x && y, "NULL tensor pointer");
  in_pitch = pitch_or(in_pitch, C);
  out_pitch = pitch_or(out_pitch, C);
  PCV_REQUIRE(pixels > 0 && C > 0 && C % 8 == 0 && in_pitch >= C && out_pitch >= C, "channel_affine_act needs C %% 8 == 0");
  const int vec = dtype == PCV_F32 ? 4 : 8;   // 16-byte vector accesses
  PCV_REQUIRE(in_pitch % vec == 0 && out_pitch % vec == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(y) % 16 == 0,
              "channel_affine_act needs 16-byte aligned pixel rows");
  for (const float* v : {scale, shift, slope})
    PCV_REQUIRE(reinterpret_cast<uintptr_t>(v) % 16 == 0, "channel_affine_act needs 16-byte aligned channel vectors");
  PCV_REQUIRE(act >= PCV_ACT_NONE && act <= PCV_ACT_HSIGMOID, "unknown activation %d", act);
  auto op = std::make_unique<ChannelAffineOp>();
  op->dtype = dtype; op->C = C; op->in_pitch = in_pitch; op->out_pitch = out_pitch; op->act = act;
  op->pixels = static_cast<long long>(pixels); op->x = x; op->scale = scale; op->shift = shift; op->slope = slope; op->y = y;
  char nm[112];
  snprintf(nm, sizeof nm, "channel_affine_act_%s C=%d px=%lld%s%s act=%d", dn(dtype), C, op->pixels, scale ? " bn" : "",
           slope ? " slope" : "", act);
  op->name = nm;
  op->bytes = 2.0 * esize(dtype) * static_cast<double>(pixels) * C;
  return submit(plan, op.release(), static_cast<cudaStream_t>(stream));
}

int pcv_nchw_to_nhwc_ex(pcv_plan* plan, int dtype, int img_type, int N, int C, int H, int W, const void* x,
                        const float* scale_host, const float* bias_host, void* y, int c_pitch, pcv_stream stream) {
  PCV_DTYPE_OK(dtype);
  PCV_REQUIRE(img_type >= PCV_IMG_F32 && img_type <= PCV_IMG_U8, "unknown image type %d", img_type);
  PCV_REQUIRE(x && y, "NULL tensor pointer");
  c_pitch = pitch_or(c_pitch, round_up(C, 8));
  PCV_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && c_pitch >= C && c_pitch % 8 == 0, "bad ingest dims");
  auto op = std::make_unique<LayoutOp>();
  op->dtype = dtype; op->N = N; op->C = C; op->HW = H * W; op->c_pitch = c_pitch; op->to_nhwc = 1; op->x = x; op->y = y;
  op->img_type = img_type; op->af = make_affine(C, scale_host, bias_host);
  char nm[96];
  snprintf(nm, sizeof nm, "ingest_nchw_%s_to_nhwc_%s C=%d->%d @%dx%d", img_name(img_type), dn(dtype), C, c_pitch, H, W);
  op->name = nm;
  op->bytes = static_cast<double>(N) * H * W * (static_cast<double>(img_esize(img_type)) * C + esize(dtype) * c_pitch);
  return submit(plan, op.release(), static_cast<cudaStream_t>(stream));
}

int pcv_nchw_f32_to_nhwc(pcv_plan* plan, int dtype, int N, int C, int H, int W, const float* x, void* y, int c_pitch,
                         pcv_stream stream) {
  return pcv_nchw_to_nhwc_ex(plan, dtype, PCV_IMG_F32, N, C, H, W, x, nullptr, nullptr, y, c_pitch, stream);
}

int pcv_zero_pad2d(pcv_plan* plan, int dtype, int N, int H, int W, int C, const void* x, int in_pitch, int pad_left,
                   int pad_right, int pad_top, int pad_bottom, void* y, int out_pitch, pcv_stream stream) {
  PCV_DTYPE_OK(dtype);
  PCV_REQUIRE(x && y, "NULL tensor pointer");
  in_pitch = pitch_or(in_pitch, C);
  out_pitch = pitch_or(out_pitch, C);
  PCV_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && pad_left >= 0 && pad_right >= 0 && pad_top >= 0 && pad_bottom >= 0,
              "bad zero-pad dims");
  PCV_REQUIRE(C % 8 == 0 && in_pitch % 8 == 0 && out_pitch % 8 == 0, "zero pad needs channels / pitches that are multiples of 8");
  auto op = std::make_unique<ZeroPadOp>();
  op->dtype = dtype; op->N = N; op->H = H; op->W = W; op->C = C; op->in_pitch = in_pitch; op->pl = pad_left; op->pt = pad_top;
  op->Ho = H + pad_top + pad_bottom; op->Wo = W + pad_left + pad_right; op->out_pitch = out_pitch; op->x = x; op->y = y;
  char nm[96];
  snprintf(nm, sizeof nm, "zero_pad_%s C=%d %dx%d +(l%d r%d t%d b%d)", dn(dtype), C, H, W, pad_left, pad_right, pad_top, pad_bottom);
  op->name = nm;
  op->bytes = static_cast<double>(esize(dtype)) * N * C * (static_cast<double>(H) * W + static_cast<double>(op->Ho) * op->Wo);
  return submit(plan, op.release(), static_cast<cudaStream_t>(stream));
}

int pcv_nhwc_to_nchw_f32(pcv_plan* plan, int dtype, int N, int C, int H, int W, const void* x, int c_pitch, float* y,
                         pcv_stream stream) {
  PCV_DTYPE_OK(dtype);
  PCV_REQUIRE(x && y, "NULL tensor pointer");
  c_pitch = pitch_or(c_pitch, C);
  PCV_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && c_pitch >= C, "bad egress dims");
  auto op = std::make_unique<LayoutOp>();
  op->dtype = dtype; op->N = N; op->C = C; op->HW = H * W; op->c_pitch = c_pitch; op->to_nhwc = 0; op->x = x; op->y = y;
  char nm[96];
  snprintf(nm, sizeof nm, "egress_nhwc_%s_to_nchw_f32 C=%d @%dx%d", dn(dtype), C, H, W);
  op->name = nm;
  op->bytes = static_cast<double>(N) * H * W * C * (4.0 + esize(dtype));
  return submit(plan, op.release(), static_cast<cudaStream_t>(stream));
}

int pcv_bilinear_upsample_ac(pcv_plan* plan, int dtype, int N, int Hin, int Win, int C, const void* x, int in_pitch,
                             int Hout, int Wout, void* y, int out_pitch, int out_nchw_f32, pcv_stream stream) {
  PCV_DTYPE_OK(dtype);
  PCV_REQUIRE(x && y, "NULL tensor pointer");
  PCV_REQUIRE(N > 0 && Hin > 0 && Win > 0 && C > 0 && Hout > 0 && Wout > 0, "bad bilinear dims");
  in_pitch = pitch_or(in_pitch, C);
  out_pitch = pitch_or(out_pitch, C);
  if (!out_nchw_f32)
    PCV_REQUIRE(C % 8 == 0 && in_pitch % 8 == 0 && out_pitch % 8 == 0, "NHWC bilinear needs channels/pitches % 8 == 0");
  auto op = std::make_unique<BilinearOp>();
  op->dtype = dtype; op->N = N; op->Hin = Hin; op->Win = Win; op->C = C; op->in_pitch = in_pitch; op->Hout = Hout;
  op->Wout = Wout; op->out_pitch = out_pitch; op->nchw = out_nchw_f32; op->x = x; op->y = y;
  char nm[96];
  snprintf(nm, sizeof nm, "bilinear_%s C=%d %dx%d->%dx%d%s", dn(dtype), C, Hin, Win, Hout, Wout, out_nchw_f32 ? " nchw_f32" : "");
  op->name = nm;
  op->bytes = static_cast<double>(N) * C * (esize(dtype) * Hin * Win + (out_nchw_f32 ? 4.0 : esize(dtype)) * Hout * Wout);
  return submit(plan, op.release(), static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// space-to-depth stem (see include/pcv_b200.h): ingest + weight re-expression
// ---------------------------------------------------------------------------------------------------------------
namespace pcv {

struct S2dGeom {
  int p, kb, pad_lo, delta, rows, cols;
};
static inline S2dGeom s2d_geom(int H, int W, int k) {
  S2dGeom g;
  g.p = k / 2;
  g.kb = g.p + 1;                       // s2d blocks touched per output pixel and axis
  g.pad_lo = (g.p + 1) / 2;             // zero border before the interior (ceil(p/2)); after: floor(p/2)
  g.delta = 2 * g.pad_lo - g.p;         // filter index = 2*block + parity - delta
  g.rows = H / 2 + g.kb - 1;
  g.cols = W / 2 + g.kb - 1;
  return g;
}

template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) { return pack_bf16x2(lo, hi); }
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) { return pack_f16x2(lo, hi); }

template <typename TI, typename T>
__global__ void __launch_bounds__(256)
s2d_ingest_kernel(int N, int C, int H, int W, int pad_lo, int rows, int cols, const TI* __restrict__ x,
                  T* __restrict__ y, ImgAffine af) {
  const int Hb = H >> 1, Wb = W >> 1;
  const long long total = static_cast<long long>(N) * Hb * Wb;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wb = static_cast<int>(idx % Wb);
    const long long r = idx / Wb;
    const int hb = static_cast<int>(r % Hb);
    const int n = static_cast<int>(r / Hb);
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c >= C) break;
      const TI* px = x + ((static_cast<size_t>(n) * C + c) * H + 2 * hb) * W + 2 * wb;
      float2 top = Img<TI>::ld2(px), bot = Img<TI>::ld2(px + W);
      top.x = fmaf(top.x, af.scale[c], af.bias[c]); top.y = fmaf(top.y, af.scale[c], af.bias[c]);
      bot.x = fmaf(bot.x, af.scale[c], af.bias[c]); bot.y = fmaf(bot.y, af.scale[c], af.bias[c]);
      // channel = (dy*2+dx)*C + c ; C <= 4 so the index is < 16 (static indexing keeps v[] in registers)
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if (e == 0 * C + c) v[e] = top.x;
        if (e == 1 * C + c) v[e] = top.y;
        if (e == 2 * C + c) v[e] = bot.x;
        if (e == 3 * C + c) v[e] = bot.y;
      }
    }
    uint4 lo, hi;
    lo.x = pack2<T>(v[0], v[1]); lo.y = pack2<T>(v[2], v[3]); lo.z = pack2<T>(v[4], v[5]); lo.w = pack2<T>(v[6], v[7]);
    hi.x = pack2<T>(v[8], v[9]); hi.y = pack2<T>(v[10], v[11]); hi.z = pack2<T>(v[12], v[13]); hi.w = pack2<T>(v[14], v[15]);
    uint4* dst = reinterpret_cast<uint4*>(y + ((static_cast<size_t>(n) * rows + hb + pad_lo) * cols + wb + pad_lo) * 16);
    dst[0] = lo;
    dst[1] = hi;
  }
}

// Two s2d pixels (four image columns) per thread: one 16-byte (fp32) / 8-byte (16-bit) / 4-byte (u8) load per image row
// and channel - a warp reads 128 contiguous pixels of each row - 64 contiguous output bytes per thread, 32-bit index
// arithmetic.  C == 3 or 4, W % 4 == 0, x aligned to 4 pixels.
template <int C, typename TI, typename T>
__global__ void __launch_bounds__(256)
s2d_ingest2_kernel(int N, int H, int W, int pad_lo, int rows, int cols, const TI* __restrict__ x, T* __restrict__ y,
                   ImgAffine af) {
  const int Hb = H >> 1, Wq = W >> 2;
  const int total = N * Hb * Wq;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int wq = idx % Wq;
    const int r = idx / Wq;
    const int hb = r % Hb;
    const int n = r / Hb;
    float4 top[C], bot[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const TI* px = x + ((static_cast<size_t>(n) * C + c) * H + 2 * hb) * W + 4 * wq;
      top[c] = Img<TI>::ld4(px);       // the image is read exactly once
      bot[c] = Img<TI>::ld4(px + W);
      const float sc = af.scale[c], bi = af.bias[c];
      top[c].x = fmaf(top[c].x, sc, bi); top[c].y = fmaf(top[c].y, sc, bi); top[c].z = fmaf(top[c].z, sc, bi); top[c].w = fmaf(top[c].w, sc, bi);
      bot[c].x = fmaf(bot[c].x, sc, bi); bot[c].y = fmaf(bot[c].y, sc, bi); bot[c].z = fmaf(bot[c].z, sc, bi); bot[c].w = fmaf(bot[c].w, sc, bi);
    }
    // s2d channel = (dy*2+dx)*C + c, zero padded to 16
    float a[16], b[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) a[e] = b[e] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      a[0 * C + c] = top[c].x; a[1 * C + c] = top[c].y; a[2 * C + c] = bot[c].x; a[3 * C + c] = bot[c].y;
      b[0 * C + c] = top[c].z; b[1 * C + c] = top[c].w; b[2 * C + c] = bot[c].z; b[3 * C + c] = bot[c].w;
    }
    uint4* dst = reinterpret_cast<uint4*>(y + ((static_cast<size_t>(n) * rows + hb + pad_lo) * cols + 2 * wq + pad_lo) * 16);
    dst[0] = make_uint4(pack2<T>(a[0], a[1]), pack2<T>(a[2], a[3]), pack2<T>(a[4], a[5]), pack2<T>(a[6], a[7]));
    dst[1] = make_uint4(pack2<T>(a[8], a[9]), pack2<T>(a[10], a[11]), pack2<T>(a[12], a[13]), pack2<T>(a[14], a[15]));
    dst[2] = make_uint4(pack2<T>(b[0], b[1]), pack2<T>(b[2], b[3]), pack2<T>(b[4], b[5]), pack2<T>(b[6], b[7]));
    dst[3] = make_uint4(pack2<T>(b[8], b[9]), pack2<T>(b[10], b[11]), pack2<T>(b[12], b[13]), pack2<T>(b[14], b[15]));
  }
}

__global__ void s2d_weight_kernel(int Cout, int C, int k, int kb, int delta, const float* __restrict__ w,
                                  float* __restrict__ weq) {
  const int cin_eq = kb * 16;
  const int total = Cout * cin_eq * kb;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx % kb;                 // equivalent filter row (kh = kb, kw = 1)
    const int cc = (idx / kb) % cin_eq;     // equivalent input channel = window pixel j, s2d channel ch
    const int o = idx / (kb * cin_eq);
    const int j = cc / 16, ch = cc % 16;
    float v = 0.f;
    if (ch < 4 * C) {
      const int q = ch / C, c = ch % C;
      const int fr = 2 * r + (q >> 1) - delta, fs = 2 * j + (q & 1) - delta;
      if (fr >= 0 && fr < k && fs >= 0 && fs < k) v = w[((static_cast<size_t>(o) * C + c) * k + fr) * k + fs];
    }
    weq[idx] = v;
  }
}

struct S2dIngestOp : Op {
  int N, C, H, W, dtype, img_type;
  S2dGeom g;
  ImgAffine af;
  const void* x;
  void* y;
  template <typename TI, typename T>
  cudaError_t run(cudaStream_t s) {
    const long long quads = static_cast<long long>(N) * (H / 2) * (W / 4);
    const TI* xi = reinterpret_cast<const TI*>(x);
    T* yo = reinterpret_cast<T*>(y);
    if ((C == 3 || C == 4) && W % 4 == 0 && reinterpret_cast<uintptr_t>(x) % (4 * sizeof(TI)) == 0 && quads < (1ll << 30)) {
      const int grid = grid_for(quads);
      if (C == 3) s2d_ingest2_kernel<3, TI, T><<<grid, 256, 0, s>>>(N, H, W, g.pad_lo, g.rows, g.cols, xi, yo, af);
      else s2d_ingest2_kernel<4, TI, T><<<grid, 256, 0, s>>>(N, H, W, g.pad_lo, g.rows, g.cols, xi, yo, af);
      return cudaGetLastError();
    }
    const int grid = grid_for(static_cast<long long>(N) * (H / 2) * (W / 2));
    s2d_ingest_kernel<TI, T><<<grid, 256, 0, s>>>(N, C, H, W, g.pad_lo, g.rows, g.cols, xi, yo, af);
    return cudaGetLastError();
  }
  template <typename TI>
  cudaError_t run_in(cudaStream_t s) { return dtype == PCV_F16 ? run<TI, __half>(s) : run<TI, __nv_bfloat16>(s); }
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    switch (img_type) {
      case PCV_IMG_BF16: return run_in<__nv_bfloat16>(s);
      case PCV_IMG_F16: return run_in<__half>(s);
      case PCV_IMG_U8: return run_in<uint8_t>(s);
      default: return run_in<float>(s);
    }
  }
};

}  // namespace pcv

extern "C" {

int pcv_stem_s2d_dims(int C, int H, int W, int k, int* rows, int* cols, int* cin_eq, int* taps_eq) {
  PCV_REQUIRE(C >= 1 && C <= 4 && (k == 3 || k == 5 || k == 7) && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0,
              "space-to-depth stem needs C <= 4, k in {3,5,7}, even H and W (got C=%d k=%d %dx%d)", C, k, H, W);
  const S2dGeom g = s2d_geom(H, W, k);
  if (rows) *rows = g.rows;
  if (cols) *cols = g.cols;
  if (cin_eq) *cin_eq = g.kb * 16;
  if (taps_eq) *taps_eq = g.kb;
  return PCV_OK;
}

int pcv_stem_s2d_ingest_ex(pcv_plan* plan, int dtype, int img_type, int N, int C, int H, int W, int k, const void* x,
                           const float* scale_host, const float* bias_host, void* s2d, pcv_stream stream) {
  if (int rc = pcv_stem_s2d_dims(C, H, W, k, nullptr, nullptr, nullptr, nullptr)) return rc;
  PCV_REQUIRE(dtype == PCV_BF16 || dtype == PCV_F16, "the space-to-depth stem exists in the 16-bit tiers only");
  PCV_REQUIRE(img_type >= PCV_IMG_F32 && img_type <= PCV_IMG_U8, "unknown image type %d", img_type);
  PCV_REQUIRE(x && s2d && N > 0, "NULL tensor pointer");
  PCV_REQUIRE(reinterpret_cast<uintptr_t>(x) % (2 * img_esize(img_type)) == 0 && reinterpret_cast<uintptr_t>(s2d) % 16 == 0,
              "misaligned stem tensors");
  auto op = std::make_unique<S2dIngestOp>();
  op->N = N; op->C = C; op->H = H; op->W = W; op->g = s2d_geom(H, W, k); op->x = x; op->y = s2d;
  op->dtype = dtype; op->img_type = img_type; op->af = make_affine(C, scale_host, bias_host);
  char nm[96];
  snprintf(nm, sizeof nm, "ingest_s2d_%s_to_%s C=%d k=%d @%dx%d", img_name(img_type), dn(dtype), C, k, H, W);
  op->name = nm;
  op->bytes = static_cast<double>(N) * (static_cast<double>(img_esize(img_type)) * C * H * W + 32.0 * (H / 2) * (W / 2));
  return submit(plan, op.release(), static_cast<cudaStream_t>(stream));
}

int pcv_stem_s2d_ingest(pcv_plan* plan, int N, int C, int H, int W, int k, const float* x, void* s2d,
                        pcv_stream stream) {
  return pcv_stem_s2d_ingest_ex(plan, PCV_BF16, PCV_IMG_F32, N, C, H, W, k, x, nullptr, nullptr, s2d, stream);
}

int pcv_stem_s2d_pool_ok(int C, int H, int W, int k, int Cout) { return pcv::bf::stem_pool_ok(C, H, W, k, Cout); }

int pcv_stem_s2d_weights(int Cout, int C, int k, const float* w, float* w_eq, pcv_stream stream) {
  PCV_REQUIRE(w && w_eq && Cout > 0 && C >= 1 && C <= 4 && (k == 3 || k == 5 || k == 7), "bad stem weight arguments");
  const S2dGeom g = s2d_geom(2, 2, k);
  const int total = Cout * g.kb * 16 * g.kb;
  s2d_weight_kernel<<<ceil_div(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(Cout, C, k, g.kb, g.delta, w, w_eq);
  g_launches++;
  PCV_CHECK_CUDA(cudaGetLastError());
  return PCV_OK;
}

}  // extern "C"
