// Internal runtime shared by the kernels' host launchers and the C ABI: the Op record, the Plan, error plumbing.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "../../include/pcv_b200.h"

namespace pcv {

// ---- errors (thread-local message, integer status across the ABI) -------------------------------------------
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);

#define PCV_CHECK_CUDA(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return ::pcv::fail(PCV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                                 \
  } while (0)

#define PCV_REQUIRE(cond, ...)                                 \
  do {                                                         \
    if (!(cond)) return ::pcv::fail(PCV_ERR_INVALID, __VA_ARGS__); \
  } while (0)

extern std::atomic<int64_t> g_launches;

// ---- one recorded fused-kernel launch --------------------------------------------------------------------------
struct Op {
  std::string name;
  double flops = 0.0;   // algorithmic FLOPs (2*MACs) of this op
  double bytes = 0.0;   // algorithmic HBM bytes of this op
  int launches = 1;     // kernels per launch() call
  virtual ~Op() {}
  virtual cudaError_t launch(cudaStream_t s) = 0;
};

}  // namespace pcv

struct pcv_plan {
  std::vector<std::unique_ptr<pcv::Op>> ops;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  int launches = 0;
};

namespace pcv {

// Run the op now (plan == nullptr) or append it to the plan.  Takes ownership of `op`.
int submit(pcv_plan* plan, Op* op, cudaStream_t stream);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }
inline int conv_out(int H, int k, int stride, int pad, int dil) { return (H + 2 * pad - dil * (k - 1) - 1) / stride + 1; }
inline int pitch_or(int pitch, int c) { return pitch > 0 ? pitch : c; }
inline size_t esize(int dtype) { return dtype == PCV_F32 ? 4 : 2; }
inline const char* dtype_name(int dtype) { return dtype == PCV_F32 ? "f32" : (dtype == PCV_F16 ? "f16" : "bf16"); }
int sm_count();   // SMs of the CURRENT device (cached per device)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE function attribute: `done` holds one bit per device
// ordinal for one kernel instantiation, so a process that drives several GPUs sets it on each of them.
inline cudaError_t set_max_smem_once(const void* func, int bytes, std::atomic<uint64_t>& done) {
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev)) return e;
  const uint64_t bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  if (cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) return e;
  done.fetch_or(bit, std::memory_order_release);
  return cudaSuccess;
}
template <typename... KArgs>
inline cudaError_t set_max_smem_once(void (*kernel)(KArgs...), int bytes, std::atomic<uint64_t>& done) {
  return set_max_smem_once(reinterpret_cast<const void*>(kernel), bytes, done);
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// The hot kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization so that the next kernel's
// prologue (barrier init, TMEM allocation, descriptor prefetch, resident-weight loads) overlaps the tail of the
// previous one; every such kernel executes griddepcontrol.wait before touching activation memory.  PCV_PDL=0 disables.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- driver entry points for tensor maps (resolved at run time so the .so loads without libcuda) ----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();
EncodeIm2colFn encode_im2col_fn();

// ---- launchers implemented in the .cu files --------------------------------------------------------------------
// The tensor-core and TMA-window kernels exist once per 16-bit storage tier (ptx.cuh, "16-bit storage tier"): the same
// sources compiled into pcv::bf (bf16) and pcv::hf (IEEE fp16).
struct IgemmDual {   // second source of a dual-source 1x1 conv (pcv_conv1x1_dual): y = act(W1 x1 + W2 x2[::s] + b)
  const pcv_conv_desc* d2;
  const void* x2;
  const float* bias2;   // gated variant only (PCV_CONV_SE_GATE on the first desc): the shortcut's bias, added outside the gate
};
#define PCV_DECLARE_TIER_API(NS)                                                                                         \
  namespace NS {                                                                                                         \
  /* conv_igemm.cu : tcgen05 implicit GEMM (dense + block-diagonal grouped) */                                           \
  int igemm_supported(const pcv_conv_desc& d, std::string* why);                                                         \
  int igemm_packed_bytes(const pcv_conv_desc& d, size_t* w_bytes, size_t* b_bytes);                                      \
  int igemm_pack(const pcv_conv_desc& d, const float* w, const float* conv_bias, const float* g, const float* b,        \
                 const float* m, const float* v, float eps, void* w_packed, float* bias_out, cudaStream_t s);            \
  int igemm_make(const pcv_conv_desc& d, const void* x, const void* w, const float* bias, const void* res, void* y,     \
                 Op** out, const float* gate = nullptr, const IgemmDual* dual = nullptr);                                \
  int igemm_gate_ok(const pcv_conv_desc& d);                                                                             \
  int igemm_dual_ok(const pcv_conv_desc& d, const pcv_conv_desc& d2);                                                    \
  /* window_tma.cu : TMA halo-staged 3x3 / 5x5 window ops (op_kind 0 = depthwise conv, 1 = max pool).  Returns */        \
  /* PCV_ERR_UNSUPPORTED (without touching the error message) for shapes the caller serves with its generic kernel. */   \
  int win_make(int op_kind, int N, int H, int W, int C, int k, int stride, int pad, int act, const void* x,              \
               int in_pitch, const float* w, const float* bias, const void* res, int res_pitch, void* y, int out_pitch,  \
               Op** out);                                                                                                \
  /* conv_igemm3x.cu : bottleneck tail  relu(conv3_1x1(relu(conv2_3x3(x))) + identity)  in one kernel */               \
  int fused_tail_ok(const pcv_conv_desc& d2, const pcv_conv_desc& d3);                                                   \
  int fused_tail_make(const pcv_conv_desc& d2, const pcv_conv_desc& d3, const void* x, const void* w2,                  \
                      const float* bias2, const void* w3, const float* bias3, const void* res, void* y, Op** out);      \
  /* conv_dwpw.cu : depthwise 3x3 -> pointwise 1x1 in one kernel (the depthwise tensor stays in shared memory) */       \
  int dwpw_ok(const pcv_conv_desc& dw, const pcv_conv_desc& pw);                                                         \
  int dwpw_make(const pcv_conv_desc& dw, const pcv_conv_desc& pw, const void* x, const float* w_dw, const float* b_dw,  \
                const void* w_pw, const float* b_pw, const void* res, void* y, Op** out);                                \
  /* conv_xdwpw.cu : 1x1 expansion -> depthwise 3x3 -> pointwise 1x1 in one kernel (neither wide tensor touches HBM) */           \
  int xdwpw_ok(const pcv_conv_desc& ex, const pcv_conv_desc& dw, const pcv_conv_desc& pw);                                       \
  int xdwpw_make(const pcv_conv_desc& ex, const pcv_conv_desc& dw, const pcv_conv_desc& pw, const void* x, const void* w_ex,    \
                 const float* b_ex, const float* w_dw, const float* b_dw, const void* w_pw, const float* b_pw,                   \
                 const void* res, void* y, Op** out);                                                                            \
  /* conv_igemm3s.cu : can the s2d stem take the fused max pool (pcv_stem_s2d_pool_ok) */                                \
  int stem_pool_ok(int C, int H, int W, int k, int Cout);                                                                \
  }
PCV_DECLARE_TIER_API(bf)
PCV_DECLARE_TIER_API(hf)
#undef PCV_DECLARE_TIER_API
namespace bf {
// conv_igemm.cu : the fp32 tier on the tensor cores (PCV_CONV_F32_SPLIT): 3-way bf16 split of fp32 operands
int igemm_split_supported(const pcv_conv_desc& d, std::string* why);
int igemm_split_packed_bytes(const pcv_conv_desc& d, size_t* w_bytes, size_t* b_bytes, size_t* ws_bytes);
int igemm_split_pack(const pcv_conv_desc& d, const float* w, const float* conv_bias, const float* g, const float* b,
                     const float* m, const float* v, float eps, void* w_packed, float* bias_out, cudaStream_t s);
int igemm_split_make(const pcv_conv_desc& d, const void* x, const void* w, const float* bias, const void* res, void* y,
                     void* workspace, Op** out);
}  // namespace bf
inline bool is16(int dtype) { return dtype == PCV_BF16 || dtype == PCV_F16; }
// conv_simt.cu : CUDA-core direct convolution (fp32 tier; generic bf16 fallback / cross-check)
int simt_packed_bytes(const pcv_conv_desc& d, int dtype, size_t* w_bytes, size_t* b_bytes);
int simt_pack(const pcv_conv_desc& d, int dtype, const float* w, const float* conv_bias, const float* g,
              const float* b, const float* m, const float* v, float eps, void* w_packed, float* bias_out,
              cudaStream_t s);
int simt_make(const pcv_conv_desc& d, int dtype, const void* x, const void* w, const float* bias, const void* res,
              void* y, Op** out);
// dwconv.cu : depthwise kxk, bf16 / fp32, NHWC vectorised
int dw_packed_bytes(const pcv_conv_desc& d, int dtype, size_t* w_bytes, size_t* b_bytes);
int dw_pack(const pcv_conv_desc& d, int dtype, const float* w, const float* conv_bias, const float* g, const float* b,
            const float* m, const float* v, float eps, void* w_packed, float* bias_out, cudaStream_t s);
int dw_make(const pcv_conv_desc& d, int dtype, const void* x, const void* w, const float* bias, const void* res,
            void* y, Op** out);

enum ConvRoute { ROUTE_IGEMM = 0, ROUTE_DW = 1, ROUTE_SIMT = 2, ROUTE_SPLIT = 3 };
int conv_route(const pcv_conv_desc& d, int dtype, std::string* why);

}  // namespace pcv
