// Host runtime of libpcv_b200: error plumbing, plan execution (eager / CUDA graph / per-op profile), conv routing.
#include "runtime.h"

#include <algorithm>
#include <cstring>
#include <mutex>

namespace pcv {

std::atomic<int64_t> g_launches{0};

std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("PCV_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

int sm_count() {
  static std::atomic<int> cache[64];   // per device ordinal; 0 = not queried yet
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::atomic<int>& slot = cache[dev & 63];
  int n = slot.load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    slot.store(n, std::memory_order_relaxed);
  }
  return n;
}

static void* driver_entry(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  return fn;
}
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(driver_entry("cuTensorMapEncodeTiled"));
  return fn;
}
EncodeIm2colFn encode_im2col_fn() {
  static EncodeIm2colFn fn = reinterpret_cast<EncodeIm2colFn>(driver_entry("cuTensorMapEncodeIm2col"));
  return fn;
}

int submit(pcv_plan* plan, Op* op, cudaStream_t stream) {
  std::unique_ptr<Op> holder(op);
  if (plan) {
    if (plan->graph_exec) return fail(PCV_ERR_INVALID, "plan already captured into a CUDA graph; create a new plan");
    plan->launches += op->launches;
    plan->ops.push_back(std::move(holder));
    return PCV_OK;
  }
  cudaError_t e = op->launch(stream);
  if (e != cudaSuccess) return fail(PCV_ERR_CUDA, "%s: launch failed: %s", op->name.c_str(), cudaGetErrorString(e));
  return PCV_OK;
}

// Which kernel family serves a ConvBlock in a tier.  bf16: depthwise -> dwconv, tcgen05-compatible dense/grouped ->
// implicit GEMM, anything else -> CUDA-core direct conv.  fp32: depthwise -> dwconv, else CUDA-core direct conv.
int conv_route(const pcv_conv_desc& d, int dtype, std::string* why) {
  const bool depthwise = d.groups > 1 && d.groups == d.Cin && d.Cin == d.Cout;
  // the depthwise kernels take square windows and store in the tier's own type; anything else falls to the generic kernel
  if (depthwise && d.kh == d.kw && !(d.flags & PCV_CONV_OUT_F32) && pitch_or(d.in_pitch, d.Cin) % 8 == 0 &&
      pitch_or(d.out_pitch, d.Cout) % 8 == 0 && d.Cin % 8 == 0)
    return ROUTE_DW;
  if (is16(dtype) && !(d.flags & PCV_CONV_FORCE_SIMT) && bf::igemm_supported(d, why)) return ROUTE_IGEMM;
  if (dtype == PCV_F32 && (d.flags & PCV_CONV_F32_SPLIT) && bf::igemm_split_supported(d, why)) return ROUTE_SPLIT;
  return ROUTE_SIMT;  // (validate_conv rejects overlapped / row-pitched views that reach this route)
}

static int validate_conv(const pcv_conv_desc* d, int dtype) {
  PCV_REQUIRE(d != nullptr, "conv desc is NULL");
  PCV_REQUIRE(dtype == PCV_BF16 || dtype == PCV_F32 || dtype == PCV_F16, "unknown dtype %d", dtype);
  PCV_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "non-positive conv dims");
  PCV_REQUIRE(d->kh > 0 && d->kw > 0 && d->stride > 0 && d->dil > 0 && d->pad >= 0, "bad kernel/stride/pad/dilation");
  PCV_REQUIRE(d->groups > 0 && d->Cin % d->groups == 0 && d->Cout % d->groups == 0,
              "channels (%d -> %d) not divisible by groups=%d", d->Cin, d->Cout, d->groups);
  PCV_REQUIRE(d->act >= PCV_ACT_NONE && d->act <= PCV_ACT_LEAKY_RELU, "unknown activation %d", d->act);
  if (d->act == PCV_ACT_LEAKY_RELU) {
    std::string why;
    PCV_REQUIRE(conv_route(*d, dtype, &why) != ROUTE_DW,
                "the LeakyReLU epilogue serves dense / grouped convs; a depthwise conv takes pcv_channel_affine_act behind it");
  }
  PCV_REQUIRE((pitch_or(d->in_pitch, d->Cin) >= d->Cin || (d->flags & PCV_CONV_IN_OVERLAP)) &&
                  pitch_or(d->out_pitch, d->Cout) >= d->Cout,
              "channel pitch smaller than channel count");
  PCV_REQUIRE(!(d->flags & PCV_CONV_IN_OVERLAP) || (is16(dtype) && d->groups == 1),
              "overlapping input views are only supported by the bf16 tensor-core path");
  PCV_REQUIRE(d->in_row_pitch == 0 || d->in_row_pitch >= (d->W - 1) * pitch_or(d->in_pitch, d->Cin) + d->Cin ||
                  (d->flags & PCV_CONV_IN_OVERLAP),
              "in_row_pitch smaller than a row");
  PCV_REQUIRE(!(d->flags & PCV_CONV_OUT_F32) || is16(dtype), "OUT_F32 only applies to the 16-bit tiers");
  PCV_REQUIRE(!(d->flags & PCV_CONV_F32_SPLIT) || dtype == PCV_F32, "F32_SPLIT only applies to the fp32 tier");
  if (d->flags & PCV_CONV_SE_GATE) {
    std::string why;
    PCV_REQUIRE(is16(dtype) && conv_route(*d, dtype, &why) == ROUTE_IGEMM && bf::igemm_gate_ok(*d),
                "PCV_CONV_SE_GATE: layer outside the gated-epilogue kernel's domain (ask pcv_conv_se_gate_ok)");
  }
  PCV_REQUIRE(!(d->flags & PCV_CONV_POOL3S2) || ((d->flags & PCV_CONV_IN_OVERLAP) && is16(dtype)),
              "POOL3S2 only applies to the bf16 space-to-depth stem");
  if ((d->flags & PCV_CONV_IN_OVERLAP) || d->in_row_pitch != 0) {
    std::string why;
    PCV_REQUIRE(conv_route(*d, dtype, &why) == ROUTE_IGEMM, "row-pitched input views need the tcgen05 route: %s",
                why.c_str());
  }
  return PCV_OK;
}

}  // namespace pcv

using namespace pcv;

extern "C" {

const char* pcv_last_error(void) { return last_error_ref().c_str(); }
int pcv_version(void) { return 100; }
int64_t pcv_launch_count(void) { return g_launches.load(); }

int pcv_device_info(int dev, int* sm_arch, int* sm_cnt, size_t* hbm_bytes) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= dev)
    return fail(PCV_ERR_NO_DEVICE, "no CUDA device %d (this path has no CPU fallback)", dev);
  cudaDeviceProp prop;
  PCV_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_arch) *sm_arch = prop.major * 10 + prop.minor;
  if (sm_cnt) *sm_cnt = prop.multiProcessorCount;
  if (hbm_bytes) *hbm_bytes = prop.totalGlobalMem;
  return PCV_OK;
}

int pcv_conv_packed_bytes(const pcv_conv_desc* d, int dtype, size_t* w_bytes, size_t* bias_bytes) {
  if (int rc = validate_conv(d, dtype)) return rc;
  PCV_REQUIRE(w_bytes && bias_bytes, "NULL output pointer");
  switch (conv_route(*d, dtype, nullptr)) {
    case ROUTE_DW: return dw_packed_bytes(*d, dtype, w_bytes, bias_bytes);
    case ROUTE_IGEMM: return bf::igemm_packed_bytes(*d, w_bytes, bias_bytes);
    case ROUTE_SPLIT: return bf::igemm_split_packed_bytes(*d, w_bytes, bias_bytes, nullptr);
    default: return simt_packed_bytes(*d, dtype, w_bytes, bias_bytes);
  }
}

int pcv_pack_conv_weights(const pcv_conv_desc* d, int dtype, const float* w, const float* conv_bias,
                          const float* bn_gamma, const float* bn_beta, const float* bn_mean, const float* bn_var,
                          float eps, void* w_packed, float* bias_out, pcv_stream stream) {
  if (int rc = validate_conv(d, dtype)) return rc;
  PCV_REQUIRE(w && w_packed && bias_out, "NULL weight pointer");
  const int nbn = (bn_gamma != nullptr) + (bn_beta != nullptr) + (bn_mean != nullptr) + (bn_var != nullptr);
  PCV_REQUIRE(nbn == 0 || nbn == 4, "BatchNorm needs all of gamma/beta/mean/var or none");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (conv_route(*d, dtype, nullptr)) {
    case ROUTE_DW: return dw_pack(*d, dtype, w, conv_bias, bn_gamma, bn_beta, bn_mean, bn_var, eps, w_packed, bias_out, s);
    case ROUTE_SPLIT: return bf::igemm_split_pack(*d, w, conv_bias, bn_gamma, bn_beta, bn_mean, bn_var, eps, w_packed, bias_out, s);
    case ROUTE_IGEMM:
      return (dtype == PCV_F16 ? hf::igemm_pack : bf::igemm_pack)(*d, w, conv_bias, bn_gamma, bn_beta, bn_mean, bn_var, eps,
                                                                  w_packed, bias_out, s);
    default: return simt_pack(*d, dtype, w, conv_bias, bn_gamma, bn_beta, bn_mean, bn_var, eps, w_packed, bias_out, s);
  }
}

int pcv_conv_workspace_bytes(const pcv_conv_desc* d, int dtype, size_t* bytes) {
  if (int rc = validate_conv(d, dtype)) return rc;
  PCV_REQUIRE(bytes != nullptr, "NULL output pointer");
  *bytes = 0;
  if (conv_route(*d, dtype, nullptr) == ROUTE_SPLIT) return bf::igemm_split_packed_bytes(*d, nullptr, nullptr, bytes);
  return PCV_OK;
}

int pcv_conv_se_gate_ok(const pcv_conv_desc* d, int dtype) {
  if (!d || !is16(dtype) || !(d->flags & PCV_CONV_SE_GATE)) return 0;
  return validate_conv(d, dtype) == PCV_OK ? 1 : 0;
}

int pcv_conv2d_bias_act_ws(pcv_plan* plan, const pcv_conv_desc* d, int dtype, const void* x, const void* w_packed,
                           const float* bias, const void* residual, void* y, void* workspace, pcv_stream stream) {
  if (int rc = validate_conv(d, dtype)) return rc;
  PCV_REQUIRE(x && w_packed && bias && y, "NULL tensor pointer");
  Op* op = nullptr;
  int rc;
  switch (conv_route(*d, dtype, nullptr)) {
    case ROUTE_DW: rc = dw_make(*d, dtype, x, w_packed, bias, residual, y, &op); break;
    case ROUTE_IGEMM:
      rc = (dtype == PCV_F16 ? hf::igemm_make : bf::igemm_make)(*d, x, w_packed, bias, residual, y, &op,
                                                                (d->flags & PCV_CONV_SE_GATE) ? static_cast<const float*>(workspace) : nullptr, nullptr);
      break;
    case ROUTE_SPLIT: rc = bf::igemm_split_make(*d, x, w_packed, bias, residual, y, workspace, &op); break;
    default: rc = simt_make(*d, dtype, x, w_packed, bias, residual, y, &op); break;
  }
  if (rc) return rc;
  return submit(plan, op, static_cast<cudaStream_t>(stream));
}

int pcv_conv2d_bias_act(pcv_plan* plan, const pcv_conv_desc* d, int dtype, const void* x, const void* w_packed,
                        const float* bias, const void* residual, void* y, pcv_stream stream) {
  return pcv_conv2d_bias_act_ws(plan, d, dtype, x, w_packed, bias, residual, y, nullptr, stream);
}

int pcv_bottleneck_tail_fusable(const pcv_conv_desc* d2, const pcv_conv_desc* d3, int dtype) {
  if (!d2 || !d3 || !is16(dtype)) return 0;
  if (validate_conv(d2, dtype) != PCV_OK || validate_conv(d3, dtype) != PCV_OK) return 0;
  return bf::fused_tail_ok(*d2, *d3);
}

int pcv_bottleneck_tail(pcv_plan* plan, const pcv_conv_desc* d2, const pcv_conv_desc* d3, int dtype, const void* x,
                        const void* w2_packed, const float* bias2, const void* w3_packed, const float* bias3,
                        const void* residual, void* y, pcv_stream stream) {
  if (int rc = validate_conv(d2, dtype)) return rc;
  if (int rc = validate_conv(d3, dtype)) return rc;
  PCV_REQUIRE(is16(dtype), "the fused bottleneck tail exists in the 16-bit tiers only");
  Op* op = nullptr;
  const int rc = (dtype == PCV_F16 ? hf::fused_tail_make : bf::fused_tail_make)(*d2, *d3, x, w2_packed, bias2, w3_packed, bias3,
                                                                                residual, y, &op);
  if (rc) return rc;
  return submit(plan, op, static_cast<cudaStream_t>(stream));
}

int pcv_dw_pw_fusable(const pcv_conv_desc* dw, const pcv_conv_desc* pw, int dtype) {
  if (!dw || !pw || !is16(dtype)) return 0;
  if (validate_conv(dw, dtype) != PCV_OK || validate_conv(pw, dtype) != PCV_OK) return 0;
  return bf::dwpw_ok(*dw, *pw);
}

int pcv_dw_pw_fused(pcv_plan* plan, const pcv_conv_desc* dw, const pcv_conv_desc* pw, int dtype, const void* x,
                    const void* w_dw_packed, const float* bias_dw, const void* w_pw_packed, const float* bias_pw,
                    const void* residual, void* y, pcv_stream stream) {
  if (int rc = validate_conv(dw, dtype)) return rc;
  if (int rc = validate_conv(pw, dtype)) return rc;
  PCV_REQUIRE(is16(dtype), "the fused depthwise -> pointwise kernel exists in the 16-bit tiers only");
  Op* op = nullptr;
  const int rc = (dtype == PCV_F16 ? hf::dwpw_make : bf::dwpw_make)(*dw, *pw, x, reinterpret_cast<const float*>(w_dw_packed),
                                                                    bias_dw, w_pw_packed, bias_pw, residual, y, &op);
  if (rc) return rc;
  return submit(plan, op, static_cast<cudaStream_t>(stream));
}

int pcv_exp_dw_pw_fusable(const pcv_conv_desc* ex, const pcv_conv_desc* dw, const pcv_conv_desc* pw, int dtype) {
  if (!ex || !dw || !pw || !is16(dtype)) return 0;
  if (validate_conv(ex, dtype) != PCV_OK || validate_conv(dw, dtype) != PCV_OK || validate_conv(pw, dtype) != PCV_OK) return 0;
  std::string why;
  // the kernel reads the expansion / projection weights in the tcgen05 route's packed layout, the depthwise ones in dwconv's
  if (conv_route(*ex, dtype, &why) != ROUTE_IGEMM || conv_route(*pw, dtype, &why) != ROUTE_IGEMM ||
      conv_route(*dw, dtype, &why) != ROUTE_DW)
    return 0;
  return bf::xdwpw_ok(*ex, *dw, *pw);
}

int pcv_exp_dw_pw_fused(pcv_plan* plan, const pcv_conv_desc* ex, const pcv_conv_desc* dw, const pcv_conv_desc* pw, int dtype,
                        const void* x, const void* w_ex_packed, const float* bias_ex, const void* w_dw_packed,
                        const float* bias_dw, const void* w_pw_packed, const float* bias_pw, const void* residual, void* y,
                        pcv_stream stream) {
  if (int rc = validate_conv(ex, dtype)) return rc;
  if (int rc = validate_conv(dw, dtype)) return rc;
  if (int rc = validate_conv(pw, dtype)) return rc;
  PCV_REQUIRE(is16(dtype), "the fused expansion -> depthwise -> pointwise kernel exists in the 16-bit tiers only");
  Op* op = nullptr;
  const int rc = (dtype == PCV_F16 ? hf::xdwpw_make : bf::xdwpw_make)(
      *ex, *dw, *pw, x, w_ex_packed, bias_ex, reinterpret_cast<const float*>(w_dw_packed), bias_dw, w_pw_packed, bias_pw,
      residual, y, &op);
  if (rc) return rc;
  return submit(plan, op, static_cast<cudaStream_t>(stream));
}

int pcv_conv1x1_dual_ok(const pcv_conv_desc* d, const pcv_conv_desc* d2, int dtype) {
  if (!d || !d2 || !is16(dtype)) return 0;
  if (validate_conv(d, dtype) != PCV_OK || validate_conv(d2, dtype) != PCV_OK) return 0;
  std::string why;
  if (conv_route(*d, dtype, &why) != ROUTE_IGEMM || conv_route(*d2, dtype, &why) != ROUTE_IGEMM) return 0;
  return bf::igemm_dual_ok(*d, *d2);
}

int pcv_conv1x1_dual(pcv_plan* plan, const pcv_conv_desc* d, const pcv_conv_desc* d2, int dtype, const void* x, const void* x2,
                     const void* w_cat_packed, const float* bias_sum, void* y, pcv_stream stream) {
  if (int rc = validate_conv(d, dtype)) return rc;
  if (int rc = validate_conv(d2, dtype)) return rc;
  PCV_REQUIRE(is16(dtype), "the dual-source 1x1 conv exists in the 16-bit tiers only");
  PCV_REQUIRE(x && x2 && w_cat_packed && bias_sum && y, "NULL tensor pointer");
  PCV_REQUIRE(pcv_conv1x1_dual_ok(d, d2, dtype), "layer pair outside the dual-source kernel's domain (ask pcv_conv1x1_dual_ok)");
  Op* op = nullptr;
  PCV_REQUIRE(!(d->flags & PCV_CONV_SE_GATE), "gated unit: use pcv_conv1x1_dual_se");
  const IgemmDual dual{d2, x2, nullptr};
  const int rc = (dtype == PCV_F16 ? hf::igemm_make : bf::igemm_make)(*d, x, w_cat_packed, bias_sum, nullptr, y, &op, nullptr, &dual);
  if (rc) return rc;
  return submit(plan, op, static_cast<cudaStream_t>(stream));
}

int pcv_conv1x1_dual_se(pcv_plan* plan, const pcv_conv_desc* d, const pcv_conv_desc* d2, int dtype, const void* x, const void* x2,
                        const void* w_cat_packed, const float* bias, const float* bias2, const float* gate, void* y,
                        pcv_stream stream) {
  if (int rc = validate_conv(d, dtype)) return rc;
  if (int rc = validate_conv(d2, dtype)) return rc;
  PCV_REQUIRE(is16(dtype), "the dual-source 1x1 conv exists in the 16-bit tiers only");
  PCV_REQUIRE(x && x2 && w_cat_packed && bias && bias2 && gate && y, "NULL tensor pointer");
  PCV_REQUIRE((d->flags & PCV_CONV_SE_GATE) && pcv_conv1x1_dual_ok(d, d2, dtype),
              "layer pair outside the gated dual-source kernel's domain (PCV_CONV_SE_GATE on d, ask pcv_conv1x1_dual_ok)");
  PCV_REQUIRE(reinterpret_cast<uintptr_t>(gate) % 16 == 0, "the SE gate must be 16-byte aligned");
  Op* op = nullptr;
  const IgemmDual dual{d2, x2, bias2};
  const int rc = (dtype == PCV_F16 ? hf::igemm_make : bf::igemm_make)(*d, x, w_cat_packed, bias, nullptr, y, &op, gate, &dual);
  if (rc) return rc;
  return submit(plan, op, static_cast<cudaStream_t>(stream));
}

int pcv_plan_create(pcv_plan** plan) {
  PCV_REQUIRE(plan != nullptr, "NULL plan out-pointer");
  *plan = new pcv_plan();
  return PCV_OK;
}

int pcv_plan_destroy(pcv_plan* plan) {
  if (!plan) return PCV_OK;
  if (plan->graph_exec) cudaGraphExecDestroy(plan->graph_exec);
  if (plan->graph) cudaGraphDestroy(plan->graph);
  delete plan;
  return PCV_OK;
}

int pcv_plan_num_ops(const pcv_plan* plan) { return plan ? static_cast<int>(plan->ops.size()) : 0; }
int pcv_plan_num_launches(const pcv_plan* plan) { return plan ? plan->launches : 0; }

int pcv_plan_run(pcv_plan* plan, pcv_stream stream) {
  PCV_REQUIRE(plan != nullptr, "NULL plan");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (auto& op : plan->ops) {
    cudaError_t e = op->launch(s);
    if (e != cudaSuccess) return fail(PCV_ERR_CUDA, "%s: launch failed: %s", op->name.c_str(), cudaGetErrorString(e));
  }
  return PCV_OK;
}

int pcv_plan_graph_launch(pcv_plan* plan, pcv_stream stream) {
  PCV_REQUIRE(plan != nullptr, "NULL plan");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!plan->graph_exec) {
    PCV_REQUIRE(s != nullptr, "graph capture needs a non-default stream");
    // warm every kernel variant once so lazy function-attribute setup does not happen during capture
    if (int rc = pcv_plan_run(plan, stream)) return rc;
    PCV_CHECK_CUDA(cudaStreamSynchronize(s));
    const int64_t before = g_launches.load();
    PCV_CHECK_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int rc = pcv_plan_run(plan, stream);
    cudaError_t e = cudaStreamEndCapture(s, &plan->graph);
    g_launches.store(before);  // capture enqueued nothing
    if (rc) return rc;
    if (e != cudaSuccess) return fail(PCV_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    PCV_CHECK_CUDA(cudaGraphInstantiate(&plan->graph_exec, plan->graph, 0));
  }
  PCV_CHECK_CUDA(cudaGraphLaunch(plan->graph_exec, s));
  g_launches += plan->launches;
  return PCV_OK;
}

int pcv_plan_profile(pcv_plan* plan, pcv_stream stream, float* ms_host, int capacity) {
  PCV_REQUIRE(plan != nullptr && ms_host != nullptr, "NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = static_cast<int>(plan->ops.size());
  PCV_REQUIRE(capacity >= n, "capacity %d < %d ops", capacity, n);
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) PCV_CHECK_CUDA(cudaEventCreate(&e));
  PCV_CHECK_CUDA(cudaEventRecord(ev[0], s));
  for (int i = 0; i < n; ++i) {
    cudaError_t e = plan->ops[i]->launch(s);
    if (e != cudaSuccess) return fail(PCV_ERR_CUDA, "%s: launch failed: %s", plan->ops[i]->name.c_str(), cudaGetErrorString(e));
    PCV_CHECK_CUDA(cudaEventRecord(ev[i + 1], s));
  }
  PCV_CHECK_CUDA(cudaStreamSynchronize(s));
  for (int i = 0; i < n; ++i) PCV_CHECK_CUDA(cudaEventElapsedTime(&ms_host[i], ev[i], ev[i + 1]));
  for (auto& e : ev) cudaEventDestroy(e);
  return PCV_OK;
}

const char* pcv_plan_op_name(const pcv_plan* plan, int i) {
  if (!plan || i < 0 || i >= static_cast<int>(plan->ops.size())) return "";
  return plan->ops[i]->name.c_str();
}

int pcv_plan_op_cost(const pcv_plan* plan, int i, double* flops, double* bytes) {
  PCV_REQUIRE(plan && i >= 0 && i < static_cast<int>(plan->ops.size()), "op index out of range");
  if (flops) *flops = plan->ops[i]->flops;
  if (bytes) *bytes = plan->ops[i]->bytes;
  return PCV_OK;
}

}  // extern "C"
