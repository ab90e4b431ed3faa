// sm_100a PTX wrappers used by the pytorchcv_b200 kernels: mbarrier, TMA (tiled + im2col),
// tcgen05 (alloc / mma / commit / ld), proxy fences.  Hand-written inline PTX; no CUTLASS.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/pcv_b200.h"   // pcv_act codes for fast_act_n

namespace pcv {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor drains; it must not touch the predecessor's memory before pdl_wait().  No-ops otherwise.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// Optional suspend-time hint of mbarrier.try_wait (a translation unit defines PCV_MBAR_SUSPEND_NS before including this file):
// without one a waiting warp re-polls every few tens of ns - in the CUDA-core-bound fused dw -> pw kernels the poll loops of
// the idle role warps were 32 % of all issued instructions (ncu source page) and compete with the stencil warps for issue
// slots; with it the warp sleeps in hardware until the phase completes (or the hint expires).  The latency-bound GEMM
// kernels keep the plain form (ResNet-18 fp32 bs8: -0.5 % with the hint).
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
#ifdef PCV_MBAR_SUSPEND_NS
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
#endif
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
#ifdef PCV_MBAR_SUSPEND_NS
      : "r"(smem_u32(bar)), "r"(parity), "r"(PCV_MBAR_SUSPEND_NS)
#else
      : "r"(smem_u32(bar)), "r"(parity)
#endif
      : "memory");
  return ok != 0;
}
// Spin on a phase parity.  try_wait suspends in hardware, so this is not a hot poll.
// A wait that outlives ~2 s of SM clocks is a protocol bug (lost TMA, wrong tx count): trap so the launch fails with
// an error instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// ----------------------------------------------------------------------------------------------
// proxy fences / named barriers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// im2col-mode load of a (pixelsPerColumn x channelsPerPixel) tile whose first base pixel is (w,h,n) and whose
// filter-tap displacement is (off_w, off_h).
__device__ __forceinline__ void tma_load_im2col_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n),
      "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the 64-bit smem descriptors passed as (lo, hi) words: the issuing thread keeps `hi` constant and only
// adds to `lo` (start address >> 4) per stage / K step, so one MMA costs ~3 scalar instructions to set up.
constexpr uint32_t SMEM_DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO=1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
}
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(SMEM_DESC_HI_SW128)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(SMEM_DESC_HI_SW128)
      : "memory");
}
// cta_group::2 MMA with an explicit high descriptor word for A (experiments with the matrix base-offset field)
__device__ __forceinline__ void umma2_bf16_lohi_a(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %6};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(SMEM_DESC_HI_SW128), "r"(a_hi)
      : "memory");
}
// cta_group::2 MMA with one explicit high descriptor word shared by A and B (layouts other than SWIZZLE_128B)
constexpr uint32_t SMEM_DESC_HI_SW32 = (256u >> 4) | (1u << 14) | (6u << 29);   // SBO=256 B, version 1, SWIZZLE_32B
__device__ __forceinline__ void umma2_bf16_lohi_h(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                  uint32_t accumulate, uint32_t hi) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(hi)
      : "memory");
}
// mbarrier arrive once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// tcgen05.wait::ld that also "touches" the destination registers of an earlier tcgen05.ld, so that neither the
// compiler nor ptxas can schedule a use of them above the wait when other loads are issued in between
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.wait::ld.sync.aligned;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
        "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
        "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
        "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
      :
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane_base + i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// Shared-memory operand descriptor for a K-major tile whose rows are `row_bytes` (32/64/128) wide and were written
// by TMA with the matching swizzle (SWIZZLE_32B/64B/128B).  Bit layout (sm_100 "SmemDescriptor"): [0,14) start>>4,
// [16,30) leading byte offset>>4 (1 for swizzled K-major), [32,46) stride byte offset>>4 (8 rows), [46,48) version=1,
// [61,64) layout type (2=128B, 4=64B, 6=32B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t row_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((8u * row_bytes) >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= layout << 61;
  return d;
}

// Instruction descriptor for kind::f16 (K-major A and B, fp32 accumulate, M x N tile).  Bits [7,10) / [10,13) are the
// A / B element formats: 1 = bf16, 0 = f16 - the same instruction and rate serve both 16-bit storage tiers.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// max(x, 0) and the bf16 rounding in one instruction (F2FP.RELU): the ReLU epilogue's clamp for free
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// fp16-storage tier twins (same roles, IEEE half: 10-bit mantissa, max 65504 - overflow rounds to inf like torch's .half())
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack_relu_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float f16lo(uint32_t v) {
  return __half2float(__ushort_as_half(static_cast<unsigned short>(v & 0xFFFFu)));
}
__device__ __forceinline__ float f16hi(uint32_t v) {
  return __half2float(__ushort_as_half(static_cast<unsigned short>(v >> 16)));
}
__device__ __forceinline__ uint32_t hmax2_f16(uint32_t a, uint32_t b) {
  const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t hmin2_f16(uint32_t a, uint32_t b) {
  const __half2 r = __hmin2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
// packed fp32x2 add (Blackwell FADD2); each lane rounds like a scalar add
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) {
  uint64_t aa = (static_cast<uint64_t>(__float_as_uint(a.y)) << 32) | __float_as_uint(a.x);
  const uint64_t bb = (static_cast<uint64_t>(__float_as_uint(b.y)) << 32) | __float_as_uint(b.x);
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(aa) : "l"(bb));
  return make_float2(__uint_as_float(static_cast<uint32_t>(aa)), __uint_as_float(static_cast<uint32_t>(aa >> 32)));
}
__device__ __forceinline__ uint32_t hmax2_bf16(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t hmin2_bf16(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hmin2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
// Explicit shared-window accesses with 32-bit addresses.  Pointers derived from the aligned dynamic-smem base lose their
// address space (the compiler emits generic LD.E/ST.E with 64-bit address arithmetic); these keep LDS/STS.
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// bf16-tier versions of the smooth activations (activ.py:16-47): one MUFU.TANH instead of ex2 + a full-precision
// division.  sigmoid(x) = 0.5 + 0.5 tanh(x/2); tanh.approx.f32 has ~2^-11 relative error, below the bf16 the result is
// rounded to.  The fp32 tier (conv_simt.cu, misc.cu) keeps expf.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
__device__ __forceinline__ float fast_swish(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}
__device__ __forceinline__ float fast_hsigmoid(float x) { return fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f); }
__device__ __forceinline__ float fast_hswish(float x) { return x * fast_hsigmoid(x); }
// One activation over N register values with the switch OUTSIDE the element loop (an inlined per-element switch is what
// once blew the epilogue up to 100 KB of SASS).  Codes are pcv_act (include/pcv_b200.h): 3 sigmoid, 4 swish, 5 h-swish,
// 6 h-sigmoid; the clamp family (0..2) is normally handled by the callers' packed paths.
template <int N>
__device__ __forceinline__ void fast_act_n(float (&v)[N], int act, float leaky_slope = 0.f) {
  if (act == PCV_ACT_SWISH) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fast_swish(v[i]);
  } else if (act == PCV_ACT_HSWISH) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fast_hswish(v[i]);
  } else if (act == PCV_ACT_SIGMOID) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fast_sigmoid(v[i]);
  } else if (act == PCV_ACT_HSIGMOID) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fast_hsigmoid(v[i]);
  } else if (act == PCV_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (act == PCV_ACT_RELU6) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fminf(fmaxf(v[i], 0.f), 6.f);
  } else if (act == PCV_ACT_LEAKY_RELU) {   // nn.LeakyReLU (activ.py:101-120)
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = v[i] >= 0.f ? v[i] : v[i] * leaky_slope;
  }
}
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

}  // namespace pcv

// ----------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: clusters of two CTAs on one TPC sharing one 256-row MMA
// ----------------------------------------------------------------------------------------------
namespace pcv {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion may be signalled on the PEER CTA's mbarrier (`mbar_cluster_addr` from mapa_u32)
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* m, uint32_t mbar_cluster_addr, void* dst, int c0,
                                             int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(const CUtensorMap* m, uint32_t mbar_cluster_addr, void* dst, int c,
                                                    int w, int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
      "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// One 256 x N x 16 MMA across the CTA pair: A rows 0-127 / B columns 0..N/2-1 from CTA 0's smem, the rest from CTA 1.
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the pair's MMAs so far have completed) on the barrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void umma2_commit(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

}  // namespace pcv

// ----------------------------------------------------------------------------------------------
// 16-bit storage tier of this translation unit.  The tensor-core / TMA-window kernels are compiled twice, once per tier
// (Makefile: -DPCV_TIER=bf -DPCV_HALF=0 and -DPCV_TIER=hf -DPCV_HALF=1), into namespaces pcv::bf and pcv::hf; inside them
// `e16` is the element type and the e16 helpers below resolve to the bf16 or the f16 instruction.
// ----------------------------------------------------------------------------------------------
#ifdef PCV_TIER
namespace pcv {
namespace PCV_TIER {
#if PCV_HALF
using e16 = __half;
using e16x2 = __half2;
constexpr CUtensorMapDataType TMAP_E16 = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
constexpr uint32_t E16_NEG_INF2 = 0xFC00FC00u;
constexpr int TIER_DTYPE = PCV_F16;
#define PCV_TIER_NAME "f16"
__host__ __device__ constexpr uint32_t make_idesc_e16(int M, int N) { return make_idesc_f16(M, N); }
__device__ __forceinline__ uint32_t pack_e16x2(float lo, float hi) { return pack_f16x2(lo, hi); }
__device__ __forceinline__ uint32_t pack_relu_e16x2(float lo, float hi) { return pack_relu_f16x2(lo, hi); }
__device__ __forceinline__ float e16lo(uint32_t v) { return f16lo(v); }
__device__ __forceinline__ float e16hi(uint32_t v) { return f16hi(v); }
__device__ __forceinline__ uint32_t hmax2_e16(uint32_t a, uint32_t b) { return hmax2_f16(a, b); }
__device__ __forceinline__ uint32_t hmin2_e16(uint32_t a, uint32_t b) { return hmin2_f16(a, b); }
__device__ __forceinline__ float e16_to_float(e16 v) { return __half2float(v); }
__device__ __forceinline__ e16 float_to_e16(float v) { return __float2half_rn(v); }
#else
using e16 = __nv_bfloat16;
using e16x2 = __nv_bfloat162;
constexpr CUtensorMapDataType TMAP_E16 = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
constexpr uint32_t E16_NEG_INF2 = 0xFF80FF80u;
constexpr int TIER_DTYPE = PCV_BF16;
#define PCV_TIER_NAME "bf16"
__host__ __device__ constexpr uint32_t make_idesc_e16(int M, int N) { return make_idesc_bf16(M, N); }
__device__ __forceinline__ uint32_t pack_e16x2(float lo, float hi) { return pack_bf16x2(lo, hi); }
__device__ __forceinline__ uint32_t pack_relu_e16x2(float lo, float hi) { return pack_relu_bf16x2(lo, hi); }
__device__ __forceinline__ float e16lo(uint32_t v) { return bf16lo(v); }
__device__ __forceinline__ float e16hi(uint32_t v) { return bf16hi(v); }
__device__ __forceinline__ uint32_t hmax2_e16(uint32_t a, uint32_t b) { return hmax2_bf16(a, b); }
__device__ __forceinline__ uint32_t hmin2_e16(uint32_t a, uint32_t b) { return hmin2_bf16(a, b); }
__device__ __forceinline__ float e16_to_float(e16 v) { return __bfloat162float(v); }
__device__ __forceinline__ e16 float_to_e16(float v) { return __float2bfloat16(v); }
#endif
__device__ __forceinline__ uint32_t hmax3_e16(uint32_t a, uint32_t b, uint32_t c) {   // one VHMNMX
  return hmax2_e16(hmax2_e16(a, b), c);
}
// The clamp-family activations (none / ReLU / ReLU6: lo in {-inf, 0}, hi in {+inf, 6}) on 16 packed pairs: the lower
// bound rides on the conversion (F2FP.RELU), the upper bound is a packed e16 min AFTER rounding (6.0 is exact in e16
// and rounding is monotonic, so min-then-round == round-then-min).  16 or 32 instructions instead of 64 FMNMX + 16 F2FP.
__device__ __forceinline__ void clamp_pack32(const float (&v)[32], uint32_t (&o)[16], bool relu, bool capped, uint32_t cap2) {
  if (relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = pack_relu_e16x2(v[2 * i], v[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = pack_e16x2(v[2 * i], v[2 * i + 1]);
  }
  if (capped) {
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = hmin2_e16(o[i], cap2);
  }
}
}  // namespace PCV_TIER
}  // namespace pcv
#endif  // PCV_TIER
