// Multi-GPU edge of the path: the ONE exchange of batch-sharded inference - an all-gather of each rank's
// [N/G, classes] logits - done by a single device-initiated kernel over NVLink peer memory instead of a host-launched
// NCCL collective.  Each rank owns one exchange buffer (cudaMalloc + CUDA IPC, mapped by every peer):
//
//   [ control: epoch | flags[world] ] [ slab 0: world x slot ] [ slab 1: world x slot ]
//
// Step e (e = local epoch + 1, identical on all ranks because every rank runs the same sequence of steps):
//   1. push   : copy the local shard into slot `rank` of slab e&1 of EVERY rank's buffer (16-byte peer stores; NVSwitch
//               gives each peer full bandwidth, so the 8 destinations proceed concurrently);
//   2. signal : system-scope fence, then flags[rank] := e in every peer's control block (st.release.sys);
//   3. wait   : spin (ld.acquire.sys) until the LOCAL flags of all ranks are >= e: every shard of step e has landed here;
//   4. drain  : copy slab e&1 (rank order == image order) into the caller's output tensor; epoch := e.
// The kernel runs as a small grid (64 CTAs): step 2 is done by the last CTA to finish its part of the push and the epoch
// is closed by the last CTA to finish draining (two counters in the control block).  No host synchronisation, no second stream, no launch per peer.  Slab parity makes the push of step e+1 safe: a peer can
// only be one step ahead (it needs this rank's step-e flag to finish step e), and slab (e+1)&1 was drained here in step
// e-1, before this rank signalled step e.
#include <algorithm>

#include "ptx.cuh"
#include "runtime.h"

namespace pcv {

constexpr int PEER_MAX_WORLD = 16;
constexpr size_t PEER_CTRL_BYTES = 4096;

struct PeerArgs {
  unsigned char* bufs[PEER_MAX_WORLD];   // every rank's exchange buffer, as mapped in THIS process
  int rank, world;
  size_t bytes;   // payload of one rank (multiple of 16)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_cg(const uint4* p) {   // L2 only: remote GPUs wrote this memory, L1 may be stale
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// Grid of PEER_CTAS CTAs.  Intra-grid ordering uses two counters in the control block ("last CTA to arrive acts"): no CTA
// ever waits for another CTA of its own grid, so the kernel needs no co-residency guarantee.
constexpr int PEER_CTAS = 64;
constexpr int PEER_THREADS = 512;

__global__ void __launch_bounds__(PEER_THREADS)
peer_allgather_kernel(const PeerArgs a, const uint4* __restrict__ local, uint4* __restrict__ out) {
  unsigned char* mine = a.bufs[a.rank];
  uint32_t* ctrl = reinterpret_cast<uint32_t*>(mine);   // [0] epoch, [1] pushed-CTA count, [2] drained-CTA count, [16 + r] flag of rank r
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(ctrl) + 1;   // stable: only rewritten after EVERY CTA has drained
  const size_t vecs = a.bytes >> 4;
  const size_t slab_off = PEER_CTRL_BYTES + static_cast<size_t>(e & 1) * a.world * a.bytes;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthr = static_cast<size_t>(gridDim.x) * blockDim.x;
  __shared__ int is_last;
  // 1. push: this rank's shard into slot `rank` of slab e&1 of every rank's buffer (the local vector is read once)
  for (size_t i = tid; i < vecs; i += nthr) {
    const uint4 v = local[i];
    for (int p = 0; p < a.world; ++p)
      reinterpret_cast<uint4*>(a.bufs[p] + slab_off + static_cast<size_t>(a.rank) * a.bytes)[i] = v;
  }
  // 2. signal: the CTA that completes the push tells every peer that this rank's shard of step e has been written
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ctrl + 1, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence_system();
    if (threadIdx.x < a.world) st_release_sys(reinterpret_cast<uint32_t*>(a.bufs[threadIdx.x]) + 16 + a.rank, e);
  }
  // 3. wait until every rank's shard of step e has landed HERE (a peer that never arrives is a protocol / launch error:
  //    trap after a few seconds instead of hanging the GPU)
  if (threadIdx.x < a.world) {
    const uint32_t* flag = ctrl + 16 + threadIdx.x;
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(flag) - e) < 0) {
      if (clock64() - t0 > 8000000000LL) __trap();
    }
  }
  __syncthreads();
  // 4. drain slab e&1 (rank order == image order) into the caller's tensor
  const uint4* slab = reinterpret_cast<const uint4*>(mine + slab_off);
  const size_t total = vecs * a.world;
  for (size_t i = tid; i < total; i += nthr) out[i] = ld_cg(slab + i);
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(ctrl + 2, 1u) == gridDim.x - 1) {   // last CTA out: close the step
    ctrl[1] = 0;
    ctrl[2] = 0;
    __threadfence();
    *reinterpret_cast<volatile uint32_t*>(ctrl) = e;
  }
}

struct PeerGatherOp : Op {
  PeerArgs a;
  const void* local;
  void* out;
  cudaError_t launch(cudaStream_t s) override {
    g_launches++;
    const int ctas = static_cast<int>(std::min<size_t>(PEER_CTAS, std::max<size_t>(1, (a.bytes >> 4) / PEER_THREADS)));
    peer_allgather_kernel<<<ctas, PEER_THREADS, 0, s>>>(a, reinterpret_cast<const uint4*>(local), reinterpret_cast<uint4*>(out));
    return cudaGetLastError();
  }
};

}  // namespace pcv

using namespace pcv;

extern "C" {

int pcv_peer_buffer_bytes(int world, size_t bytes_per_rank, size_t* total) {
  PCV_REQUIRE(total && world >= 1 && world <= PEER_MAX_WORLD && bytes_per_rank > 0 && bytes_per_rank % 16 == 0,
              "peer buffer: world in [1, %d] and a payload that is a multiple of 16 bytes", PEER_MAX_WORLD);
  *total = PEER_CTRL_BYTES + 2 * static_cast<size_t>(world) * bytes_per_rank;
  return PCV_OK;
}

int pcv_peer_buffer_alloc(size_t bytes, void** ptr, void* ipc_handle_out) {
  PCV_REQUIRE(ptr && ipc_handle_out && bytes >= PEER_CTRL_BYTES, "peer buffer: NULL argument or too small");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "pcv_peer_buffer_alloc documents a 64-byte handle");
  void* p = nullptr;
  PCV_CHECK_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(ipc_handle_out), p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(PCV_ERR_CUDA, "peer buffer setup failed: %s", cudaGetErrorString(e));
  }
  *ptr = p;
  return PCV_OK;
}

int pcv_peer_buffer_open(const void* ipc_handle, void** ptr) {
  PCV_REQUIRE(ipc_handle && ptr, "peer buffer: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof h);
  PCV_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return PCV_OK;
}

int pcv_peer_buffer_close(void* ptr) {
  PCV_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return PCV_OK;
}

int pcv_peer_buffer_free(void* ptr) {
  PCV_CHECK_CUDA(cudaFree(ptr));
  return PCV_OK;
}

int pcv_peer_allgather(pcv_plan* plan, const void* local, size_t bytes_per_rank, int rank, int world,
                       void* const* peer_bufs_host, void* out, pcv_stream stream) {
  PCV_REQUIRE(local && out && peer_bufs_host, "NULL tensor pointer");
  PCV_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
  PCV_REQUIRE(bytes_per_rank > 0 && bytes_per_rank % 16 == 0 && reinterpret_cast<uintptr_t>(local) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(out) % 16 == 0,
              "peer all-gather needs 16-byte aligned tensors and a payload that is a multiple of 16 bytes");
  auto op = std::make_unique<PeerGatherOp>();
  for (int p = 0; p < world; ++p) {
    PCV_REQUIRE(peer_bufs_host[p] != nullptr, "peer buffer %d is NULL", p);
    op->a.bufs[p] = reinterpret_cast<unsigned char*>(peer_bufs_host[p]);
  }
  op->a.rank = rank; op->a.world = world; op->a.bytes = bytes_per_rank;
  op->local = local; op->out = out;
  char nm[96];
  snprintf(nm, sizeof nm, "peer_allgather x%d %zu B/rank", world, bytes_per_rank);
  op->name = nm;
  op->bytes = 3.0 * world * static_cast<double>(bytes_per_rank);
  return submit(plan, op.release(), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
