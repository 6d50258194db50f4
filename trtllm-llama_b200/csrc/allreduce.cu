// One-shot tensor-parallel all-reduce over NVLink peer memory, fused with the residual add, for decode-size
// messages (M x hidden fp16 <= 64 KB).  Each rank's row-parallel projection writes its partial result straight into
// a peer-mapped buffer; this kernel signals the peers (flag store over NVLink), waits for their flags, then every
// rank reads all partials with 16-byte peer loads, sums them in RANK ORDER in fp32 (bit-identical on every rank,
// deterministic) and adds the residual:  out = residual + sum_r partial_r.
//
// Replaces, on the decode path, AllreducePlugin::enqueue -> ncclAllReduce (P/ncclPlugin/allreducePlugin.cpp:80-97)
// plus the TensorRT-native residual add that follows it (LQ/llama_model.py:96-118): 64 latency-bound collectives per
// token (SURVEY.md §7 "TP decode latency").  Large (prefill) messages keep the NCCL plugin.
//
// Protocol: two buffer/flag sets used alternately by consecutive calls (the engine's call sites alternate 0,1,0,1 and
// there is an even number per step, so a CUDA graph can bake the addresses).  A flag carries a monotonically
// increasing epoch kept in device memory (graph replay cannot pass a fresh host value).  Reuse safety: a rank passes
// the barrier of call k+1 only after every peer finished reading call k, so set k%2 is free again at call k+2.
#include "common.cuh"
#include "kernels.h"

namespace tb {

constexpr int kArMaxWorld = 8;
constexpr int kArThreads = 512;

struct ArSet {
  const uint4* data[kArMaxWorld];   // rank r's partial for this set (peer-mapped; [rank] is local)
  uint32_t* flags_of[kArMaxWorld];  // base of rank r's flag array for this set (we store into [my rank])
  uint32_t* my_flags;               // local flag array [world]
  uint32_t* epoch;                  // local epoch counter for this set
  uint32_t* arrive;                 // local block-arrival counter
};

struct ArParams {
  ArSet set;
  const __half* residual;
  __half* out;
  int64_t n16;   // 16-byte chunks
  int rank, world;
};

__device__ __forceinline__ uint4 ld_peer_v4(const uint4* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_flag_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_flag_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kArThreads) allreduce_oneshot_kernel(const ArParams p) {
  const ArSet& s = p.set;
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(s.epoch) + 1;
  // 1. this rank's partial was written by the previous kernel on this stream: publish it to every rank
  if (blockIdx.x == 0 && threadIdx.x < p.world) {
    __threadfence_system();
    st_flag_sys(s.flags_of[threadIdx.x] + p.rank, epoch);
  }
  // 2. wait until every rank (this one included) has published this epoch
  if (threadIdx.x < p.world) {
    const long long t0 = clock64();
    while (ld_flag_sys(s.my_flags + threadIdx.x) < epoch) {
      if (clock64() - t0 > 20000000000ll) {   // a peer died: trap instead of hanging the GPU
        printf("[trtllm_b200] all-reduce flag wait timed out (rank %d waiting for %d)\n", p.rank, threadIdx.x);
        __trap();
      }
    }
  }
  __syncthreads();
  // 3. rank-ordered fp32 sum of the partials + residual
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < p.n16; i += (int64_t) gridDim.x * blockDim.x) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    uint4 v[kArMaxWorld];
#pragma unroll
    for (int r = 0; r < kArMaxWorld; ++r)
      if (r < p.world) v[r] = ld_peer_v4(s.data[r] + i);
#pragma unroll
    for (int r = 0; r < kArMaxWorld; ++r) {
      if (r < p.world) {
        const __half2* h = reinterpret_cast<const __half2*>(&v[r]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
    if (p.residual) {
      // same rounding points as the unfused path: fp16(all-reduce result) then fp16(sum + residual)
      const uint4 rv = *reinterpret_cast<const uint4*>(p.residual + i * 8);
      const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = __half22float2(__floats2half2_rn(acc[2 * j], acc[2 * j + 1])), b = __half22float2(rh[j]);
        oh[j] = __floats2half2_rn(a.x + b.x, a.y + b.y);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
    }
    *reinterpret_cast<uint4*>(p.out + i * 8) = o;
  }
  // 4. the last block to finish advances the epoch for the next use of this set
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(s.arrive, 1u);
    if (prev == gridDim.x - 1) {
      *s.arrive = 0;
      *reinterpret_cast<volatile uint32_t*>(s.epoch) = epoch;
      __threadfence();
    }
  }
}

}  // namespace tb

using namespace tb;

struct tb_ar {
  int rank = 0, world = 1;
  size_t max_bytes = 0, extra_bytes = 0;
  uint8_t* local = nullptr;                 // [2 sets][max_bytes] data, then [2][8] flags, [2] epochs, [2] arrive, then `extra`
  uint8_t* peer[kArMaxWorld] = {nullptr};   // mapped bases (peer[rank] == local)
  bool opened = false;
};

static size_t ar_total_bytes(size_t max_bytes, size_t extra = 0) { return 2 * max_bytes + 1024 + extra; }

extern "C" {

int tb_ar_create(tb_ar** out, int rank, int world, size_t max_bytes) {
  return tb_ar_create_ex(out, rank, world, max_bytes, 0);
}
int tb_ar_create_ex(tb_ar** out, int rank, int world, size_t max_bytes, size_t extra_bytes) {
  if (!out || world < 2 || world > kArMaxWorld || rank < 0 || rank >= world) return -1;
  auto* a = new tb_ar();
  a->rank = rank; a->world = world; a->max_bytes = (max_bytes + 255) & ~(size_t) 255;
  a->extra_bytes = (extra_bytes + 255) & ~(size_t) 255;
  TB_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&a->local), ar_total_bytes(a->max_bytes, a->extra_bytes)));
  TB_CHECK_CUDA(cudaMemset(a->local, 0, ar_total_bytes(a->max_bytes, a->extra_bytes)));
  TB_CHECK_CUDA(cudaDeviceSynchronize());
  a->peer[rank] = a->local;
  *out = a;
  return 0;
}
void tb_ar_destroy(tb_ar* a) {
  if (!a) return;
  for (int r = 0; r < a->world; ++r)
    if (r != a->rank && a->peer[r]) cudaIpcCloseMemHandle(a->peer[r]);
  if (a->local) cudaFree(a->local);
  delete a;
}
/* 64-byte cudaIpcMemHandle_t of this rank's buffer, to be exchanged by the host (torch.distributed / a file). */
int tb_ar_ipc_handle(tb_ar* a, void* out64) {
  cudaIpcMemHandle_t h;
  TB_CHECK_CUDA(cudaIpcGetMemHandle(&h, a->local));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out64, &h, 64);
  return 0;
}
int tb_ar_open_peers(tb_ar* a, const void* handles /* world x 64 bytes, rank order */) {
  for (int r = 0; r < a->world; ++r) {
    if (r == a->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const uint8_t*>(handles) + 64 * r, 64);
    void* p = nullptr;
    TB_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    a->peer[r] = static_cast<uint8_t*>(p);
  }
  a->opened = true;
  return 0;
}
/* rank r's `extra` area as mapped into this process (r == own rank: the local one); NULL before tb_ar_open_peers */
void* tb_ar_extra(tb_ar* a, int r) {
  if (!a || r < 0 || r >= a->world || !a->peer[r] || a->extra_bytes == 0) return nullptr;
  if (r != a->rank && !a->opened) return nullptr;
  return a->peer[r] + 2 * a->max_bytes + 1024;
}
size_t tb_ar_extra_bytes(tb_ar* a) { return a ? a->extra_bytes : 0; }
/* local buffer a row-parallel projection should write its partial result to, for call-site parity `set` */
void* tb_ar_buffer(tb_ar* a, int set) { return a->local + (size_t) (set & 1) * a->max_bytes; }

int tb_ar_allreduce(tb_ar* a, int set, void* out, const void* residual, int64_t n_half, cudaStream_t stream) {
  if (!a || !a->opened) return -1;
  if (n_half % 8 != 0 || (size_t) n_half * 2 > a->max_bytes) return -2;
  set &= 1;
  ArParams p{};
  uint8_t* ctl_off = nullptr;
  for (int r = 0; r < a->world; ++r) {
    p.set.data[r] = reinterpret_cast<const uint4*>(a->peer[r] + (size_t) set * a->max_bytes);
    uint8_t* ctl = a->peer[r] + 2 * a->max_bytes;
    p.set.flags_of[r] = reinterpret_cast<uint32_t*>(ctl + set * 64);
    if (r == a->rank) ctl_off = ctl;
  }
  p.set.my_flags = reinterpret_cast<uint32_t*>(ctl_off + set * 64);
  p.set.epoch = reinterpret_cast<uint32_t*>(ctl_off + 128 + set * 16);
  p.set.arrive = reinterpret_cast<uint32_t*>(ctl_off + 256 + set * 16);
  p.residual = static_cast<const __half*>(residual);
  p.out = static_cast<__half*>(out);
  p.n16 = n_half / 8;
  p.rank = a->rank; p.world = a->world;
  int grid = (int) ((p.n16 + kArThreads - 1) / kArThreads);
  if (grid > 16) grid = 16;
  if (grid < 1) grid = 1;
  allreduce_oneshot_kernel<<<grid, kArThreads, 0, stream>>>(p);
  return (int) cudaGetLastError();
}
}
