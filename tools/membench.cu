// Micro-benchmark (development aid, not part of the product): what read bandwidth do the candidate weight-streaming
// mechanisms reach on this B200?  (a) per-thread 16-byte loads, (b) cp.async.bulk rings with idle consumers.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../trtllm-llama_b200/csrc/common.cuh"
using namespace tb;

template <int U>
__global__ void ldg_stream(const uint4* __restrict__ p, size_t n16, unsigned* out) {
  unsigned acc = 0;
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (; i + (U - 1) * stride < n16; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldg_nc_v4(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345) *out = acc;
}

// warp-per-row pattern: warp w reads row r = w + k*total_warps, each row `row16` uint4 long
template <int U>
__global__ void ldg_rows(const uint4* __restrict__ p, int rows, int row16, unsigned* out) {
  unsigned acc = 0;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, tw = (gridDim.x * blockDim.x) >> 5;
  for (int r = gw; r < rows; r += tw) {
    const uint4* row = p + (size_t) r * row16;
    for (int c = lane; c < row16; c += 32 * U) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = (c + 32 * u < row16) ? ldg_nc_v4(row + c + 32 * u) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
  }
  if (acc == 0x12345) *out = acc;
}

// MMA-fragment pattern: a warp owns 16 rows; per k-step lane (g = lane/4, t = lane%4) loads 16 B from row g and row g+8
// at byte offset kstep*64 + t*16  (each instruction touches 8 rows x 64 contiguous bytes)
template <int U>
__global__ void ldg_frag(const uint4* __restrict__ p, int rows, int row16, unsigned* out) {
  unsigned acc = 0;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, tw = (gridDim.x * blockDim.x) >> 5;
  const int ksteps = row16 / 4;
  for (int r0 = gw * 16; r0 < rows; r0 += tw * 16) {
    const uint4* lo = p + (size_t) (r0 + g) * row16 + t;
    const uint4* hi = p + (size_t) (r0 + g + 8) * row16 + t;
    for (int ks = 0; ks < ksteps; ks += U) {
      uint4 a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { a[u] = ldg_nc_v4(lo + (ks + u) * 4); b[u] = ldg_nc_v4(hi + (ks + u) * 4); }
#pragma unroll
      for (int u = 0; u < U; ++u) acc += a[u].x ^ a[u].y ^ a[u].z ^ a[u].w ^ b[u].x ^ b[u].y ^ b[u].z ^ b[u].w;
    }
  }
  if (acc == 0x12345) *out = acc;
}

__device__ __forceinline__ void bulk_ld(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// each CTA streams a contiguous range in `stage_bytes` pieces through an `nst`-deep ring; consumers only touch one word
__global__ void bulk_stream(const uint8_t* __restrict__ p, size_t total, int stage_bytes, int nst, int ncopies, unsigned* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t) nst * stage_bytes);
  uint64_t* empty = full + nst;
  const size_t per = (total / gridDim.x) / stage_bytes * stage_bytes;
  const uint8_t* base = p + per * blockIdx.x;
  const int nstage = (int) (per / stage_bytes);
  const int nwarps = blockDim.x / 32 - 1;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nwarps); }
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == nwarps) {
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      int st = 0, ph = 0;
      for (int i = 0; i < nstage; ++i) {
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], stage_bytes);
        const int cb = stage_bytes / ncopies;
        for (int c = 0; c < ncopies; ++c)
          bulk_ld(smem + (size_t) st * stage_bytes + c * cb, base + (size_t) i * stage_bytes + c * cb, cb, &full[st], pol);
        if (++st == nst) { st = 0; ph ^= 1; }
      }
    }
    return;
  }
  unsigned acc = 0;
  int st = 0, ph = 0;
  for (int i = 0; i < nstage; ++i) {
    mbar_wait(&full[st], ph);
    acc += *reinterpret_cast<const unsigned*>(smem + (size_t) st * stage_bytes + threadIdx.x * 4);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
    if (++st == nst) { st = 0; ph ^= 1; }
  }
  if (acc == 0x12345) *out = acc;
}

template <class F> float time_it(F f, int reps = 10) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const size_t total = (size_t) 2048 << 20;   // 2 GiB buffer, kernels read `sz` bytes at rotating offsets (> L2)
  uint8_t* buf; unsigned* out;
  cudaMalloc(&buf, total); cudaMalloc(&out, 4);
  cudaMemset(buf, 1, total);
  size_t szs[3] = {(size_t) 32 << 20, (size_t) 96 << 20, (size_t) 176 << 20};
  for (size_t sz : szs) {
    printf("== %zu MiB per launch\n", sz >> 20);
    size_t off = 0;
    auto next = [&]() { off = (off + sz + (64 << 20)) % (total - sz); off &= ~(size_t) 4095; return buf + off; };
    for (int bps : {2, 4, 6, 8}) {
      float ms = time_it([&] { ldg_stream<8><<<148 * bps, 256>>>((const uint4*) next(), sz / 16, out); });
      printf("ldg_stream U=8 blocks/SM=%d : %.2f us  %.0f GB/s\n", bps, ms * 1e3, sz / ms / 1e6);
    }
    { float ms = time_it([&] { ldg_stream<4><<<148 * 8, 256>>>((const uint4*) next(), sz / 16, out); });
      printf("ldg_stream U=4 blocks/SM=8 : %.2f us  %.0f GB/s\n", ms * 1e3, sz / ms / 1e6); }
    { float ms = time_it([&] { ldg_stream<16><<<148 * 4, 256>>>((const uint4*) next(), sz / 16, out); });
      printf("ldg_stream U=16 blocks/SM=4 : %.2f us  %.0f GB/s\n", ms * 1e3, sz / ms / 1e6); }
    for (int bps : {4, 6, 8}) {
      float ms = time_it([&] { ldg_rows<8><<<148 * bps, 256>>>((const uint4*) next(), (int) (sz / 8192), 512, out); });
      printf("ldg_rows (8KB rows) U=8 blocks/SM=%d : %.2f us  %.0f GB/s\n", bps, ms * 1e3, sz / ms / 1e6);
    }
    for (int bps : {2, 3, 4}) {
      float ms = time_it([&] { ldg_frag<4><<<148 * bps, 256>>>((const uint4*) next(), (int) (sz / 8192), 512, out); });
      printf("ldg_frag (16 rows x 64B) U=4 blocks/SM=%d : %.2f us  %.0f GB/s\n", bps, ms * 1e3, sz / ms / 1e6);
      ms = time_it([&] { ldg_frag<8><<<148 * bps, 256>>>((const uint4*) next(), (int) (sz / 8192), 512, out); });
      printf("ldg_frag (16 rows x 64B) U=8 blocks/SM=%d : %.2f us  %.0f GB/s\n", bps, ms * 1e3, sz / ms / 1e6);
    }
    cudaFuncSetAttribute(bulk_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    struct C { int grid_per_sm, stage, nst, ncopies; };
    C cs[] = {{2, 16384, 4, 1}, {2, 16384, 5, 1}, {2, 8192, 8, 1}, {1, 32768, 6, 1}, {1, 16384, 12, 1}, {2, 16384, 4, 16}, {4, 8192, 4, 1}, {2, 32768, 3, 1}, {1, 65536, 3, 1}};
    for (C c : cs) {
      size_t smem = (size_t) c.nst * c.stage + 2 * c.nst * 8;
      float ms = time_it([&] { bulk_stream<<<148 * c.grid_per_sm, 288, smem>>>(next(), sz, c.stage, c.nst, c.ncopies, out); });
      cudaError_t e = cudaGetLastError();
      printf("bulk ring ctas/SM=%d stage=%dKB x%d copies/stage=%d : %.2f us  %.0f GB/s %s\n", c.grid_per_sm, c.stage >> 10, c.nst, c.ncopies,
             ms * 1e3, sz / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
