// One persistent kernel per generated token: the whole LLaMA decode step (every layer's QKV projection, fused masked
// attention, dense, gate/up + SwiGLU, down projection, then ln_f + lm_head + greedy argmax + step bookkeeping) in ONE
// cooperative launch of one CTA per SM, for 1..4 token rows.
//
// Why: the kernel-per-operator step (164 kernels under one CUDA graph) loses
// 25-75 % of a step to launch / drain / prologue gaps between dependent kernels, because every kernel must finish before
// the next may read its activations and the HBM pipe empties at each boundary.  Weights do not depend on activations:
// here a producer warp per CTA streams this CTA's share of EVERY projection of EVERY layer, in execution order, through
// a ring of 4 KB shared-memory stages with cp.async.bulk (one elected thread, completion on mbarriers), running ahead
// across phase boundaries for as far as the ring is deep (~170 KB per SM = ~25 MB in flight chip-wide).  The 16
// consumer warps wait at a grid-wide barrier between phases (an L2 counter: ~1 us instead of a kernel boundary), stage
// the phase's activations (RMSNorm / per-token int8 quantisation fused, as in gemv.cu), and then find the first ~170 KB
// of the phase's weights already resident in shared memory.
//
// Work split: a projection with n_out output channels gives CTA c the contiguous channel range
// [n_out*c/G, n_out*(c+1)/G) — one contiguous byte range of the [N,K] weight matrix (two for gate|up), balanced to
// within one row (>= 98.8 % for LLaMA-7B on 148 SMs; the warp-per-row round-robin of gemv.cu is 86 % balanced at N =
// 4096).  A stage is one <= 4 KB segment of one weight row, consumed by one warp: exact fp16 x fp16 products in fp32
// (int32 dp4a for W8A8), warp-reduced to one partial per (row, segment) in shared memory; after the phase's last stage
// the CTA sums the segments in fixed order and applies the fused epilogue (per-channel / per-token scales, SwiGLU,
// residual add) — deterministic, no atomics.
//
// Attention phase: one (sequence, head) per CTA round-robin, 512 threads, RoPE + in-place KV append + QK^T.softmax.V
// as mmha.cu (FMA loops, no split: contexts this path serves are <= a few thousand positions of one head).
//
// Replaces, for M <= 4 rows: the per-step plugin schedule of T/tensorrt_llm/runtime/generation.py:852-963 (one
// enqueue per plugin per layer) — GPTAttention generation phase (P/gptAttentionCommon/gptAttentionCommon.cpp:649-780),
// Gemm / WeightOnlyQuantMatmul / SmoothQuantGemm at decode shapes, RmsnormQuantization, QuantizePerToken, the
// TensorRT-native glue and the greedy DynamicDecodeOp.  The plugin path stays (tbrt decode_mode 0) for the
// boundary tests and for shapes this kernel does not take.
// Algorithmic bytes per launch = weights of the model + lm_head + K/V rows read (13.26 GB for cfg2).
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "kernels.h"

namespace tb {
namespace ds {

constexpr int kCW = 16;               // consumer warps
constexpr int kCT = kCW * 32;         // consumer threads (named barrier 1)
constexpr int kThreads = kCT + 32;    // + one producer warp
constexpr int kSeg = 4096;            // bytes per ring stage
constexpr int kDh = 128;
constexpr int kMaxStages = 60;
constexpr int kPartFloats = 512;      // x MB: per-(row, segment) partial sums of one phase of one CTA
constexpr int kLgFloats = 256;        // x MB: this CTA's logits (argmax candidates)
constexpr int kMaxRows = 4;
constexpr int kTraceSlots = 2048;

enum Kind { kF16 = 0, kW8 = 1, kW4 = 2, kA8W8 = 3 };
enum XFormat { kXHalf = 0, kXFloat = 1, kXInt8 = 2 };

struct Layer {
  const uint8_t *w_qkv, *w_dense, *w_fc, *w_proj;
  const void *s_qkv, *s_dense, *s_fc, *s_proj;   // per-channel scales: fp16 (weight-only) / fp32 (SmoothQuant)
  const __half *ln_in, *ln_post;
  uint8_t* kv;                                   // [B, 2, Hl, S_max, Dh]
  const float *kv_oq, *kv_qo;
};

struct Params {
  const Layer* layers;
  int n_layers, kind, B, hidden, hid_l, inter_l, Hl, vocab_l, vocab, S_max, int8_kv, out_stride;
  float eps, inv_sqrt_dh;
  const __half *emb, *ln_f;
  const uint8_t* lm_head;
  __half *hA, *hB, *qkv, *att, *act;
  float* logits;
  float* cand_v;
  int* cand_i;
  int *ids, *seq_lens, *step_pos, *out_ids, *next;
  const int *in_lens, *max_in;
  unsigned long long* bar;
  int stages, prefetch;          // ring slots; L2 prefetch distance in stages (0: off)
  int interleave;                // output channels dealt round-robin over the CTAs (1) or in contiguous blocks (0)
  // tensor parallel: rank r's peer-mapped scratch (TpLayout offsets); tp == 1: unused
  int tp, rank;
  uint8_t* peer[8];
  int debug;                     // diagnostics (TB_DS_DEBUG): bit 0 = grid barriers do not wait (timing of the pure weight stream; results are garbage)
  unsigned long long* trace;     // optional [gridDim.x][kTraceSlots]: globaltimer stamps of the consumer pipeline (diagnostics)
  uint32_t xs_off, ring_off;
};

// smem layout (dynamic): [0,1024) mbarriers | [1024,1536) reduction scratch + per-token scales | part | lgs | xs | ring
constexpr uint32_t kOffRed = 1024, kOffPart = 1536;
__host__ __device__ constexpr uint32_t off_lg(int MB) { return kOffPart + (uint32_t) kPartFloats * MB * 4; }
__host__ __device__ constexpr uint32_t off_xs(int MB) { return off_lg(MB) + (uint32_t) kLgFloats * MB * 4; }

__host__ __device__ constexpr int epc_of(int kind) { return kind == kF16 ? 8 : (kind == kW4 ? 32 : 16); }
// fp16 activations (int8 for W8A8).  An fp32 copy for one row saves the consumers a conversion per element but costs
// 22 KB of shared memory = the difference between 32 and 48 ring slots, which matters more.
template <int KIND, int MB> struct XF { static constexpr int v = KIND == kA8W8 ? kXInt8 : kXHalf; };
__host__ __device__ constexpr int xbytes_of(int xf) { return xf == kXFloat ? 4 : (xf == kXHalf ? 2 : 1); }

// Staged activations are stored chunk-interleaved: a "chunk" is the E activations that meet one 16-byte weight chunk
// (E = 8 / 16 / 32 elements = E * XB bytes = NV 16-byte vectors).  Lane l of a warp works on chunk c = l + 32 u, so vector
// j of 32 consecutive chunks is stored contiguously: vec(c, j) = ((c / 32) * NV + j) * 32 + (c % 32).  A warp's LDS.128
// for one j then covers 512 contiguous bytes (conflict-free); the natural layout put consecutive lanes E * XB bytes
// apart (2-, 4- and 8-way bank conflicts for fp16 / int8 / int4 weights against fp32 activations).
__host__ __device__ constexpr int xrow_bytes(int K, int E, int XB) { return ((K / E + 31) / 32) * 32 * E * XB; }
__device__ __forceinline__ uint32_t xvec_off(int c, int j, int NV) {
  return (uint32_t) (((c >> 5) * NV + j) * 32 + (c & 31)) * 16u;
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define DS_STAMP()                                                                                   \
  do {                                                                                               \
    if (p.trace && ctid == 0 && tr_n < kTraceSlots) p.trace[(size_t) blockIdx.x * kTraceSlots + tr_n++] = gtime(); \
  } while (0)

__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ float silu_f(float v) { return v / (1.f + __expf(-v)); }
__device__ __forceinline__ uint4 lds128(const void* p) { return *reinterpret_cast<const uint4*>(p); }
// fp16 activation written by another CTA earlier in this launch: read through L2, never a stale L1 line
__device__ __forceinline__ float ldcg_h(const __half* p) {
  return __half2float(__ushort_as_half(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}

__device__ __forceinline__ void bulk_load_1d_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                                  uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// ---- tensor-parallel scratch, identical on every rank (offsets into the peer-mapped area) --------------------------------
//   flags   u32[8]                              flags[r] = number of cross-GPU barriers rank r has completed (monotonic)
//   xepoch  u32                                 cross-GPU barriers completed before this launch (local bookkeeping)
//   partial [2 sets][world][kMaxRows][hidden]   4-byte granules {fp16 partial sum, 16-bit exchange tag}, pushed by every rank
//   cand    [world][grid][kMaxRows] {f32, i32}  arg-max candidates of every rank's vocabulary shard
//   logits  [kMaxRows][vocab] f32               all-gathered logits
struct TpLayout {
  size_t flags, xepoch, partial, cand, logits, total;
};
__host__ __device__ inline TpLayout tp_layout(int hidden, int vocab, int grid) {
  TpLayout l;
  l.flags = 0;
  l.xepoch = 128;
  l.partial = 256;
  l.cand = l.partial + (size_t) 2 * 8 * kMaxRows * hidden * 4;
  l.logits = (l.cand + (size_t) 8 * grid * kMaxRows * 8 + 255) & ~(size_t) 255;
  l.total = l.logits + (size_t) kMaxRows * vocab * 4;
  return l;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// Cross-GPU barrier (after a phase whose outputs were pushed to the peers): every writer thread fences its remote stores
// system-wide, the CTA arrives on the local counter, and the LAST CTA of this rank publishes the rank's barrier count to
// every peer (one 4-byte release store over NVLink each).  Waiters poll their LOCAL flag array until every rank has
// published `xe`.  The local counter keeps the same accounting as grid_arrive (one arrival per CTA per barrier).
__device__ __forceinline__ void xgrid_arrive(const Params& p, unsigned long long* bar, unsigned long long target_after,
                                             uint32_t xe, int ctid) {
  cbar();
  if (ctid == 0) {
    __threadfence_system();      // the CTA's remote stores (ordered before this thread by the CTA barrier), system-wide
    unsigned long long old;
    asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], 1;" : "=l"(old) : "l"(bar) : "memory");
    if (old + 1 == target_after) {
      __threadfence_system();
      const TpLayout l = tp_layout(p.hidden, p.vocab, gridDim.x);
      for (int r = 0; r < p.tp; ++r) st_release_sys_u32(reinterpret_cast<uint32_t*>(p.peer[r] + l.flags) + p.rank, xe);
    }
  }
}
__device__ __forceinline__ void xgrid_wait(const Params& p, uint32_t xe, int ctid) {
  if (ctid < p.tp && !(p.debug & 1)) {
    const TpLayout l = tp_layout(p.hidden, p.vocab, gridDim.x);
    const uint32_t* f = reinterpret_cast<const uint32_t*>(p.peer[p.rank] + l.flags) + ctid;
    const long long t0 = clock64();
    for (;;) {
      uint32_t v;
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int32_t) (v - xe) >= 0) {
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        break;
      }
      if (clock64() - t0 > 40000000000ll) {      // a peer died or never launched: trap instead of hanging the GPU
        printf("[trtllm_b200] decode_step cross-GPU barrier timed out (rank %d waiting for %d, block %d)\n", p.rank, ctid,
               blockIdx.x);
        __trap();
      }
    }
  }
  cbar();
}

// ---- grid-wide barrier: a monotonically increasing arrival counter in L2 (reset to 0 by CTA 0 at the end of the launch)
// The CTA's writes reach thread 0 through the CTA barrier; its gpu-scope fence + release then publishes them (the
// cooperative-groups grid.sync pattern): 511 threads do not pay a membar each.
__device__ __forceinline__ void grid_arrive(unsigned long long* bar, int ctid) {
  cbar();
  if (ctid == 0) asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(bar) : "memory");
}
__device__ __forceinline__ void grid_wait(const unsigned long long* bar, unsigned long long target, int ctid, int debug = 0) {
  if (ctid == 0 && !(debug & 1)) {
    unsigned long long v;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
      if (v >= target) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");      // one acquire for the whole wait, not one per poll
        break;
      }
      if (clock64() - t0 > 4000000000ll) {     // a lost CTA must surface as a trapped kernel, never as a hung GPU
        printf("[trtllm_b200] decode_step grid barrier timed out (block %d, %llu < %llu)\n", blockIdx.x, v, target);
        __trap();
      }
    }
  }
  cbar();
}

// ---- one projection, as seen by this CTA ---------------------------------------------------------------------------
struct Phase {
  const uint8_t* w;
  const void* scale;
  int K, n_out, R, kind, nseg, o0, rows, ostride;   // this CTA's output channels: o0 + ostride * j, j < rows
  uint32_t rowbytes;
};

__device__ __forceinline__ Phase make_phase(const uint8_t* w, const void* scale, int kind, int K, int n_out, int R,
                                            int interleave) {
  Phase f;
  f.w = w; f.scale = scale; f.kind = kind; f.K = K; f.n_out = n_out; f.R = R;
  f.rowbytes = (uint32_t) (K / epc_of(kind)) * 16u;
  f.nseg = (int) ((f.rowbytes + kSeg - 1) / kSeg);
  if (interleave) {
    // channels dealt round-robin: at any instant the chip streams ONE window of ~G consecutive weight rows (DRAM pages stay
    // open) instead of G far-apart streams
    f.o0 = blockIdx.x; f.ostride = gridDim.x;
    f.rows = (int) blockIdx.x < n_out ? (n_out - (int) blockIdx.x + (int) gridDim.x - 1) / (int) gridDim.x : 0;
  } else {
    f.o0 = (int) ((long long) n_out * blockIdx.x / gridDim.x);
    f.ostride = 1;
    f.rows = (int) ((long long) n_out * (blockIdx.x + 1) / gridDim.x) - f.o0;
  }
  return f;
}
// idx = 4 * layer + {0 qkv, 1 dense, 2 gate|up, 3 down}; idx = 4 * n_layers: lm_head (always fp16, LQ/quant.py:58-59)
__device__ __forceinline__ Phase phase_of(const Params& p, int idx) {
  if (idx == 4 * p.n_layers) return make_phase(p.lm_head, nullptr, kF16, p.hidden, p.vocab_l, 1, p.interleave);
  const Layer& L = p.layers[idx >> 2];
  switch (idx & 3) {
    case 0: return make_phase(L.w_qkv, L.s_qkv, p.kind, p.hidden, 3 * p.hid_l, 1, p.interleave);
    case 1: return make_phase(L.w_dense, L.s_dense, p.kind, p.hid_l, p.hidden, 1, p.interleave);
    case 2: return make_phase(L.w_fc, L.s_fc, p.kind, p.hidden, p.inter_l, 2, p.interleave);
    default: return make_phase(L.w_proj, L.s_proj, p.kind, p.inter_l, p.hidden, 1, p.interleave);
  }
}

// ---- producer warp: every weight byte this CTA will need during the step, in order, regardless of phase barriers -------
// One thread needs ~200 cycles per copy (mbarrier wait on the slot, expect_tx, cp.async.bulk): at one 4 KB stage per
// ~180 cycles needed per SM that serial chain capped the kernel at ~3 TB/s.  The ring therefore has S = 32 (or 16) slots
// and producer lane l OWNS slot l: it issues the stages l, l + S, l + 2S, ... of the step's global stage sequence, so the
// uses of one slot are issued in order by one thread (mbarrier parity waits are only sound for consecutive phases) while
// up to 32 copies are in flight.  The lanes poll their slots with the non-blocking test_wait and the warp stays
// converged: a lane whose slot is still being read simply skips the round.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

struct Cursor {                      // a lane's position in the step's global stage sequence
  uint32_t gi, gbase, nst;
  int idx;
  bool done;
  Phase f;
};
__device__ __forceinline__ void cursor_seek(const Params& p, Cursor& c) {
  const int last = 4 * p.n_layers;
  while (!c.done && c.gi >= c.gbase + c.nst) {          // the stage lies in a later projection
    c.gbase += c.nst;
    if (++c.idx > last) { c.done = true; break; }
    c.f = phase_of(p, c.idx);
    c.nst = (uint32_t) (c.f.R * c.f.rows * c.f.nseg);
  }
}
__device__ __forceinline__ const uint8_t* cursor_src(const Cursor& c, uint32_t& bytes) {
  const int i = (int) (c.gi - c.gbase);
  const int per_r = c.f.rows * c.f.nseg;
  const int r = i >= per_r ? 1 : 0;
  const int rem = i - r * per_r;
  const int row = rem / c.f.nseg, seg = rem - row * c.f.nseg;
  const uint32_t off = (uint32_t) seg * kSeg;
  bytes = min((uint32_t) kSeg, c.f.rowbytes - off);
  return c.f.w + (size_t) (r * c.f.n_out + c.f.o0 + row * c.f.ostride) * c.f.rowbytes + off;
}

__device__ __forceinline__ void producer(const Params& p, uint8_t* ring, uint64_t* full, uint64_t* empty, int lane) {
  const uint64_t pol = policy_evict_first();     // weights are read once per step: keep K/V and activations in L2
  const uint32_t S = (uint32_t) p.stages;
  // slot cursors: c[k] walks the stages lane + 32 k, + S, + 2S, ... of the step's global stage sequence (slot lane + 32 k)
  Cursor c[2];
  uint32_t use[2] = {0, 0};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    c[k].gi = (uint32_t) (lane + 32 * k); c[k].gbase = 0; c[k].idx = 0; c[k].done = lane + 32 * k >= (int) S;
    c[k].f = phase_of(p, 0);
    c[k].nst = (uint32_t) (c[k].f.R * c[k].f.rows * c[k].f.nseg);
  }
  // Optional third cursor: cp.async.bulk.prefetch.L2 of the stages lane, lane + 32, ... (all lanes together: every stage
  // once) up to p.prefetch stages ahead of what the ring has requested — L2 as the buffer behind the 192 KB ring, so that
  // HBM keeps streaming while the consumers sit in a phase boundary (A/B switch TB_DS_PREFETCH; 0 = off).
  Cursor pf = c[0];
  pf.gi = (uint32_t) lane;
  pf.done = false;
  const long long t0 = clock64();
  for (;;) {
#pragma unroll
    for (int k = 0; k < 2; ++k) cursor_seek(p, c[k]);
    if (__all_sync(0xffffffffu, c[0].done && c[1].done)) break;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int slot = lane + 32 * k;
      if (!c[k].done && mbar_test_wait(&empty[slot], (use[k] & 1u) ^ 1u)) {
        uint32_t bytes;
        const uint8_t* src = cursor_src(c[k], bytes);
        mbar_expect_tx(&full[slot], bytes);
        bulk_load_1d_hint(ring + (size_t) slot * kSeg, src, bytes, &full[slot], pol);
        c[k].gi += S;
        ++use[k];
      }
    }
    const uint32_t head = __shfl_sync(0xffffffffu, c[0].gi, 0);      // lane 0's next stage ~ the ring's request frontier
    if (p.prefetch > 0 && !pf.done && pf.gi < head + S + (uint32_t) p.prefetch) {
      while (pf.gi < head + S) pf.gi += 32;                           // never behind the ring
      cursor_seek(p, pf);
      if (!pf.done) {
        uint32_t bytes;
        const uint8_t* src = cursor_src(pf, bytes);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
        pf.gi += 32;
      }
    }
    if (clock64() - t0 > 8000000000ll) {          // a wedged pipeline must surface as a trapped kernel, never as a hang
      if (lane == 0) printf("[trtllm_b200] decode_step producer timed out (block %d)\n", blockIdx.x);
      __trap();
    }
  }
}

// ---- sum / max over the 512 consumer threads ---------------------------------------------------------------------------
__device__ __forceinline__ float cta_reduce(float v, float* red, bool is_max, int ctid) {
  v = is_max ? warp_max(v) : warp_sum(v);
  cbar();
  if ((ctid & 31) == 0) red[ctid >> 5] = v;
  cbar();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kCW; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// ---- activation staging with the fused prologue (same arithmetic as gemv.cu / norm_quant.cu) ----------------------------
// mode 0: copy; 1: RMSNorm; 2: RMSNorm + dynamic per-token int8; 3: dynamic per-token int8.  rows[m]: fp16 [K] in global
// memory, written by other CTAs earlier in this launch (read through L2).
// RMSNorm weights of the coming phase: independent of every activation, so they are requested BEFORE the grid barrier
__device__ __forceinline__ void load_gamma(const __half* gamma, int K, uint4 (&g)[3], int ctid) {
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int i = (it * kCT + ctid) * 8;
    g[it] = (gamma && i < K) ? *reinterpret_cast<const uint4*>(gamma + i) : make_uint4(0, 0, 0, 0);
  }
}

// Tensor parallel: the row is not in memory yet — it is residual + sum over ranks of the fp16 partials every rank pushed
// into this rank's scratch (rank order, fp32; rounding points as allreduce.cu: fp16(sum), then fp16(sum + residual)).
// CTA 0 also stores the reduced row (the new residual stream) for the epilogues that add it later.
// No barrier guards the partials: every 4-byte granule carries the 16-bit tag of its exchange next to the fp16 value (the
// NCCL "LL" idea: data and flag in one atomic store), so the reader simply polls the granules it needs until all of them
// show the tag — one NVLink hop after the producer's store, no fence, no flag round trip.
struct ReduceSrc {
  const uint8_t* partial;     // [world][kMaxRows][hidden] granules of the current set (local scratch), or NULL: plain rows
  __half* store;              // [kMaxRows][hidden] (hA or hB)
  int world, hidden;
  uint32_t tag;               // 16-bit tag of the exchange
};

template <int XFMT, int MB, int E>
__device__ __forceinline__ void stage_x(const Params& p, const __half* const (&rows)[MB], int K, int mode,
                                        const uint4 (&gam)[3], uint8_t* xs, float* srow, float* red, int ctid,
                                        const ReduceSrc rs = ReduceSrc{nullptr, nullptr, 1, 0, 0}) {
  constexpr int XB = xbytes_of(XFMT);
  constexpr int NV = E * XB / 16;                          // 16-byte vectors per chunk
  const int xstride = xrow_bytes(K, E, XB);
  constexpr int IT = 3;                                   // K <= 3 * 512 * 8 = 12288 stays in registers
#pragma unroll 1
  for (int m = 0; m < MB; ++m) {
    if (m >= p.B) {
      for (int i = ctid * 16; i < xstride; i += kCT * 16) *reinterpret_cast<uint4*>(xs + (size_t) m * xstride + i) = make_uint4(0, 0, 0, 0);
      if (ctid == 0) srow[m] = 0.f;
      continue;
    }
    const __half* xr = rows[m];
    uint4 raw[IT];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int i = (it * kCT + ctid) * 8;
      raw[it] = i < K ? __ldcg(reinterpret_cast<const uint4*>(xr + i)) : make_uint4(0, 0, 0, 0);
    }
    if (rs.partial) {
      // rows[m] is the residual; add the rank-ordered sum of the partials
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int i = (it * kCT + ctid) * 8;
        if (i < K) {
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = 0.f;
          for (int r = 0; r < rs.world; ++r) {
            const uint8_t* src = rs.partial + ((size_t) (r * kMaxRows + m) * rs.hidden + i) * 4;
            uint4 a, b;
            const long long t0 = clock64();
            for (;;) {
              a = ld_volatile_v4(src);
              b = ld_volatile_v4(src + 16);
              const uint32_t t = rs.tag;
              if ((a.x >> 16) == t && (a.y >> 16) == t && (a.z >> 16) == t && (a.w >> 16) == t && (b.x >> 16) == t &&
                  (b.y >> 16) == t && (b.z >> 16) == t && (b.w >> 16) == t)
                break;
              if (clock64() - t0 > 40000000000ll) {      // a peer died or never launched: trap instead of hanging the GPU
                printf("[trtllm_b200] decode_step partial-sum exchange timed out (waiting for rank %d, block %d)\n", r, blockIdx.x);
                __trap();
              }
            }
            const uint32_t g8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += __half2float(__ushort_as_half((unsigned short) (g8[j] & 0xffffu)));
          }
          __half2* h = reinterpret_cast<__half2*>(&raw[it]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 a = __half22float2(__floats2half2_rn(acc[2 * j], acc[2 * j + 1])), b = __half22float2(h[j]);
            h[j] = __floats2half2_rn(a.x + b.x, a.y + b.y);
          }
          if (blockIdx.x == 0) *reinterpret_cast<uint4*>(rs.store + (size_t) m * rs.hidden + i) = raw[it];
        }
      }
    }
    if (mode == 1 || mode == 2) {
      float sq = 0.f;
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const __half2* h = reinterpret_cast<const __half2*>(&raw[it]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h[j]);
          sq += f.x * f.x + f.y * f.y;
        }
      }
      sq = cta_reduce(sq, red, false, ctid);
      const float inv = rsqrtf(sq / K + p.eps);
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int i = (it * kCT + ctid) * 8;
        if (i < K) {
          const __half2* g = reinterpret_cast<const __half2*>(&gam[it]);
          __half2* h = reinterpret_cast<__half2*>(&raw[it]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]), gg = __half22float2(g[j]);
            h[j] = __floats2half2_rn(f.x * inv * gg.x, f.y * inv * gg.y);
          }
        }
      }
    }
    if constexpr (XFMT == kXInt8) {
      float amax = 0.f;
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const __half2* h = reinterpret_cast<const __half2*>(&raw[it]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h[j]);
          amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
        }
      }
      amax = fmaxf(cta_reduce(amax, red, true, ctid), __half2float(__float2half_rn(1e-6f)));
      const float qs = 127.f / amax;
      if (ctid == 0) srow[m] = amax / 127.f;
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int i = (it * kCT + ctid) * 8;
        if (i < K) {
          const __half2* h = reinterpret_cast<const __half2*>(&raw[it]);
          float f[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 t = __half22float2(h[j]);
            f[2 * j] = t.x * qs;
            f[2 * j + 1] = t.y * qs;
          }
          uint2 o;
          o.x = pack4_i8(f[0], f[1], f[2], f[3]);
          o.y = pack4_i8(f[4], f[5], f[6], f[7]);
          const int c = i / E, e = i % E;                   // 8 int8 = half of the chunk's single vector
          *reinterpret_cast<uint2*>(xs + (size_t) m * xstride + xvec_off(c, e / 16, NV) + (e % 16)) = o;
        }
      }
    } else if constexpr (XFMT == kXFloat) {
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int i = (it * kCT + ctid) * 8;
        if (i < K) {
          const __half2* h = reinterpret_cast<const __half2*>(&raw[it]);
          const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), c = __half22float2(h[2]), d = __half22float2(h[3]);
          const int ch = i / E, j = (i % E) / 4;              // 8 floats = vectors j, j + 1 of chunk ch
          uint8_t* row = xs + (size_t) m * xstride;
          *reinterpret_cast<float4*>(row + xvec_off(ch, j, NV)) = make_float4(a.x, a.y, b.x, b.y);
          *reinterpret_cast<float4*>(row + xvec_off(ch, j + 1, NV)) = make_float4(c.x, c.y, d.x, d.y);
        }
      }
    } else {
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int i = (it * kCT + ctid) * 8;
        if (i < K) *reinterpret_cast<uint4*>(xs + (size_t) m * xstride + xvec_off(i / E, (i % E) / 8, NV)) = raw[it];
      }
    }
  }
  cbar();
}

// ---- one 16-byte weight chunk against the staged activations of MB rows ------------------------------------------------
template <int KIND, int XFMT, int MB>
__device__ __forceinline__ void chunk_dot(const uint4& wq, const uint8_t* xs, int c, int xstride, float (&acc)[MB],
                                          int (&iacc)[MB]) {
  constexpr int NV = epc_of(KIND) * xbytes_of(XFMT) / 16;
  if constexpr (KIND == kA8W8) {
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const uint4 xv = lds128(xs + (size_t) m * xstride + xvec_off(c, 0, NV));
      iacc[m] = __dp4a((int) wq.x, (int) xv.x, iacc[m]);
      iacc[m] = __dp4a((int) wq.y, (int) xv.y, iacc[m]);
      iacc[m] = __dp4a((int) wq.z, (int) xv.z, iacc[m]);
      iacc[m] = __dp4a((int) wq.w, (int) xv.w, iacc[m]);
    }
  } else {
    constexpr int E = epc_of(KIND);
    __half2 wh[E / 2];
    if constexpr (KIND == kF16) {
      const __half2* w2 = reinterpret_cast<const __half2*>(&wq);
#pragma unroll
      for (int j = 0; j < 4; ++j) wh[j] = w2[j];
    } else if constexpr (KIND == kW8) {
      i8x4_to_h2x2(wq.x, wh[0], wh[1]);
      i8x4_to_h2x2(wq.y, wh[2], wh[3]);
      i8x4_to_h2x2(wq.z, wh[4], wh[5]);
      i8x4_to_h2x2(wq.w, wh[6], wh[7]);
    } else {
      i4x8_to_h2x4(wq.x, wh + 0);
      i4x8_to_h2x4(wq.y, wh + 4);
      i4x8_to_h2x4(wq.z, wh + 8);
      i4x8_to_h2x4(wq.w, wh + 12);
    }
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      float a0 = 0.f, a1 = 0.f;
      if constexpr (XFMT == kXFloat) {
        const uint8_t* xp = xs + (size_t) m * xstride + xvec_off(c, 0, NV);
#pragma unroll
        for (int q = 0; q < E / 4; ++q) {
          const float4 xv = *reinterpret_cast<const float4*>(xp + q * 512);
          const float2 wa = __half22float2(wh[2 * q]), wb = __half22float2(wh[2 * q + 1]);
          a0 = fmaf(wa.x, xv.x, a0);
          a1 = fmaf(wa.y, xv.y, a1);
          a0 = fmaf(wb.x, xv.z, a0);
          a1 = fmaf(wb.y, xv.w, a1);
        }
      } else {
        const uint8_t* xp = xs + (size_t) m * xstride + xvec_off(c, 0, NV);
#pragma unroll
        for (int q = 0; q < E / 8; ++q) {
          const uint4 xv = lds128(xp + q * 512);
          const __half2* x2 = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 a = __half22float2(wh[q * 4 + j]), b = __half22float2(x2[j]);
            a0 = fmaf(a.x, b.x, a0);
            a1 = fmaf(a.y, b.y, a1);
          }
        }
      }
      acc[m] += a0 + a1;
    }
  }
}

// ---- the stages of one phase: warp w consumes stages w, w + 16, ... and leaves one partial per (stage, row) -------------
template <int KIND, int XFMT, int MB>
__device__ __forceinline__ void run_phase(const Phase& f, uint32_t g, int S, const uint8_t* ring, uint64_t* full,
                                          uint64_t* empty, const uint8_t* xs, float* part, int cwarp, int lane,
                                          int debug = 0) {
  const int rows = f.rows;
  const int nst = f.R * rows * f.nseg;
  const int xstride = xrow_bytes(f.K, epc_of(KIND), xbytes_of(XFMT));
  // stage i of this phase: ring slot (g + i) % S with use count (g + i) / S, row segment i % nseg — kept incrementally
  uint32_t q = (g + (uint32_t) cwarp) / (uint32_t) S;
  int slot = (int) (g + (uint32_t) cwarp - q * (uint32_t) S);
  int seg = cwarp % f.nseg;
  const int seg_step = kCW % f.nseg;
  for (int i = cwarp; i < nst; i += kCW) {
    const uint32_t off = (uint32_t) seg * kSeg;
    const int nchunks = (int) (min((uint32_t) kSeg, f.rowbytes - off) >> 4);
    const int cbase = (int) (off >> 4);
    const uint8_t* st = ring + (size_t) slot * kSeg;
    const uint32_t parity = q & 1u;
    const int slot_now = slot;
    slot += kCW;
    while (slot >= S) { slot -= S; ++q; }
    seg += seg_step;
    if (seg >= f.nseg) seg -= f.nseg;
    float acc[MB];
    int iacc[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) { acc[m] = 0.f; iacc[m] = 0; }
    mbar_wait(&full[slot_now], parity);
    if (!(debug & 2)) {
#pragma unroll 4
      for (int c = lane; c < nchunks; c += 32) {
        const uint4 wq = lds128(st + (size_t) c * 16);
        chunk_dot<KIND, XFMT, MB>(wq, xs, cbase + c, xstride, acc, iacc);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot_now]);      // the slot is free as soon as every lane has read its chunks
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      if constexpr (KIND == kA8W8) {
        int v = iacc[m];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) reinterpret_cast<int*>(part)[i * MB + m] = v;
      } else {
        const float v = warp_sum(acc[m]);
        if (lane == 0) part[i * MB + m] = v;
      }
    }
  }
  cbar();
}

// ---- fused epilogue over this CTA's output channels -------------------------------------------------------------------
// mode 0: y[m][n] = fp16(v) (+ residual);  1: SwiGLU (R = 2);  2: fp32 logits (+ copy in lgs for the argmax)
// A thread owns at most kEpiItems (output channel, row) pairs; their per-channel scales and residual values are loaded
// BEFORE the phase's stages are consumed (they depend on nothing the phase computes), so the tail of the phase — the part
// every other CTA waits for at the grid barrier — is shared-memory sums, arithmetic and stores only.
constexpr int kEpiItems = 2;
struct EpiPre { float sc[kEpiItems][2]; float res[kEpiItems]; };

template <int KIND, int MB>
__device__ __forceinline__ void epilogue_prefetch(const Params& p, const Phase& f, const __half* const (&resid)[MB], EpiPre& e,
                                                  int ctid) {
  const int rows = f.rows;
#pragma unroll
  for (int k = 0; k < kEpiItems; ++k) {
    const int t = ctid + k * kCT;
    e.sc[k][0] = e.sc[k][1] = 1.f;
    e.res[k] = 0.f;
    if (t >= rows * MB) continue;
    const int row = t / MB, m = t % MB;
    if (m >= p.B) continue;
    const int n = f.o0 + row * f.ostride;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r < f.R) {
        if constexpr (KIND == kA8W8) e.sc[k][r] = reinterpret_cast<const float*>(f.scale)[n + r * f.n_out];
        else if constexpr (KIND == kW8 || KIND == kW4) e.sc[k][r] = __half2float(reinterpret_cast<const __half*>(f.scale)[n + r * f.n_out]);
      }
    }
    if (resid[m]) e.res[k] = ldcg_h(resid[m] + n);
  }
}

template <int KIND, int MB>
__device__ __forceinline__ void epilogue(const Params& p, const Phase& f, int mode, const float* part, const float* srow,
                                         const EpiPre& e, bool has_resid, __half* y, int ldy, float* lgs, int ctid) {
  const int rows = f.rows;
#pragma unroll
  for (int k = 0; k < kEpiItems; ++k) {
    const int t = ctid + k * kCT;
    if (t >= rows * MB) continue;
    const int row = t / MB, m = t % MB;
    if (m >= p.B) continue;
    const int n = f.o0 + row * f.ostride;
    float v[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r >= f.R) continue;
      const int base = ((r * rows + row) * f.nseg) * MB + m;
      if constexpr (KIND == kA8W8) {
        int s = 0;
        for (int sg = 0; sg < f.nseg; ++sg) s += reinterpret_cast<const int*>(part)[base + sg * MB];
        // reference grouping: accum * (scale_col * scale_row)  (epilogue_per_row_per_col_scale.h:325,341)
        v[r] = (float) s * (e.sc[k][r] * srow[m]);
      } else {
        float s = 0.f;
        for (int sg = 0; sg < f.nseg; ++sg) s += part[base + sg * MB];
        if constexpr (KIND == kW8 || KIND == kW4) s *= e.sc[k][r];
        v[r] = s;
      }
    }
    if (mode == 2) {
      if (p.tp > 1) {
        const TpLayout l = tp_layout(p.hidden, p.vocab, gridDim.x);
        for (int r = 0; r < p.tp; ++r)
          reinterpret_cast<float*>(p.peer[r] + l.logits)[(size_t) m * p.vocab + (size_t) p.rank * p.vocab_l + n] = v[0];
      } else {
        p.logits[(size_t) m * p.vocab + n] = v[0];
      }
      lgs[row * MB + m] = v[0];
    } else if (mode == 3) {
      // row-parallel projection under tensor parallelism: this rank's fp16 partial, pushed into every rank's scratch
      const uint32_t gran = (uint32_t) __half_as_ushort(__float2half_rn(v[0])) | ((uint32_t) (ldy >> 1) << 16);   // ldy = tag << 1 | set
      const TpLayout l = tp_layout(p.hidden, p.vocab, gridDim.x);
      const size_t off = l.partial + ((size_t) (((ldy & 1) * 8 + p.rank) * kMaxRows + m) * p.hidden + n) * 4;
      for (int r = 0; r < p.tp; ++r) *reinterpret_cast<volatile uint32_t*>(p.peer[r] + off) = gran;
    } else if (mode == 1) {
      const float gte = __half2float(__float2half_rn(v[0])), up = __half2float(__float2half_rn(v[1]));
      y[(size_t) m * ldy + n] = __float2half_rn(__half2float(__float2half_rn(silu_f(gte))) * up);
    } else {
      __half oh = __float2half_rn(v[0]);
      if (has_resid) oh = __float2half_rn(__half2float(oh) + e.res[k]);
      y[(size_t) m * ldy + n] = oh;
    }
  }
}

// ---- fused masked multi-head attention of one (sequence, head): RoPE, KV append, QK^T . softmax . V ---------------------
// Same arithmetic as mmha.cu's FMA variant with one split (reference order: p * 1/(sum + 1e-6) -> fp16 -> P.V).
template <bool INT8>
__device__ __forceinline__ void unpack16(const uint4& r, float* f) {
  if constexpr (INT8) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t u = w[i] ^ 0x80808080u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t bits;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(bits) : "r"(u), "r"(0x4B000000u), "r"(0x7650u + j));
        f[i * 4 + j] = __uint_as_float(bits) - 8388736.f;
      }
    }
  } else {
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}
template <bool INT8>
__device__ __forceinline__ void attention_item(const Params& p, const Layer& L, int b, int h, float* scr, int ctid,
                                               float rope_c, float rope_s) {
  constexpr int LPK = INT8 ? 8 : 16, DPL = kDh / LPK, KPI = kCT / LPK, ELT = INT8 ? 1 : 2, UN = INT8 ? 4 : 8;
  float* q_s = scr;                                               // [128]
  __half* kcur_s = reinterpret_cast<__half*>(scr + kDh);          // [128]
  __half* vcur_s = kcur_s + kDh;                                  // [128]
  float* red = scr + 2 * kDh;                                     // [2 * 16]
  float* s_s = red + 2 * kCW;                                     // [S_max + 1]
  float* o_red = s_s + ((p.S_max + 1 + 3) & ~3);                  // [KPI][128]
  const int lane = ctid & 31, warp = ctid >> 5;
  const int H = p.Hl, hidden = H * kDh;
  const int tlen = p.seq_lens[b];                                 // positions [0, tlen) are cached
  const int max_in = p.max_in[0];
  const int in_len = p.in_lens[b];
  const int len = tlen;
  const float kv_dq = INT8 ? L.kv_qo[0] : 1.f;
  const size_t seq_stride = (size_t) 2 * H * p.S_max * kDh * ELT;
  uint8_t* kbase = L.kv + (size_t) b * seq_stride + (size_t) h * p.S_max * kDh * ELT;
  uint8_t* vbase = kbase + (size_t) H * p.S_max * kDh * ELT;

  const __half* qrow = p.qkv + (size_t) b * 3 * hidden + (size_t) h * kDh;
  if (ctid < kDh / 2) {
    const float c = rope_c, s = rope_s;
    const int i0 = ctid, i1 = ctid + kDh / 2;
    const float qa = ldcg_h(qrow + i0), qb = ldcg_h(qrow + i1);
    q_s[i0] = __half2float(__float2half_rn(c * qa - s * qb));
    q_s[i1] = __half2float(__float2half_rn(c * qb + s * qa));
    const float ka = ldcg_h(qrow + hidden + i0), kb = ldcg_h(qrow + hidden + i1);
    kcur_s[i0] = __float2half_rn(c * ka - s * kb);
    kcur_s[i1] = __float2half_rn(c * kb + s * ka);
    vcur_s[i0] = __float2half_rn(ldcg_h(qrow + 2 * hidden + i0));
    vcur_s[i1] = __float2half_rn(ldcg_h(qrow + 2 * hidden + i1));
  }
  cbar();
  if (ctid < kDh / 8) {
    const int d0 = ctid * 8;
    if constexpr (INT8) {
      const float qs = L.kv_oq[0];
      uint2 kq, vq;
      kq.x = pack4_i8(__half2float(kcur_s[d0]) * qs, __half2float(kcur_s[d0 + 1]) * qs, __half2float(kcur_s[d0 + 2]) * qs,
                      __half2float(kcur_s[d0 + 3]) * qs);
      kq.y = pack4_i8(__half2float(kcur_s[d0 + 4]) * qs, __half2float(kcur_s[d0 + 5]) * qs, __half2float(kcur_s[d0 + 6]) * qs,
                      __half2float(kcur_s[d0 + 7]) * qs);
      vq.x = pack4_i8(__half2float(vcur_s[d0]) * qs, __half2float(vcur_s[d0 + 1]) * qs, __half2float(vcur_s[d0 + 2]) * qs,
                      __half2float(vcur_s[d0 + 3]) * qs);
      vq.y = pack4_i8(__half2float(vcur_s[d0 + 4]) * qs, __half2float(vcur_s[d0 + 5]) * qs, __half2float(vcur_s[d0 + 6]) * qs,
                      __half2float(vcur_s[d0 + 7]) * qs);
      *reinterpret_cast<uint2*>(kbase + (size_t) tlen * kDh + d0) = kq;
      *reinterpret_cast<uint2*>(vbase + (size_t) tlen * kDh + d0) = vq;
    } else {
      *reinterpret_cast<uint4*>(kbase + ((size_t) tlen * kDh + d0) * 2) = *reinterpret_cast<uint4*>(&kcur_s[d0]);
      *reinterpret_cast<uint4*>(vbase + ((size_t) tlen * kDh + d0) * 2) = *reinterpret_cast<uint4*>(&vcur_s[d0]);
    }
  }

  const int grp = ctid / LPK, gl = ctid % LPK;
  float qreg[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) qreg[i] = q_s[gl * DPL + i];
  const float qk_scale = kv_dq * p.inv_sqrt_dh;
  float lmax = -3.0e38f;
  for (int i = grp; i - grp < len; i += KPI * UN) {            // trip count uniform across the warp (shuffles)
    uint4 raw[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int ii = i + u * KPI;
      raw[u] = make_uint4(0, 0, 0, 0);
      if (ii < len) raw[u] = ldg_nc_v4(kbase + ((size_t) ii * kDh + gl * DPL) * ELT);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int ii = i + u * KPI;
      float kf[DPL];
      unpack16<INT8>(raw[u], kf);
      float d = 0.f;
#pragma unroll
      for (int j = 0; j < DPL; ++j) d = fmaf(qreg[j], kf[j], d);
#pragma unroll
      for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (ii < len && gl == 0) {
        d *= qk_scale;
        if (ii >= in_len && ii < max_in) d = -3.0e38f;           // padding of a shorter prompt in the padded batch
        s_s[ii] = d;
        lmax = fmaxf(lmax, d);
      }
    }
  }
  if (warp == 0) {
    // current token: unquantised k (decoderMaskedMultiheadAttentionTemplate.h:1511-1549)
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) d = fmaf(q_s[lane * 4 + j], __half2float(kcur_s[lane * 4 + j]), d);
    d = warp_sum(d) * p.inv_sqrt_dh;
    if (lane == 0) {
      s_s[len] = d;
      lmax = fmaxf(lmax, d);
    }
  }
  const int n_s = len + 1;
  lmax = warp_max(lmax);
  if (lane == 0) red[warp] = lmax;
  cbar();
  float m_s = red[0];
#pragma unroll
  for (int w = 1; w < kCW; ++w) m_s = fmaxf(m_s, red[w]);
  float lsum = 0.f;
  for (int i = ctid; i < n_s; i += kCT) {
    const float sv = s_s[i];
    const float e = sv <= -1.0e38f ? 0.f : __expf(sv - m_s);
    s_s[i] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red[kCW + warp] = lsum;
  cbar();
  float l_s = 0.f;
#pragma unroll
  for (int w = 0; w < kCW; ++w) l_s += red[kCW + w];
  const float inv_sum = __fdividef(1.f, l_s + 1.e-6f);
  for (int i = ctid; i < n_s; i += kCT) s_s[i] = __half2float(__float2half_rn(s_s[i] * inv_sum));   // p -> fp16 (Template.h:1765)
  cbar();

  float acc[DPL];
#pragma unroll
  for (int j = 0; j < DPL; ++j) acc[j] = 0.f;
  for (int i = grp; i - grp < len; i += KPI * UN) {
    uint4 raw[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int ii = i + u * KPI;
      raw[u] = make_uint4(0, 0, 0, 0);
      if (ii < len) raw[u] = ldg_nc_v4(vbase + ((size_t) ii * kDh + gl * DPL) * ELT);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int ii = i + u * KPI;
      const float pv = ii < len ? s_s[ii] : 0.f;
      float vf[DPL];
      unpack16<INT8>(raw[u], vf);
#pragma unroll
      for (int j = 0; j < DPL; ++j) acc[j] = fmaf(pv, vf[j], acc[j]);
    }
  }
  // the 32 / LPK key groups of a warp hold partial outputs for the same dims: fold them with shuffles, one row per warp
#pragma unroll
  for (int j = 0; j < DPL; ++j) {
#pragma unroll
    for (int o = LPK; o < 32; o <<= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  if (lane < LPK) {
#pragma unroll
    for (int j = 0; j < DPL; ++j) o_red[warp * kDh + gl * DPL + j] = acc[j] * kv_dq;
  }
  cbar();
  if (ctid < kDh) {
    float o = 0.f;
#pragma unroll
    for (int g2 = 0; g2 < kCW; ++g2) o += o_red[g2 * kDh + ctid];
    o = fmaf(s_s[len], __half2float(vcur_s[ctid]), o);
    p.att[(size_t) b * hidden + h * kDh + ctid] = __float2half_rn(o);
  }
  cbar();     // scratch is reused by the next item
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
template <int KIND, int MB>
__global__ void __launch_bounds__(kThreads, 1) decode_step_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kMaxStages;
  float* red = reinterpret_cast<float*>(smem + kOffRed);          // [32]
  float* srow = red + 64;                                         // [MB] per-token scales (W8A8)
  float* part = reinterpret_cast<float*>(smem + kOffPart);
  float* lgs = reinterpret_cast<float*>(smem + off_lg(MB));
  uint8_t* xs = smem + p.xs_off;
  uint8_t* ring = smem + p.ring_off;
  const int tid = threadIdx.x;
  const int S = p.stages;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (tid >= kCT) {
    producer(p, ring, full, empty, tid - kCT);
    return;
  }

  constexpr int XM = XF<KIND, MB>::v;       // activation format of the model's projections
  constexpr int XL = XF<kF16, MB>::v;       // ... of the fp16 lm_head
  const int ctid = tid, lane = tid & 31, cwarp = tid >> 5;
  const unsigned long long G = gridDim.x;
  unsigned long long target = 0;
  uint32_t g = 0;                           // running stage counter (the producer counts the same sequence)
  auto stages_of = [](const Phase& f) { return (uint32_t) (f.R * f.rows * f.nseg); };

  const __half* none[MB];
  const __half* hrow[MB];
  uint4 gam[3];
  EpiPre epre;
  int tr_n = 0;
  DS_STAMP();
#pragma unroll
  for (int m = 0; m < MB; ++m) none[m] = nullptr;
  // tensor parallel: cross-GPU barriers completed before this launch (same on every rank), barriers done in this launch
  const bool tp = p.tp > 1;
  const TpLayout tl = tp_layout(p.hidden, p.vocab, gridDim.x);
  const uint32_t xbase = tp ? *reinterpret_cast<const volatile uint32_t*>(p.peer[p.rank] + tl.xepoch) : 0u;
  uint32_t xk = 0;
  bool xpending = false;                    // the barrier in front of the next phase is the cross-GPU one (end of step)
  bool llpending = false;                   // the next phase's input arrives as tagged partial sums: no barrier at all
  auto wait_phase = [&]() {
    if (xpending) xgrid_wait(p, xbase + 1, ctid);
    else if (!llpending && target) grid_wait(p.bar, target, ctid, p.debug);
    xpending = false;
    llpending = false;
  };
  auto arrive_cross = [&]() {
    target += G;
    xgrid_arrive(p, p.bar, target, xbase + 1, ctid);
    xpending = true;
  };
  // exchanges of row-parallel partial sums: numbered xbase * (2 L) + xk (same on every rank); set = parity, tag = low 16 bits
  const uint32_t ex_base = xbase * (uint32_t) (2 * p.n_layers) + 1u;
  auto ex_tag = [&](uint32_t k) { return (ex_base + k) & 0xffffu; };
  auto ex_set = [&](uint32_t k) { return (int) ((ex_base + k) & 1u); };
  auto reduce_src = [&](uint32_t k, __half* store) {
    return ReduceSrc{p.peer[p.rank] + tl.partial + (size_t) ex_set(k) * 8 * kMaxRows * p.hidden * 4, store, p.tp, p.hidden,
                     ex_tag(k)};
  };
  uint32_t last_ex = 0;                     // the most recent exchange this CTA published

#pragma unroll 1
  for (int li = 0; li < p.n_layers; ++li) {
    const Layer& L = p.layers[li];
    // the residual stream entering the layer: the embedding rows for layer 0, hA afterwards
#pragma unroll
    for (int m = 0; m < MB; ++m)
      hrow[m] = m < p.B ? (li == 0 ? p.emb + (size_t) p.ids[m] * p.hidden : p.hA + (size_t) m * p.hidden) : nullptr;

    // ---- QKV projection (RMSNorm / RmsnormQuantization prologue) --------------------------------------------------------
    {
      const Phase f = phase_of(p, 4 * li);
      load_gamma(L.ln_in, p.hidden, gam, ctid);
      epilogue_prefetch<KIND, MB>(p, f, none, epre, ctid);
      wait_phase();
      DS_STAMP();
      if (tp && li > 0) {
        // the stream entering this layer = hB + all-reduce of the previous layer's down projection; CTA 0 stores it in hA
        const __half* rrow[MB];
#pragma unroll
        for (int m = 0; m < MB; ++m) rrow[m] = m < p.B ? p.hB + (size_t) m * p.hidden : nullptr;
        stage_x<XM, MB, epc_of(KIND)>(p, rrow, p.hidden, KIND == kA8W8 ? 2 : 1, gam, xs, srow, red, ctid,
                                      reduce_src(last_ex, p.hA));
      } else {
        stage_x<XM, MB, epc_of(KIND)>(p, hrow, p.hidden, KIND == kA8W8 ? 2 : 1, gam, xs, srow, red, ctid);
      }
      DS_STAMP();
      run_phase<KIND, XM, MB>(f, g, S, ring, full, empty, xs, part, cwarp, lane, p.debug);
      DS_STAMP();
      epilogue<KIND, MB>(p, f, 0, part, srow, epre, false, p.qkv, 3 * p.hid_l, lgs, ctid);
      DS_STAMP();
      g += stages_of(f);
      grid_arrive(p.bar, ctid);
      target += G;
    }
    // ---- attention ----------------------------------------------------------------------------------------------------
    {
      // the rotary angle of this CTA's first (sequence, head) depends on the position only: computed (powf, sincos)
      // while waiting for the QKV projection
      float rope_c = 1.f, rope_s = 0.f;
      auto rope_of = [&](int b) {
        if (ctid < kDh / 2) {
          // inv_freq = t / pow(10000, 2j/rot)  (decoderMaskedMultiheadAttentionUtils.h:1511-1515)
          const int pos = p.seq_lens[b] - (p.max_in[0] - p.in_lens[b]);
          const float ang = (float) pos / powf(10000.0f, (2 * ctid) / (float) kDh);
          rope_c = cosf(ang);
          rope_s = sinf(ang);
        }
      };
      if ((int) blockIdx.x < p.B * p.Hl) rope_of(blockIdx.x / p.Hl);
      wait_phase();
      DS_STAMP();
      float* scr = reinterpret_cast<float*>(xs);
      for (int it = blockIdx.x; it < p.B * p.Hl; it += gridDim.x) {
        const int b = it / p.Hl, h = it % p.Hl;
        if (it != (int) blockIdx.x) rope_of(b);
        if (p.int8_kv) attention_item<true>(p, L, b, h, scr, ctid, rope_c, rope_s);
        else attention_item<false>(p, L, b, h, scr, ctid, rope_c, rope_s);
      }
      DS_STAMP();
      grid_arrive(p.bar, ctid);
      target += G;
    }
    // ---- dense + residual ------------------------------------------------------------------------------------------------
    {
      const Phase f = phase_of(p, 4 * li + 1);
      const __half* arow[MB];
#pragma unroll
      for (int m = 0; m < MB; ++m) arow[m] = m < p.B ? p.att + (size_t) m * p.hid_l : nullptr;
      load_gamma(nullptr, 0, gam, ctid);
      epilogue_prefetch<KIND, MB>(p, f, tp ? none : hrow, epre, ctid);   // residual = the stream entering the layer (older than QKV)
      wait_phase();
      DS_STAMP();
      stage_x<XM, MB, epc_of(KIND)>(p, arow, p.hid_l, KIND == kA8W8 ? 3 : 0, gam, xs, srow, red, ctid);
      DS_STAMP();
      run_phase<KIND, XM, MB>(f, g, S, ring, full, empty, xs, part, cwarp, lane, p.debug);
      DS_STAMP();
      g += stages_of(f);
      if (tp) {
        // row-parallel: fp16 partial to every rank; the residual add happens where the all-reduce is consumed
        last_ex = xk++;
        epilogue<KIND, MB>(p, f, 3, part, srow, epre, false, nullptr, (int) (ex_tag(last_ex) << 1) | ex_set(last_ex), lgs, ctid);
        DS_STAMP();
        llpending = true;
      } else {
        epilogue<KIND, MB>(p, f, 0, part, srow, epre, true, p.hB, p.hidden, lgs, ctid);
        DS_STAMP();
        grid_arrive(p.bar, ctid);
        target += G;
      }
    }
#pragma unroll
    for (int m = 0; m < MB; ++m) hrow[m] = m < p.B ? p.hB + (size_t) m * p.hidden : nullptr;
    // ---- gate | up + SwiGLU (RMSNorm prologue) --------------------------------------------------------------------------
    {
      const Phase f = phase_of(p, 4 * li + 2);
      load_gamma(L.ln_post, p.hidden, gam, ctid);
      epilogue_prefetch<KIND, MB>(p, f, none, epre, ctid);
      wait_phase();
      DS_STAMP();
      if (tp) {
        // hB = (stream entering the layer) + all-reduce of the attention output projection
        const __half* rrow[MB];
#pragma unroll
        for (int m = 0; m < MB; ++m)
          rrow[m] = m < p.B ? (li == 0 ? p.emb + (size_t) p.ids[m] * p.hidden : p.hA + (size_t) m * p.hidden) : nullptr;
        stage_x<XM, MB, epc_of(KIND)>(p, rrow, p.hidden, KIND == kA8W8 ? 2 : 1, gam, xs, srow, red, ctid,
                                      reduce_src(last_ex, p.hB));
      } else {
        stage_x<XM, MB, epc_of(KIND)>(p, hrow, p.hidden, KIND == kA8W8 ? 2 : 1, gam, xs, srow, red, ctid);
      }
      DS_STAMP();
      run_phase<KIND, XM, MB>(f, g, S, ring, full, empty, xs, part, cwarp, lane, p.debug);
      DS_STAMP();
      epilogue<KIND, MB>(p, f, 1, part, srow, epre, false, p.act, p.inter_l, lgs, ctid);
      DS_STAMP();
      g += stages_of(f);
      grid_arrive(p.bar, ctid);
      target += G;
    }
    // ---- down projection + residual -----------------------------------------------------------------------------------------
    {
      const Phase f = phase_of(p, 4 * li + 3);
      const __half* arow[MB];
#pragma unroll
      for (int m = 0; m < MB; ++m) arow[m] = m < p.B ? p.act + (size_t) m * p.inter_l : nullptr;
      wait_phase();
      // residual = hB, written by the dense phase of THIS layer two barriers ago: loaded after this barrier, but still
      // before the stages (its latency hides behind the weight stream)
      DS_STAMP();
      stage_x<XM, MB, epc_of(KIND)>(p, arow, p.inter_l, KIND == kA8W8 ? 3 : 0, gam, xs, srow, red, ctid);
      epilogue_prefetch<KIND, MB>(p, f, tp ? none : hrow, epre, ctid);
      DS_STAMP();
      run_phase<KIND, XM, MB>(f, g, S, ring, full, empty, xs, part, cwarp, lane, p.debug);
      DS_STAMP();
      g += stages_of(f);
      if (tp) {
        last_ex = xk++;
        epilogue<KIND, MB>(p, f, 3, part, srow, epre, false, nullptr, (int) (ex_tag(last_ex) << 1) | ex_set(last_ex), lgs, ctid);
        DS_STAMP();
        llpending = true;
      } else {
        epilogue<KIND, MB>(p, f, 0, part, srow, epre, true, p.hA, p.hidden, lgs, ctid);
        DS_STAMP();
        grid_arrive(p.bar, ctid);
        target += G;
      }
    }
  }

  // ---- ln_f + lm_head (fp16) -> fp32 logits, per-CTA argmax candidates ---------------------------------------------------
  {
    const Phase f = phase_of(p, 4 * p.n_layers);
#pragma unroll
    for (int m = 0; m < MB; ++m) hrow[m] = m < p.B ? p.hA + (size_t) m * p.hidden : nullptr;
    load_gamma(p.ln_f, p.hidden, gam, ctid);
    epilogue_prefetch<kF16, MB>(p, f, none, epre, ctid);
    wait_phase();
    if (tp) {
      const __half* rrow[MB];
#pragma unroll
      for (int m = 0; m < MB; ++m) rrow[m] = m < p.B ? p.hB + (size_t) m * p.hidden : nullptr;
      stage_x<XL, MB, epc_of(kF16)>(p, rrow, p.hidden, 1, gam, xs, srow, red, ctid, reduce_src(last_ex, p.hA));
    } else {
      stage_x<XL, MB, epc_of(kF16)>(p, hrow, p.hidden, 1, gam, xs, srow, red, ctid);
    }
    run_phase<kF16, XL, MB>(f, g, S, ring, full, empty, xs, part, cwarp, lane, p.debug);
    epilogue<kF16, MB>(p, f, 2, part, srow, epre, false, nullptr, 0, lgs, ctid);
    cbar();
    const int rows = f.rows;
    float* rv = red;
    int* ri = reinterpret_cast<int*>(red + kCW);
    for (int m = 0; m < p.B; ++m) {
      float bv = -3.4e38f;
      int bi = 0x7fffffff;
      for (int r = ctid; r < rows; r += kCT) {
        const float v = lgs[r * MB + m];
        const int n = f.o0 + r * f.ostride;
        if (v > bv || (v == bv && n < bi)) { bv = v; bi = n; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      cbar();
      if (lane == 0) { rv[cwarp] = bv; ri[cwarp] = bi; }
      cbar();
      if (ctid == 0) {
        for (int w = 1; w < kCW; ++w)
          if (rv[w] > bv || (rv[w] == bv && ri[w] < bi)) { bv = rv[w]; bi = ri[w]; }
        if (tp) {
          // vocabulary-parallel lm_head: this CTA's best (value, global index) to every rank
          const int2 e = make_int2(__float_as_int(bv), bi == 0x7fffffff ? bi : p.rank * p.vocab_l + bi);
          for (int r = 0; r < p.tp; ++r)
            reinterpret_cast<int2*>(p.peer[r] + tl.cand)[((size_t) p.rank * gridDim.x + blockIdx.x) * kMaxRows + m] = e;
        } else {
          p.cand_v[blockIdx.x * MB + m] = bv;
          p.cand_i[blockIdx.x * MB + m] = bi;
        }
      }
    }
    if (tp) {
      arrive_cross();
    } else {
      grid_arrive(p.bar, ctid);
      target += G;
    }
  }

  // ---- greedy token (lowest index wins ties, as tb_argmax) + device-side step bookkeeping, by CTA 0 ----------------------
  if (blockIdx.x != 0) return;
  wait_phase();
  const int pos = p.step_pos[0];
  if (ctid < p.B) {
    const int m = ctid;
    float bv = -3.4e38f;
    int bi = 0x7fffffff;
    if (tp) {
      const int2* cd = reinterpret_cast<const int2*>(p.peer[p.rank] + tl.cand);
      for (int c = 0; c < p.tp * (int) gridDim.x; ++c) {
        int2 e;
        asm volatile("ld.volatile.global.v2.s32 {%0,%1}, [%2];" : "=r"(e.x), "=r"(e.y) : "l"(cd + (size_t) c * kMaxRows + m));
        const float v = __int_as_float(e.x);
        if (v > bv || (v == bv && e.y < bi)) { bv = v; bi = e.y; }
      }
    } else {
      for (int c = 0; c < (int) gridDim.x; ++c) {
        const float v = __ldcg(p.cand_v + c * MB + m);
        const int i = __ldcg(p.cand_i + c * MB + m);
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
      }
    }
    p.next[m] = bi;
    p.ids[m] = bi;
    if (pos < p.out_stride) p.out_ids[(size_t) m * p.out_stride + pos] = bi;
    p.seq_lens[m] += 1;
  }
  cbar();
  if (ctid == 0) {
    p.step_pos[0] = pos + 1;
    *p.bar = 0ull;                 // every CTA has arrived at the last barrier and none reads the counter again
    if (tp) *reinterpret_cast<volatile uint32_t*>(p.peer[p.rank] + tl.xepoch) = xbase + 1;   // launches completed
  }
}

}  // namespace ds
}  // namespace tb

using namespace tb;

struct tb_decode_step {
  ds::Params p{};
  ds::Layer* d_layers = nullptr;
  float* d_cand_v = nullptr;
  int* d_cand_i = nullptr;
  unsigned long long* d_bar = nullptr;
  unsigned long long* d_trace = nullptr;
  const float* tp_logits = nullptr;
  int grid = 0, max_batch = 0, smem_max = 0;
  size_t smem[3] = {0, 0, 0};       // per MB in {1, 2, 4}
  int stages[3] = {0, 0, 0};
  uint32_t ring_off[3] = {0, 0, 0};
};

namespace {

template <int KIND, int MB>
int launch_t(tb_decode_step* d, int B, cudaStream_t stream) {
  const int mi = MB == 1 ? 0 : (MB == 2 ? 1 : 2);
  ds::Params p = d->p;
  p.B = B;
  p.stages = d->stages[mi];
  static const int pf_env = getenv("TB_DS_PREFETCH") ? atoi(getenv("TB_DS_PREFETCH")) : 0;   // A/B switch (stages)
  p.prefetch = pf_env;
  static const int il_env = getenv("TB_DS_INTERLEAVE") ? atoi(getenv("TB_DS_INTERLEAVE")) : 1;   // A/B switch
  p.interleave = d->p.tp > 1 ? 0 : il_env;   // tensor parallel: contiguous channels per CTA -> coalesced stores to the peers
  static const int dbg_env = getenv("TB_DS_DEBUG") ? atoi(getenv("TB_DS_DEBUG")) : 0;
  p.debug = dbg_env;
  p.xs_off = ds::off_xs(MB);
  p.ring_off = d->ring_off[mi];
  auto kern = ds::decode_step_kernel<KIND, MB>;
  static bool attr_done = false;      // per template instantiation
  if (!attr_done) {
    TB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, d->smem_max));
    int per_sm = 0;
    TB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ds::kThreads, d->smem[mi]));
    if (per_sm < 1) return -20;
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(d->grid);
  cfg.blockDim = dim3(ds::kThreads);
  cfg.dynamicSmemBytes = d->smem[mi];
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;     // all CTAs co-resident or the launch fails: the grid barrier cannot hang
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int) cudaLaunchKernelEx(&cfg, kern, p);
}

template <int KIND>
int launch_m(tb_decode_step* d, int B, cudaStream_t s) {
  if (B == 1) return launch_t<KIND, 1>(d, B, s);
  if (B == 2) return launch_t<KIND, 2>(d, B, s);
  return launch_t<KIND, 4>(d, B, s);
}

}  // namespace

extern "C" {

int tb_decode_step_max_batch(void) { return ds::kMaxRows; }

size_t tb_decode_step_tp_bytes(const tb_decode_step_config* c) {
  if (!c || c->tp_size <= 1) return 0;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return 0;
  return ds::tp_layout(c->hidden, c->vocab, sms).total;
}
const float* tb_decode_step_tp_logits(const tb_decode_step* d) { return d ? d->tp_logits : nullptr; }

int tb_decode_step_create(tb_decode_step** out, const tb_decode_step_config* c, const tb_decode_step_layer* layers,
                          const tb_decode_step_buffers* b) {
  if (!out || !c || !layers || !b) return -1;
  if (c->kind < 0 || c->kind > 3 || c->layers < 1 || c->tp_size < 1 || c->tp_size > 8 || c->tp_rank < 0 ||
      c->tp_rank >= c->tp_size)
    return -2;
  if (c->tp_size > 1) {
    for (int r = 0; r < c->tp_size; ++r)
      if (!b->tp_peers[r]) return -2;
  }
  const int epc = ds::epc_of(c->kind);
  if (c->hidden % (8 * 32) || c->hidden % epc || (c->heads_local * ds::kDh) % epc || c->inter_local % epc ||
      c->inter_local % 8 || c->hidden > 12288 || c->inter_local > 12288 || c->heads_local * ds::kDh > 12288)
    return -3;
  int dev = 0, sms = 0, smem_max = 0, coop = 0;
  TB_CHECK_CUDA(cudaGetDevice(&dev));
  TB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TB_CHECK_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  TB_CHECK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return -4;
  auto* d = new tb_decode_step();
  d->grid = sms;
  d->smem_max = smem_max;
  d->max_batch = c->max_batch < ds::kMaxRows ? c->max_batch : ds::kMaxRows;
  const int kmax = c->hidden > c->inter_local ? c->hidden : c->inter_local;
  // attention scratch (aliases the activation staging area): q, k, v of the new token, reductions, scores, partial outputs
  const int kpi = ds::kCT / (c->int8_kv ? 8 : 16);
  const size_t attn = (size_t) (2 * ds::kDh + 2 * ds::kCW + ((c->max_seq_len + 1 + 3) & ~3) + kpi * ds::kDh) * 4;
  // partial sums of the largest phase: stages per CTA x rows
  auto nst = [&](int kind, int K, int n_out, int R) {
    const size_t rowbytes = (size_t) (K / ds::epc_of(kind)) * 16;
    const int nseg = (int) ((rowbytes + ds::kSeg - 1) / ds::kSeg);
    return (size_t) R * ((n_out + sms - 1) / sms) * nseg;
  };
  size_t max_nst = nst(c->kind, c->hidden, 3 * c->heads_local * ds::kDh, 1);
  max_nst = std::max(max_nst, nst(c->kind, c->heads_local * ds::kDh, c->hidden, 1));
  max_nst = std::max(max_nst, nst(c->kind, c->hidden, c->inter_local, 2));
  max_nst = std::max(max_nst, nst(c->kind, c->inter_local, c->hidden, 1));
  max_nst = std::max(max_nst, nst(0, c->hidden, c->vocab_local, 1));
  for (int mi = 0; mi < 3; ++mi) {
    const int MB = 1 << mi;
    int widest = std::max(std::max(3 * c->heads_local * ds::kDh, c->hidden), std::max(c->inter_local, c->vocab_local));
    if ((size_t) ((widest + sms - 1) / sms) * MB > (size_t) ds::kEpiItems * ds::kCT) {   // outputs per thread in the epilogue
      delete d;
      return -5;
    }
    if (max_nst > (size_t) ds::kPartFloats || (size_t) ((c->vocab_local + sms - 1) / sms) > (size_t) ds::kLgFloats) {
      delete d;
      return -5;
    }
    const int xf_model = c->kind == 3 ? ds::kXInt8 : ds::kXHalf;
    const int xf_lm = ds::kXHalf;
    size_t xs = (size_t) MB * ds::xrow_bytes(kmax, epc, ds::xbytes_of(xf_model));
    xs = std::max(xs, (size_t) MB * ds::xrow_bytes(c->hidden, 8, ds::xbytes_of(xf_lm)));
    xs = std::max(xs, attn);
    const uint32_t ring_off = (uint32_t) ((ds::off_xs(MB) + xs + 1023) & ~(size_t) 1023);
    // S must be a multiple of the consumer warps (a slot is then always read by the same warp, in order); every slot is
    // filled by one fixed producer lane, in order (48 slots: lanes 0..15 own two): 48, 32 or 16
    int stages = (int) (((size_t) smem_max - ring_off) / ds::kSeg);
    stages = stages >= 48 ? 48 : (stages >= 32 ? 32 : (stages >= 16 ? 16 : 0));
    if (stages == 0) { delete d; return -6; }
    d->stages[mi] = stages;
    d->ring_off[mi] = ring_off;
    d->smem[mi] = ring_off + (size_t) stages * ds::kSeg;
  }
  std::vector<ds::Layer> hl(c->layers);
  for (int i = 0; i < c->layers; ++i) {
    const tb_decode_step_layer& s = layers[i];
    ds::Layer& l = hl[i];
    l.w_qkv = (const uint8_t*) s.w_qkv; l.w_dense = (const uint8_t*) s.w_dense; l.w_fc = (const uint8_t*) s.w_fc_gate;
    l.w_proj = (const uint8_t*) s.w_proj; l.s_qkv = s.s_qkv; l.s_dense = s.s_dense; l.s_fc = s.s_fc_gate; l.s_proj = s.s_proj;
    l.ln_in = (const __half*) s.ln_in; l.ln_post = (const __half*) s.ln_post; l.kv = (uint8_t*) s.kv_cache;
    l.kv_oq = s.kv_orig_quant; l.kv_qo = s.kv_quant_orig;
    if (!l.w_qkv || !l.w_dense || !l.w_fc || !l.w_proj || !l.ln_in || !l.ln_post || !l.kv) { delete d; return -7; }
    if (c->kind != 0 && (!l.s_qkv || !l.s_dense || !l.s_fc || !l.s_proj)) { delete d; return -7; }
    if (c->int8_kv && (!l.kv_oq || !l.kv_qo)) { delete d; return -7; }
  }
  cudaError_t e = cudaMalloc(&d->d_layers, sizeof(ds::Layer) * c->layers);
  if (e == cudaSuccess) e = cudaMemcpy(d->d_layers, hl.data(), sizeof(ds::Layer) * c->layers, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&d->d_cand_v, sizeof(float) * sms * ds::kMaxRows);
  if (e == cudaSuccess) e = cudaMalloc(&d->d_cand_i, sizeof(int) * sms * ds::kMaxRows);
  if (e == cudaSuccess) e = cudaMalloc(&d->d_bar, 128);
  if (e == cudaSuccess) e = cudaMemset(d->d_bar, 0, 128);
  if (e != cudaSuccess) { tb_decode_step_destroy(d); return (int) e; }
  ds::Params& p = d->p;
  p.layers = d->d_layers; p.n_layers = c->layers; p.kind = c->kind; p.hidden = c->hidden; p.hid_l = c->heads_local * ds::kDh;
  p.inter_l = c->inter_local; p.Hl = c->heads_local; p.vocab_l = c->vocab_local; p.vocab = c->vocab; p.S_max = c->max_seq_len;
  p.int8_kv = c->int8_kv; p.out_stride = c->out_stride; p.eps = c->rms_eps; p.inv_sqrt_dh = 1.f / sqrtf((float) ds::kDh);
  p.emb = (const __half*) b->emb; p.ln_f = (const __half*) b->ln_f; p.lm_head = (const uint8_t*) b->lm_head;
  p.hA = (__half*) b->h_a; p.hB = (__half*) b->h_b; p.qkv = (__half*) b->qkv; p.att = (__half*) b->att; p.act = (__half*) b->act;
  p.logits = b->logits; p.cand_v = d->d_cand_v; p.cand_i = d->d_cand_i;
  p.ids = b->ids; p.seq_lens = b->seq_lens; p.step_pos = b->step_pos; p.out_ids = b->out_ids; p.next = b->next_ids;
  p.in_lens = b->in_lens; p.max_in = b->max_in; p.bar = d->d_bar;
  p.tp = c->tp_size; p.rank = c->tp_rank;
  for (int r = 0; r < 8; ++r) p.peer[r] = r < c->tp_size && c->tp_size > 1 ? static_cast<uint8_t*>(b->tp_peers[r]) : nullptr;
  d->tp_logits = c->tp_size > 1
                     ? reinterpret_cast<const float*>(p.peer[c->tp_rank] + ds::tp_layout(c->hidden, c->vocab, sms).logits)
                     : nullptr;
  if (!p.emb || !p.ln_f || !p.lm_head || !p.hA || !p.hB || !p.qkv || !p.att || !p.act || !p.logits || !p.ids || !p.seq_lens ||
      !p.step_pos || !p.out_ids || !p.next || !p.in_lens || !p.max_in) {
    tb_decode_step_destroy(d);
    return -8;
  }
  *out = d;
  return 0;
}

void tb_decode_step_destroy(tb_decode_step* d) {
  if (!d) return;
  if (d->d_layers) cudaFree(d->d_layers);
  if (d->d_cand_v) cudaFree(d->d_cand_v);
  if (d->d_cand_i) cudaFree(d->d_cand_i);
  if (d->d_bar) cudaFree(d->d_bar);
  if (d->d_trace) cudaFree(d->d_trace);
  delete d;
}

int tb_decode_step_launch(tb_decode_step* d, int batch, cudaStream_t stream) {
  if (!d || batch < 1 || batch > d->max_batch) return -1;
  switch (d->p.kind) {
    case 0: return launch_m<ds::kF16>(d, batch, stream);
    case 1: return launch_m<ds::kW8>(d, batch, stream);
    case 2: return launch_m<ds::kW4>(d, batch, stream);
    default: return launch_m<ds::kA8W8>(d, batch, stream);
  }
}

/* diagnostics: the next launches record globaltimer stamps of every CTA's consumer pipeline (per projection: after the grid
 * barrier, after activation staging, after the last stage, after the epilogue; per attention phase: after the barrier,
 * after the items); out (host, grid x 2048 u64) receives them.  enable = 0 stops recording. */
int tb_decode_step_trace(tb_decode_step* d, int enable, unsigned long long* out_host) {
  if (!d) return -1;
  const size_t bytes = (size_t) d->grid * ds::kTraceSlots * sizeof(unsigned long long);
  if (enable && !d->d_trace) {
    TB_CHECK_CUDA(cudaMalloc(&d->d_trace, bytes));
    TB_CHECK_CUDA(cudaMemset(d->d_trace, 0, bytes));
  }
  if (out_host && d->d_trace) {
    TB_CHECK_CUDA(cudaDeviceSynchronize());
    TB_CHECK_CUDA(cudaMemcpy(out_host, d->d_trace, bytes, cudaMemcpyDeviceToHost));
  }
  d->p.trace = enable ? d->d_trace : nullptr;
  return 0;
}

int tb_decode_step_info(const tb_decode_step* d, int batch, int* stages, size_t* smem_bytes, int* grid) {
  if (!d || batch < 1 || batch > ds::kMaxRows) return -1;
  const int mi = batch == 1 ? 0 : (batch == 2 ? 1 : 2);
  if (stages) *stages = d->stages[mi];
  if (smem_bytes) *smem_bytes = d->smem[mi];
  if (grid) *grid = d->grid;
  return 0;
}
}
