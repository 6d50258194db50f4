"""Python face of the C++ runtime (include/trtllm_b200_runtime.h) with the reference runtime's names.

Mirrors T/tensorrt_llm/runtime/generation.py: ``ModelConfig`` (:103-117), ``SamplingConfig`` (:119-138),
``GenerationSession.setup`` / ``.decode`` (:413-488, :782-997) for the contiguous-KV greedy path.  The
session holds a ``tbrt_engine``; torch supplies device memory and the stream, nothing else.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import torch

from ._lib import KernelError, TbrtConfig, lib
from .quantization import QuantMode, quantize_per_channel_int8, symmetric_quantize_last_axis_of_batched_matrix

MODE_FP16, MODE_W8, MODE_W4, MODE_SQ = 0, 1, 2, 3


@dataclass
class ModelConfig:
    """generation.py:103-117 (+ the sizes build.py bakes into config.json)."""
    vocab_size: int = 32000
    num_layers: int = 32
    num_heads: int = 32
    hidden_size: int = 4096
    inter_size: int = 11008
    gpt_attention_plugin: bool = True
    multi_query_mode: bool = False
    remove_input_padding: bool = False
    paged_kv_cache: bool = False
    remove_input_padding: bool = False  # LQ/build.py --remove_input_padding: the context phase runs on packed tokens
    tokens_per_block: int = 64          # paged KV cache block size (LQ/build.py --tokens_per_block; a power of two >= 16)
    rms_eps: float = 1e-6
    quant_mode: QuantMode = QuantMode(0)
    max_batch_size: int = 8
    max_input_len: int = 128
    max_output_len: int = 128
    tp_size: int = 1
    tp_rank: int = 0

    @property
    def mode(self) -> int:
        q = self.quant_mode
        if q.has_act_and_weight_quant():
            return MODE_SQ
        if q.is_int4_weight_only():
            return MODE_W4
        if q.is_int8_weight_only():
            return MODE_W8
        return MODE_FP16


@dataclass
class SamplingConfig:
    """generation.py:119-138.  top_k = 1 (default) is greedy; top_k > 1, top_p in (0, 1] and temperature select the
    sampling kernel (tb_sample: top-k, top-p, top-k + top-p), seeded by ``random_seed``.  num_beams > 1 is beam search
    (tbrt_beam_*; ``length_penalty`` normalises the beam scores as the reference's beam layer does).  Repetition penalty and
    min_length are not built and are rejected, not ignored."""
    end_id: int = 2
    pad_id: int = 2
    num_beams: int = 1
    temperature: float = 1.0
    top_k: int = 1
    top_p: float = 0.0
    length_penalty: float = 1.0
    repetition_penalty: float = 1.0
    min_length: int = 1
    random_seed: int = 0
    output_log_probs: bool = field(init=False, default=False)


def _err(what):
    return KernelError(f"{what}: {lib.tbrt_last_error().decode()}")


# ------------------------------------------------------------------------------------------------
# weights: fp16 torch-Linear tensors -> the engine's named, quantised, tensor-parallel-sharded tensors
# ------------------------------------------------------------------------------------------------
def shard_weights(w, tp_size, rank, num_heads):
    """Megatron split of one decoder's fp16 weights (T/examples/llama/weight.py:71-178):
    qkv rows per head group, dense / proj input dim, fc / gate rows, lm_head vocab rows."""
    if tp_size == 1:
        return w
    out = {k: w[k] for k in ("vocab_embedding", "ln_f")}
    out["lm_head"] = w["lm_head"].chunk(tp_size, dim=0)[rank].contiguous()
    out["layers"] = []
    for lw in w["layers"]:
        hidden = lw["dense"].shape[0]
        q, k, v = lw["qkv"].view(3, hidden, hidden).unbind(0)
        e = {"input_layernorm": lw["input_layernorm"], "post_layernorm": lw["post_layernorm"],
             "qkv": torch.cat([t.chunk(tp_size, dim=0)[rank] for t in (q, k, v)], dim=0).contiguous(),
             "dense": lw["dense"].chunk(tp_size, dim=1)[rank].contiguous(),
             "gate": lw["gate"].chunk(tp_size, dim=0)[rank].contiguous(),
             "up": lw["up"].chunk(tp_size, dim=0)[rank].contiguous(),
             "down": lw["down"].chunk(tp_size, dim=1)[rank].contiguous()}
        if "kv_scale" in lw:
            e["kv_scale"] = lw["kv_scale"]
        out["layers"].append(e)
    return out


def quantize_linear(w_nk: torch.Tensor, mode: int):
    """one Linear weight [N, K] fp16 -> {"weight": ..., "per_channel_scale": ...} in the plugin's layout."""
    if mode == MODE_FP16:
        return {"weight": w_nk.contiguous()}
    if mode in (MODE_W8, MODE_W4):
        qt = torch.int8 if mode == MODE_W8 else torch.quint4x2
        processed, scales = symmetric_quantize_last_axis_of_batched_matrix(w_nk.t(), qt)   # op takes [K, N]
        return {"weight": processed, "per_channel_scale": scales}
    q, s = quantize_per_channel_int8(w_nk)
    return {"weight": q, "per_channel_scale": s}


def build_engine_tensors(w, cfg: ModelConfig, kv_scale: float = 4.0 / 127.0, require_kv_scale: bool = False):
    """fp16 weights (this rank's shard, see ``shard_weights``) -> {engine tensor name: device tensor}.

    ``kv_scale`` is the int8-KV placeholder for random-init builds; a layer's own calibrated ``kv_scale`` wins.  With
    ``require_kv_scale`` (real checkpoints: ``--model_dir`` given) a layer without one raises, as the reference loader
    does on the missing ``scale_y_quant_orig`` file (LQ/weight_quant.py:439-446)."""
    mode = cfg.mode
    t = {"vocab_embedding.weight": w["vocab_embedding"], "ln_f.weight": w["ln_f"], "lm_head.weight": w["lm_head"]}
    for i, lw in enumerate(w["layers"]):
        p = f"layers.{i}."
        t[p + "input_layernorm.weight"] = lw["input_layernorm"]
        t[p + "post_layernorm.weight"] = lw["post_layernorm"]
        if mode == MODE_SQ and cfg.tp_size > 1:
            raise NotImplementedError("SmoothQuant with tensor parallelism: quantise before sharding (see DESIGN.md)")
        # fc = gate_proj, gate = up_proj (LQ/weight_quant.py:343-404); fused as one [2*inter, K] projection
        named = {"attention.qkv": lw["qkv"], "attention.dense": lw["dense"],
                 "mlp.fc_gate": torch.cat([lw["gate"], lw["up"]], dim=0), "mlp.proj": lw["down"]}
        pre_q = lw.get("sq") if mode == MODE_SQ else None   # converter-made int8 weights + per-channel scales
        for name, wt in named.items():
            if pre_q is not None:
                parts = {"attention.qkv": ["qkv"], "attention.dense": ["dense"], "mlp.fc_gate": ["gate", "up"],
                         "mlp.proj": ["down"]}[name]
                t[p + name + ".weight"] = torch.cat([pre_q[n][0] for n in parts], dim=0)
                t[p + name + ".per_channel_scale"] = torch.cat([pre_q[n][1] for n in parts], dim=0)
                continue
            for k, v in quantize_linear(wt, mode).items():
                t[p + name + "." + k] = v
        if cfg.quant_mode.has_int8_kv_cache():
            dev = lw["qkv"].device
            if require_kv_scale and "kv_scale" not in lw:
                raise FileNotFoundError(
                    f"layer {i}: int8 KV cache requested but the checkpoint has no calibrated "
                    "attention.query_key_value.scale_y_quant_orig (run hf_llama_convert.py with -kv or -sq)")
            layer_scale = float(lw.get("kv_scale", kv_scale))
            # LQ/weight_quant.py:439-446: kv_orig_quant_scale = 1/t, kv_quant_orig_scale = t
            t[p + "attention.kv_orig_quant_scale"] = torch.tensor([1.0 / layer_scale], dtype=torch.float32, device=dev)
            t[p + "attention.kv_quant_orig_scale"] = torch.tensor([layer_scale], dtype=torch.float32, device=dev)
    return {k: v.contiguous() for k, v in t.items()}


# ------------------------------------------------------------------------------------------------
class GenerationSession:
    """generation.py:151-242 GenerationSession, holding a tbrt_engine instead of a TensorRT context."""

    def __init__(self, model_config: ModelConfig, engine_tensors: dict, use_cuda_graph: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("GenerationSession needs a CUDA device (sm_100a); there is no CPU path")
        mc = self.cfg = model_config
        c = TbrtConfig(hidden=mc.hidden_size, heads=mc.num_heads, inter=mc.inter_size, layers=mc.num_layers,
                       vocab=mc.vocab_size, head_size=mc.hidden_size // mc.num_heads, rms_eps=mc.rms_eps, mode=mc.mode,
                       int8_kv=int(mc.quant_mode.has_int8_kv_cache()), max_batch=mc.max_batch_size,
                       max_input_len=mc.max_input_len, max_output_len=mc.max_output_len, tp_size=mc.tp_size,
                       tp_rank=mc.tp_rank, use_cuda_graph=int(use_cuda_graph),
                       paged_kv_tokens_per_block=int(mc.tokens_per_block) if mc.paged_kv_cache else 0)
        self._e = lib.tbrt_create(C.byref(c))
        if not self._e:
            raise _err("tbrt_create")
        self._tensors = engine_tensors      # keep the device memory alive
        for name, t in engine_tensors.items():
            if not t.is_cuda or not t.is_contiguous():
                raise ValueError(f"engine tensor {name} must be a contiguous CUDA tensor")
            if lib.tbrt_set_tensor(self._e, name.encode(), t.data_ptr(), t.numel() * t.element_size()):
                raise _err("tbrt_set_tensor")
        if lib.tbrt_finalize(self._e):
            raise _err("tbrt_finalize")
        self.batch_size = self.max_input_len = self.max_new_tokens = 0
        self.kv_cache_manager = None
        if mc.paged_kv_cache:
            per_seq = lib.tbrt_kv_max_blocks_per_seq(self._e)
            self.kv_cache_manager = KVCacheManager(blocks=mc.max_batch_size * per_seq, tokens_per_block=mc.tokens_per_block,
                                                   max_blocks_per_seq=per_seq)

    def __del__(self):
        e, self._e = getattr(self, "_e", None), None
        try:
            if e:
                lib.tbrt_destroy(e)
        except Exception:      # interpreter shutdown: the library wrapper may already be gone
            pass

    @property
    def device_bytes(self):
        return lib.tbrt_device_bytes(self._e)

    def set_decode_mode(self, fused):
        """True: generation steps run as one persistent kernel when the engine / batch allow it; False: the per-operator
        plugin schedule (one IPluginV2DynamicExt.enqueue per operator, CUDA graph); None (default): whichever measured
        faster on B200 (the persistent kernel under tensor parallelism, the plugin schedule on one GPU)."""
        lib.tbrt_set_decode_mode(self._e, -1 if fused is None else int(bool(fused)))

    @property
    def fused_step_max_batch(self):
        return lib.tbrt_fused_step_available(self._e)

    @property
    def last_launches(self):
        return lib.tbrt_last_launches(self._e)

    def enable_peer_allreduce(self, group=None):
        """Tensor parallel: exchange the IPC handles of the peer-memory all-reduce buffers through torch.distributed
        (one process per GPU) so the decode path uses the fused NVLink all-reduce + residual kernel instead of NCCL."""
        import torch.distributed as dist
        mine = torch.zeros(64, dtype=torch.uint8)
        if lib.tbrt_ar_handle(self._e, mine.data_ptr()):
            raise _err("tbrt_ar_handle")
        world = self.cfg.tp_size
        table = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(table, mine.cuda(), group=group)
        flat = torch.cat([t.cpu() for t in table]).contiguous()
        if lib.tbrt_ar_open(self._e, flat.data_ptr()):
            raise _err("tbrt_ar_open")
        dist.barrier(group=group)

    def setup(self, batch_size, max_input_length, max_new_tokens, beam_width=1):
        """generation.py:413-488: fixes the shapes of the next decode (buffers were sized at engine build)."""
        mc = self.cfg
        if not 1 <= beam_width <= 16:
            raise ValueError("beam_width must be in [1, 16]")
        if beam_width > 1 and self.kv_cache_manager is not None:
            raise NotImplementedError("beam search needs the contiguous KV cache (paged engines: beam width 1)")
        self.beam_width = beam_width
        if batch_size * beam_width > mc.max_batch_size or max_input_length > mc.max_input_len or max_new_tokens > mc.max_output_len:
            raise ValueError("setup() exceeds the limits the engine was built with")
        self.batch_size, self.max_input_len, self.max_new_tokens = batch_size, max_input_length, max_new_tokens

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    # granular entry points (tests compare logits step by step)
    def _upload_kv_blocks(self, batch):
        table = self.kv_cache_manager.get_block_table(batch)
        if lib.tbrt_set_kv_blocks(self._e, table.data_ptr(), table.shape[0], table.shape[1], self._stream()):
            raise _err("tbrt_set_kv_blocks")

    def context(self, input_ids: torch.Tensor, input_lengths: torch.Tensor):
        B, S = input_ids.shape
        if self.kv_cache_manager is not None:
            # generation.py:609-640: every sequence of the padded batch gets blocks for max_input_length + 1 positions
            m = self.kv_cache_manager
            m.reset()
            for b in range(B):
                m.add_sequence(GenerationSequence(seq_idx=b, batch_idx=b), S)
            self._upload_kv_blocks(B)
        ids = input_ids.to(device="cuda", dtype=torch.int32).contiguous()
        lens = input_lengths.to(device="cuda", dtype=torch.int32).contiguous()
        if self.cfg.remove_input_padding:
            # generation.py:355-363: with remove_input_padding the engine's input_ids is [1, num_tokens], prompts back to back
            keep = torch.arange(S, device="cuda")[None, :] < lens[:, None]
            packed = ids[keep].contiguous()
            if lib.tbrt_context_packed(self._e, packed.data_ptr(), lens.data_ptr(), B, packed.numel(), S, self._stream()):
                raise _err("tbrt_context_packed")
        elif lib.tbrt_context(self._e, ids.data_ptr(), lens.data_ptr(), B, S, self._stream()):
            raise _err("tbrt_context")
        self._B = B
        return self.logits()

    def step(self):
        if self.kv_cache_manager is not None:
            # generation.py:944-949: advance the manager; a sequence entering a new block gets one from the free list
            if self.kv_cache_manager.step([False] * self._B):
                self._upload_kv_blocks(self._B)
        if lib.tbrt_step(self._e, self._stream()):
            raise _err("tbrt_step")
        return self.logits()

    def force_ids(self, ids: torch.Tensor):
        """teacher forcing (parity checks): the token chosen by the last context() / step() becomes ``ids`` [B]."""
        t = ids.to(device="cuda", dtype=torch.int32).contiguous()
        if lib.tbrt_force_ids(self._e, t.data_ptr(), self._stream()):
            raise _err("tbrt_force_ids")

    def logits(self) -> torch.Tensor:
        """fp32 [B, vocab] copy of the engine's logits buffer."""
        n = self._B * self.cfg.vocab_size
        out = torch.empty((self._B, self.cfg.vocab_size), dtype=torch.float32, device="cuda")
        _memcpy_d2d(out.data_ptr(), lib.tbrt_logits(self._e), n * 4, self._stream())
        return out

    def output_ids(self, n_tokens) -> torch.Tensor:
        full = torch.empty((self._B, self.cfg.max_output_len), dtype=torch.int32, device="cuda")
        _memcpy_d2d(full.data_ptr(), lib.tbrt_output_ids(self._e), full.numel() * 4, self._stream())
        return full[:, :n_tokens]

    def kv_cache(self, layer) -> torch.Tensor:
        mc = self.cfg
        shape = (mc.max_batch_size, 2, mc.num_heads // mc.tp_size, mc.max_input_len + mc.max_output_len,
                 mc.hidden_size // mc.num_heads)
        dt = torch.int8 if mc.quant_mode.has_int8_kv_cache() else torch.float16
        out = torch.empty(shape, dtype=dt, device="cuda")
        _memcpy_d2d(out.data_ptr(), lib.tbrt_kv_cache(self._e, layer), out.numel() * out.element_size(), self._stream())
        return out

    def decode(self, input_ids: torch.Tensor, input_lengths: torch.Tensor, sampling_config: SamplingConfig = None,
               max_new_tokens: int = None, out: torch.Tensor = None) -> torch.Tensor:
        """generation.py:782-997 for greedy sampling.  ``input_ids`` [B, S] / ``input_lengths`` [B] are HOST int32
        tensors (pinned for asynchronous copies); returns HOST output ids [B, max_new_tokens] — the host<->device
        copies are part of the call, as in the reference's run.py timing (LQ/run.py:117-198)."""
        sc = sampling_config or SamplingConfig()
        if sc.repetition_penalty != 1.0 or sc.min_length > 1 or (sc.num_beams == 1 and sc.length_penalty != 1.0):
            raise NotImplementedError("repetition_penalty / min_length (and length_penalty without beams) are not built (SURVEY 8f-4)")
        if sc.num_beams != 1:
            return self._decode_beams(input_ids, input_lengths, sc, max_new_tokens or self.max_new_tokens, out)
        if lib.tbrt_set_sampling(self._e, int(sc.top_k), float(sc.top_p), float(sc.temperature), int(sc.random_seed)):
            raise _err("tbrt_set_sampling")
        B, S = input_ids.shape
        n = max_new_tokens or self.max_new_tokens
        if input_ids.is_cuda or input_ids.dtype != torch.int32 or not input_ids.is_contiguous():
            raise ValueError("decode() takes contiguous host int32 input_ids")
        if out is None:
            out = torch.empty((B, n), dtype=torch.int32, pin_memory=True)
        lens = input_lengths.to(dtype=torch.int32).contiguous()
        # with an explicit sampling config the engine applies the reference's stop criterion: it checks every 16 steps
        # whether every sequence has produced end_id, stops early, and pads finished sequences with end_id
        end_id = sampling_config.end_id if (sampling_config is not None and sampling_config.end_id is not None) else -1
        lib.tbrt_set_end_id(self._e, int(end_id))
        if self.kv_cache_manager is not None or self.cfg.remove_input_padding:
            # paged KV cache: the block tables change while the request runs, so the step loop is driven from the host as
            # in the reference (generation.py:852-997); the stop criterion is applied to the finished ids.  Packed input
            # takes the same route (the host packs the prompts)
            self.context(input_ids, lens)
            for _ in range(n - 1):
                self.step()
            ids = self.output_ids(n).cpu()
            if end_id >= 0:
                pad_finished(ids, end_id)
            out.copy_(ids)
            self.last_steps = n
            return out
        if lib.tbrt_generate(self._e, input_ids.data_ptr(), lens.data_ptr(), B, S, n, out.data_ptr(), self._stream()):
            raise _err("tbrt_generate")
        self._B = B
        self.last_steps = lib.tbrt_last_steps(self._e)
        return out

    def _decode_beams(self, input_ids, input_lengths, sc, n, out):
        """generation.py:365-409,823-997 with num_beams > 1: the context phase runs once per batch entry, the engine tiles the
        cache beam_width times and every later step advances all beams; returns HOST ids [B, num_beams, n], best beam first
        (``self.cum_log_probs`` [B, num_beams] holds the final scores)."""
        B, S = input_ids.shape
        W = int(sc.num_beams)
        if sc.top_k != 1 or sc.top_p != 0.0:
            raise NotImplementedError("beam search does not combine with top-k / top-p sampling (as in the reference's decoder)")
        if self.kv_cache_manager is not None:
            raise NotImplementedError("beam search needs the contiguous KV cache (paged engines: beam width 1)")
        if input_ids.is_cuda or input_ids.dtype != torch.int32 or not input_ids.is_contiguous():
            raise ValueError("decode() takes contiguous host int32 input_ids")
        if out is None:
            out = torch.empty((B, W, n), dtype=torch.int32, pin_memory=True)
        if lib.tbrt_set_sampling(self._e, 1, 0.0, 1.0, 0):
            raise _err("tbrt_set_sampling")
        self.context(input_ids, input_lengths.to(dtype=torch.int32))
        end_id = int(sc.end_id) if sc.end_id is not None else -1
        if lib.tbrt_beam_begin(self._e, W, float(sc.length_penalty), end_id, self._stream()):
            raise _err("tbrt_beam_begin")
        self._B = B * W
        for _ in range(n - 1):
            if lib.tbrt_step(self._e, self._stream()):
                raise _err("tbrt_step")
        cum = torch.empty((B, W), dtype=torch.float32, pin_memory=True)
        if lib.tbrt_beam_finalize(self._e, out.data_ptr(), cum.data_ptr(), n, self._stream()):
            raise _err("tbrt_beam_finalize")
        self.cum_log_probs = cum
        self.last_steps = n
        return out


@dataclass(eq=False)
class GenerationSequence:
    """T/tensorrt_llm/runtime/kv_cache_manager.py:31-52"""
    seq_idx: int
    batch_idx: int

    def get_batch_idx(self) -> int:
        return self.batch_idx

    def get_seq_idx(self) -> int:
        return self.seq_idx

    def __eq__(self, other):
        return isinstance(other, GenerationSequence) and (self.seq_idx, self.batch_idx) == (other.seq_idx, other.batch_idx)

    def __hash__(self):
        return self.seq_idx


class KVCacheManager:
    """Block bookkeeping of the paged KV cache — T/tensorrt_llm/runtime/kv_cache_manager.py:57-292 (BlocksManager +
    KVCacheManager) for beam width 1.  The pools live in the engine (one per layer, same block ids in every layer); this
    class only decides WHICH pool block holds which 2^k positions of which sequence:

      add_sequence(seq, context_len)   ceil((context_len + 1) / tokens_per_block) blocks (:262-277)
      step(finished)                   one more block for every unfinished sequence whose length is about to cross a block
                                       boundary, finished sequences return their blocks (:225-260); returns True when the
                                       table changed
      get_block_table(batch)           int32 [batch, max_blocks_per_seq], -1 = not allocated  (the engine turns ids into the
                                       [B, 1, 2, max_blocks] pointer array of get_pointer_arrays, :279-292)

    The free list is dealt in a seeded shuffled order so that a sequence's blocks are scattered over the pool and
    interleaved with other sequences' — the layout a long-running server converges to, and what the tests need to show
    that the kernels really follow the table."""

    def __init__(self, blocks: int, tokens_per_block: int, max_blocks_per_seq: int, seed: int = 0):
        if tokens_per_block < 16 or tokens_per_block & (tokens_per_block - 1):
            raise ValueError("tokens_per_block must be a power of two >= 16")
        self.blocks, self.tokens_per_block, self.max_blocks_per_seq, self.seed = blocks, tokens_per_block, max_blocks_per_seq, seed
        self.reset()

    def reset(self):
        g = torch.Generator().manual_seed(self.seed)
        self.free_blocks = torch.randperm(self.blocks, generator=g).tolist()
        self.allocated = {}          # GenerationSequence -> [block ids]
        self.sequences, self.lens = [], []

    def has_free_block(self) -> bool:
        return len(self.free_blocks) > 0

    def _allocate(self, seq):
        if not self.has_free_block():
            raise RuntimeError("Can't allocate new block for KV cache")
        if len(self.allocated[seq]) >= self.max_blocks_per_seq:
            raise RuntimeError("sequence exceeds max_blocks_per_seq")
        self.allocated[seq].append(self.free_blocks.pop(0))

    def add_sequence(self, sequence: GenerationSequence, context_len: int):
        self.sequences.append(sequence)
        self.lens.append(context_len)
        self.allocated[sequence] = []
        for _ in range(-(-(context_len + 1) // self.tokens_per_block)):     # one more token for the 1st generation step
            self._allocate(sequence)

    def step(self, finished) -> bool:
        changed = False
        for seq in self.sequences:
            b = seq.get_batch_idx()
            if not finished[b] and self.lens[b] % self.tokens_per_block == self.tokens_per_block - 1:
                self._allocate(seq)
                changed = True
            self.lens[b] += 1
        for b, f in enumerate(finished):
            if f:
                self.free_blocks.extend(self.allocated.pop(self.sequences[b]))
                changed = True
        keep = [(s, l) for s, l, f in zip(self.sequences, self.lens, finished) if not f]
        self.sequences, self.lens = [s for s, _ in keep], [l for _, l in keep]
        for i, s in enumerate(self.sequences):
            s.batch_idx = i
        return changed

    def get_block_table(self, batch: int) -> torch.Tensor:
        t = torch.full((batch, self.max_blocks_per_seq), -1, dtype=torch.int32)
        for seq, ids in self.allocated.items():
            if seq.get_batch_idx() < batch:
                t[seq.get_batch_idx(), :len(ids)] = torch.tensor(ids, dtype=torch.int32)
        return t


def pad_finished(output_ids: torch.Tensor, end_id: int) -> torch.Tensor:
    """In place on HOST ids [B, T]: once a sequence has produced ``end_id`` it is finished and every later position holds
    ``end_id`` — what the reference's decoder leaves in ``output_ids`` (finished sequences are skipped by its sampling
    kernels and keep emitting ``end_id``, generation.py:782-997).  This engine always runs all T steps (no early stop), so
    the positions after the first ``end_id`` are overwritten here; ids up to and including it are untouched."""
    hit = output_ids == end_id
    finished_before = (hit.cumsum(dim=1) - hit.to(hit.cumsum(dim=1).dtype)) > 0      # an end_id strictly earlier in the row
    output_ids[finished_before] = end_id
    return output_ids


def _memcpy_d2d(dst, src, nbytes, stream):
    rc = lib.tb_copy(dst, src, nbytes, stream)
    if rc:
        raise KernelError(f"tb_copy failed with {rc}")
