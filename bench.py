#!/usr/bin/env python
"""Headline benchmark (driver contract): LLaMA-7B decode tokens/s on B200 through the plugin engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--only-headline]

Headline workload (BASELINE.json configs[1], the N=1 default): LLaMA-7B, fp16 weights, GPTAttention plugin + int8 KV
cache, batch 1, 128-token prompt, 128 generated tokens, synthetic prompt ids and seeded random-init weights.
A "step" is one whole request (context phase + 127 generation steps).  Metric = batch * out_len / latency, the
reference's own definition (T/benchmarks/gpt_benchmark.py:339; LQ/run.py:117-198 times setup+decode).
  value : requests driven with device-resident prompt ids, timed with CUDA events on the launching stream
  e2e   : GenerationSession.decode() with pinned HOST buffers (H2D prompt, D2H output ids inside the timed region)
  roofline : the dominant kernel class of a decode step (the weight-streaming projections) timed alone with CUDA events
             over all 32 layers' projections (distinct weights, far beyond L2), algorithmic bytes = weight bytes
  workloads : the SAME measurements for the other configurations BASELINE.json's metric names — SmoothQuant + int8 KV
             (north_star target), weight-only int8 at batch 8 / 2048-token context (fp16 and int8 KV), int4 + int8 KV at
             batch 1 and 8, and the SmoothQuant prefill at batch 8 x 2048 (int8 tensor-core roofline against a tcgen05
             kind::i8 peak measured in the same run) — each with value, decode_step / prefill_ms, roofline.frac, clocks
  reference_kernels : the reference's own CUDA kernels (oracle/_ref/libref_cuda.so, recompiled for sm_100a) timed in
             the same CUDA-event harness beside this library's kernel for the same shape
  cpu_baseline / --impl reference : the reference's run_hf.py path (HF fp32 generate on the host cores), bounded sample
N > 1 (torchrun): tensor parallel over N GPUs (strong scaling: one request sharded across ranks); rank 0 also runs the
first request on a tp = 1 engine and asserts the generated ids agree ("tp_parity").
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LLAMA7B = dict(hidden=4096, heads=32, inter=11008, layers=32, vocab=32000)
GEMM_WEIGHT_ELEMS = 6476005376          # 32 layers x 202,375,168
LM_HEAD_BYTES = 262144000
WORKLOADS = {
    # name: (mode, int8_kv, batch, in_len, out_len, description)
    "cfg2": ("fp16", True, 1, 128, 128, "LLaMA-7B fp16 GPTAttentionPlugin + int8 KV-cache, batch=1, 128-in/128-out"),
    "cfg3": ("w8", False, 8, 1920, 128, "LLaMA-7B weight-only int8, batch=8, 2048-ctx decode (1920-in/128-out), fp16 KV"),
    "cfg3_int8kv": ("w8", True, 8, 1920, 128, "LLaMA-7B weight-only int8 + int8 KV, batch=8, 2048-ctx decode"),
    "cfg5": ("w4", True, 1, 128, 128, "LLaMA-7B int4 weight-only + int8 KV-cache, batch=1, 128-in/128-out"),
    "cfg5_b8": ("w4", True, 8, 128, 128, "LLaMA-7B int4 weight-only + int8 KV-cache, batch=8, 128-in/128-out"),
    "w8_b1": ("w8", True, 1, 128, 128, "LLaMA-7B weight-only int8 + int8 KV-cache, batch=1, 128-in/128-out"),
    "sq": ("sq", True, 1, 128, 128, "LLaMA-7B SmoothQuant per-token/per-channel int8 + int8 KV, batch=1, 128-in/128-out"),
    "sq_b8": ("sq", True, 8, 128, 128, "LLaMA-7B SmoothQuant per-token/per-channel int8 + int8 KV, batch=8, 128-in/128-out"),
}
SIDE_N1 = ["sq", "cfg3", "cfg3_int8kv", "cfg5", "cfg5_b8"]     # + cfg4_prefill, at N = 1
SIDE_TP = ["cfg5", "cfg5_b8"]                                            # BASELINE configs[4], at N > 1
BPW = {"fp16": 2.0, "w8": 1.0, "w4": 0.5, "sq": 1.0}
DTYPE = {"fp16": "fp16", "w8": "fp16 x int8", "w4": "fp16 x int4", "sq": "int8"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.p, self.idx = None, gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None
        return self

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def make_weights(torch, cfg, rank, tp, seed=0):
    """seeded random-init fp16 LLaMA weights (HF init: normal std 0.02), generated on the device, already sharded."""
    from trtllm_llama_b200 import runtime as rt
    g = torch.Generator(device="cuda").manual_seed(seed)
    n = lambda *s: (torch.randn(*s, generator=g, device="cuda", dtype=torch.float32) * 0.02).half()  # noqa: E731
    hid, inter, V = cfg["hidden"], cfg["inter"], cfg["vocab"]
    w = {"vocab_embedding": n(V, hid), "ln_f": torch.ones(hid, device="cuda", dtype=torch.float16), "lm_head": n(V, hid),
         "layers": []}
    for _ in range(cfg["layers"]):
        w["layers"].append({"input_layernorm": torch.ones(hid, device="cuda", dtype=torch.float16), "qkv": n(3 * hid, hid),
                            "dense": n(hid, hid), "post_layernorm": torch.ones(hid, device="cuda", dtype=torch.float16),
                            "gate": n(inter, hid), "up": n(inter, hid), "down": n(hid, inter)})
        if tp > 1:   # shard layer by layer to bound peak memory
            one = rt.shard_weights({"vocab_embedding": w["vocab_embedding"], "ln_f": w["ln_f"], "lm_head": w["lm_head"],
                                    "layers": [w["layers"][-1]]}, tp, rank, cfg["heads"])
            w["layers"][-1] = one["layers"][0]
    if tp > 1:
        w["lm_head"] = w["lm_head"].chunk(tp, dim=0)[rank].contiguous()
    return w


def ncu_traffic_per_launch(mode, rows, tp):
    """dram__bytes_read + dram__bytes_write per projection launch from the committed `ncu --set full` capture of THIS
    round's kernels (profiles/r02_gemv_<mode>_full.txt: one launch of each of the four projection shapes of a layer,
    averaged).  None when no capture of this exact configuration (mode, rows, tp = 1) is committed."""
    import re
    if tp != 1:
        return None
    path = os.path.join(ROOT, "profiles", f"r02_gemv_{mode}_m{rows}_full.txt")
    try:
        txt = open(path).read()
    except OSError:
        return None
    per_grid = {}
    for blk in txt.split("----")[1:]:
        rd = re.search(r"dram__bytes_read.sum = ([0-9.]+) (\w+)", blk)
        wr = re.search(r"dram__bytes_write.sum = ([0-9.]+) (\w+)", blk)
        if not rd or not wr:
            continue
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = float(rd.group(1)) * unit.get(rd.group(2), 1.0) + float(wr.group(1)) * unit.get(wr.group(2), 1.0)
        per_grid[round(float(rd.group(1)))] = tot          # de-duplicate repeated shapes by their read size
    if len(per_grid) < 4:
        return None
    return int(sum(per_grid.values()) / len(per_grid))


def gemv_roofline(torch, sess_tensors, cfg, mode, hbm_peak, which, rows=1, tp=1):
    """Time the dominant kernel class alone: one weight-streaming decode projection launch per projection of every
    layer (`rows` token rows), replayed as ONE CUDA graph (no host launch gaps), CUDA events on the launching stream;
    every launch reads weights no other launch touched (13 GB of distinct weights for fp16: far beyond L2)."""
    from trtllm_llama_b200 import ops
    kind = {"fp16": ops.KIND_F16, "w8": ops.KIND_W8, "w4": ops.KIND_W4, "sq": ops.KIND_A8W8}[mode]
    bpw = BPW[mode]
    hid = cfg["hidden"]
    calls, bytes_total = [], 0
    for i in range(cfg["layers"]):
        for name in ("attention.qkv", "attention.dense", "mlp.fc_gate", "mlp.proj"):
            wt = sess_tensors[f"layers.{i}.{name}.weight"]
            sc = sess_tensors.get(f"layers.{i}.{name}.per_channel_scale")
            N = wt.shape[0]
            K = int(round(wt.numel() * wt.element_size() / bpw / N))
            calls.append((wt, sc, N, K, name == "mlp.fc_gate"))
            bytes_total += int(N * K * bpw)
    not_gemv = sorted({c[3] for c in calls if rows > ops.lib.tb_gemv_max_rows(kind, c[3])})
    if not_gemv:
        return {"bound": "hbm", "kernel": f"projections with K in {not_gemv} at {rows} rows are outside the decode-shape GEMV (K not a whole "
                                          "number of its k-steps at this tensor-parallel shard size): the engine runs this step's "
                                          "projections on the tcgen05 GEMM (gemm_tc_kernel); no GEMV class to time",
                "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None}
    x16 = (torch.randn(rows, max(hid, cfg["inter"]), device="cuda") * 0.1).half()
    x8 = torch.randint(-127, 127, (rows, max(hid, cfg["inter"])), device="cuda", dtype=torch.int8)
    st = torch.ones(rows, 1, device="cuda", dtype=torch.float32)
    xs16 = {K: x16[:, :K].contiguous() for K in {c[3] for c in calls}}
    xs8 = {K: x8[:, :K].contiguous() for K in {c[3] for c in calls}}

    def run_all():
        for wt, sc, N, K, swiglu in calls:
            if mode == "sq":
                ops.gemv(kind, xs8[K], wt, sc=sc.view(1, -1), sr=st, swiglu=swiglu)
            elif mode == "fp16":
                ops.gemv(kind, xs16[K], wt, swiglu=swiglu)
            else:
                ops.gemv(kind, xs16[K], wt, w_scale=sc, swiglu=swiglu)
    for _ in range(3):
        run_all()
    torch.cuda.synchronize()
    graphed = True
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run_all()
        replay = g.replay
    except Exception:          # noqa: BLE001  (fall back to eager launches; noted in the JSON)
        graphed, replay = False, run_all
    replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n_launch = len(calls)
    achieved = bytes_total / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": f"decode projection GEMV ({'gemv_mma_kernel' if ops.lib.tb_gemv_on_tensor_cores(kind, rows, hid) else 'gemv_kernel'}, "
                                      f"{rows} token row{'s' if rows > 1 else ''})",
            "achieved": round(achieved, 1), "peak": hbm_peak, "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
            "unit": "GB/s", "frac": round(achieved / hbm_peak, 4), "bytes_per_launch": bytes_total // n_launch,
            "us_per_launch": round(ms * 1e3 / n_launch, 2), "launches_timed": n_launch * reps,
            "traffic": ncu_traffic_per_launch(mode, rows, tp),
            "timing": "one CUDA graph of the %d launches, replayed %d times" % (n_launch, reps) if graphed else "eager launches"}


# ------------------------------------------------------------------------------------------------
class Ctx:
    """process-wide state shared by the workloads of one bench run"""

    def __init__(self, args):
        import torch
        self.torch, self.args = torch, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        import trtllm_llama_b200  # noqa: F401
        from trtllm_llama_b200._lib import lib
        self.lib = lib
        assert lib.tb_check_device() == 0
        if self.world > 1:
            import ctypes as C
            idbuf = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                assert lib.tb_comm_unique_id(idbuf.data_ptr()) == 0
            idd = idbuf.cuda()
            self.dist.broadcast(idd, 0)
            idbuf = idd.cpu()
            group = (C.c_int32 * self.world)(*range(self.world))
            rc = lib.tb_comm_init(idbuf.data_ptr(), group, self.world, self.rank)
            assert rc == 0, f"tb_comm_init failed: {rc}"
        self.pk, self.which = peaks()
        self.hbm = float(self.pk["hbm_gbs"])

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = self.torch.tensor(vals, device="cuda", dtype=self.torch.float64)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()


def build_session(cx, mode, int8_kv, B, in_len, out_len, tp, rank, graph=True, peer_ar=True):
    from trtllm_llama_b200 import runtime as rt
    from trtllm_llama_b200.quantization import QuantMode
    torch = cx.torch
    qm = {"fp16": QuantMode(0), "w8": QuantMode.use_weight_only(False), "w4": QuantMode.use_weight_only(True),
          "sq": QuantMode.use_smooth_quant(True, True)}[mode]
    if int8_kv:
        qm |= QuantMode.INT8_KV_CACHE
    mc = rt.ModelConfig(vocab_size=LLAMA7B["vocab"], num_layers=LLAMA7B["layers"], num_heads=LLAMA7B["heads"],
                        hidden_size=LLAMA7B["hidden"], inter_size=LLAMA7B["inter"], quant_mode=qm, max_batch_size=B,
                        max_input_len=in_len, max_output_len=out_len, tp_size=tp, tp_rank=rank)
    w = make_weights(torch, LLAMA7B, rank, tp)
    tensors = rt.build_engine_tensors(w, mc)
    del w
    torch.cuda.empty_cache()
    sess = rt.GenerationSession(mc, tensors, use_cuda_graph=graph)
    if tp > 1 and peer_ar:
        sess.enable_peer_allreduce()
    sess.setup(B, in_len, out_len)
    return sess, tensors


def step_bytes_of(mode, int8_kv, B, in_len, out_len, tp):
    L_mid = in_len + out_len // 2
    return (GEMM_WEIGHT_ELEMS * BPW[mode] + LM_HEAD_BYTES) / tp + 2 * 32 * B * L_mid * 4096 * (1 if int8_kv else 2) / tp


def run_decode_workload(cx, name, steps, warmup, with_e2e=True, with_roofline=True, return_ids=False):
    """one decode workload on this process group: builds weights + engine, times requests; rank 0 gets the result dict"""
    torch, lib = cx.torch, cx.lib
    mode, int8_kv, B, in_len, out_len, desc = WORKLOADS[name]
    tp, rank = cx.world, cx.rank
    sess, tensors = build_session(cx, mode, int8_kv, B, in_len, out_len, tp, rank, graph=not cx.args.no_graph,
                                  peer_ar=not cx.args.nccl_only)
    g = torch.Generator().manual_seed(1234)
    host_ids = torch.randint(3, LLAMA7B["vocab"], (B, in_len), generator=g, dtype=torch.int32).pin_memory()
    host_lens = torch.full((B,), in_len, dtype=torch.int32).pin_memory()
    host_out = torch.empty((B, out_len), dtype=torch.int32).pin_memory()
    dev_ids, dev_lens = host_ids.cuda(), host_lens.cuda()
    st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731

    def request_device():
        if lib.tbrt_context(sess._e, dev_ids.data_ptr(), dev_lens.data_ptr(), B, in_len, st()):
            raise RuntimeError(lib.tbrt_last_error().decode())
        for _ in range(out_len - 1):
            if lib.tbrt_step(sess._e, st()):
                raise RuntimeError(lib.tbrt_last_error().decode())

    # ---- value: device-resident inputs, CUDA events, max over ranks --------------------------------
    for _ in range(max(warmup, 3)):
        request_device()
    clocks = ClockSampler(cx.local)
    cx.barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        request_device()
    e1.record()
    cx.barrier()
    dev_ms = e0.elapsed_time(e1)
    # ---- context phase alone and decode-only step time (graph replays), for the step-level roofline ------------------
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    lib.tbrt_context(sess._e, dev_ids.data_ptr(), dev_lens.data_ptr(), B, in_len, st())
    c1.record()
    for _ in range(3):
        lib.tbrt_step(sess._e, st())
    torch.cuda.synchronize()
    ctx_ms = c0.elapsed_time(c1)
    n_dec = out_len - 4
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for _ in range(n_dec):
        lib.tbrt_step(sess._e, st())
    d1.record()
    torch.cuda.synchronize()
    step_ms = d0.elapsed_time(d1) / n_dec
    step_launches = int(sess.last_launches)
    # ---- the other decode path on the same engine (fused persistent step kernel vs per-operator plugin schedule) -------
    ab = None
    if B <= sess.fused_step_max_batch:
        fused_default = step_launches == 1
        sess.set_decode_mode(not fused_default)
        lib.tbrt_context(sess._e, dev_ids.data_ptr(), dev_lens.data_ptr(), B, in_len, st())
        for _ in range(3):
            lib.tbrt_step(sess._e, st())
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(n_dec):
            lib.tbrt_step(sess._e, st())
        a1.record()
        torch.cuda.synchronize()
        other_ms = a0.elapsed_time(a1) / n_dec
        other_launches = int(sess.last_launches)
        sess.set_decode_mode(None)
        other_ms = cx.max_over_ranks([other_ms])[0]
        ab = {"default_path": "fused persistent step kernel" if fused_default else "per-operator plugin schedule (CUDA graph)",
              "fused_step_ms": round(step_ms if fused_default else other_ms, 4),
              "plugin_schedule_ms": round(other_ms if fused_default else step_ms, 4),
              "fused_kernels": 1, "plugin_kernels": other_launches if fused_default else step_launches}
    # ---- e2e: public API with pinned host buffers ----------------------------------------------------------
    e2e_ms, launches = 0.0, 0
    if with_e2e:
        for _ in range(2):
            sess.decode(host_ids, host_lens, out=host_out)
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            sess.decode(host_ids, host_lens, out=host_out)
        launches = int(sess.last_launches)
        cx.barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
    else:
        sess.decode(host_ids, host_lens, out=host_out)
        launches = int(sess.last_launches)
    clk = clocks.stop()
    dev_ms, e2e_ms, step_ms, ctx_ms = cx.max_over_ranks([dev_ms, e2e_ms, step_ms, ctx_ms])
    ids = host_out.clone().numpy() if return_ids else None

    step_bytes = step_bytes_of(mode, int8_kv, B, in_len, out_len, tp)
    roof = None
    if with_roofline and rank == 0:
        # rank 0 only and no collective inside: a failure here must not leave this function on one rank alone (the other
        # ranks are already on their way to the next collective)
        try:
            roof = gemv_roofline(torch, tensors, LLAMA7B, mode, cx.hbm, cx.which, rows=min(B, 8), tp=tp)
        except Exception as ex:  # noqa: BLE001
            roof = {"error": f"{type(ex).__name__}: {ex}"}
    res = None
    if rank == 0:
        hbm_ms = step_bytes / (cx.hbm * 1e9) * 1e3
        res = {"workload": desc, "value": round(B * out_len * steps / (dev_ms * 1e-3), 2), "unit": "tokens/s",
               "ms_per_request": round(dev_ms / steps, 3), "steps": steps, "warmup": max(warmup, 3),
               "dtype": DTYPE[mode], "batch": B, "in_len": in_len, "out_len": out_len, "parallelism": f"tp{tp}",
               "context_ms": round(ctx_ms, 3),
               "decode_step": {"ms": round(step_ms, 4), "tokens_per_sec": round(B / (step_ms * 1e-3), 1),
                               "kernels": step_launches, "algorithmic_bytes": int(step_bytes),
                               "achieved_gbs": round(step_bytes / (step_ms * 1e-3) / 1e9, 1),
                               "frac_of_hbm_peak": round(step_bytes / (step_ms * 1e-3) / 1e9 / cx.hbm, 4),
                               "hbm_floor_ms": round(hbm_ms, 4)},
               "roofline": roof, "clocks": clk, "gpu_launches": launches * (steps if with_e2e else 1)}
        if ab:
            res["decode_paths"] = ab
        if with_e2e:
            res["e2e"] = {"value": round(B * out_len * steps / (e2e_ms * 1e-3), 2), "unit": "tokens/s",
                          "h2d_bytes_per_step": int(B * in_len * 4 + B * 4), "d2h_bytes_per_step": int(B * out_len * 4)}
        if tp > 1:
            # which term bounds a step at this tp: HBM (per-rank bytes), NVLink (64 one-shot all-reduces: every rank reads
            # world x M x hidden x 2 bytes from its peers at the measured 770 GB/s), or the dependent-kernel chain
            nvl_ms = 64 * (tp - 1) * B * 4096 * 2 / 770e9 * 1e3
            chain_ms = step_ms - max(hbm_ms, nvl_ms)
            res["limits"] = {"hbm_ms": round(hbm_ms, 4), "nvlink_ms": round(nvl_ms, 4),
                             "dependent_kernel_chain_ms": round(chain_ms, 4), "kernels": step_launches,
                             "us_per_kernel": round(step_ms * 1e3 / max(step_launches, 1), 2),
                             "limiter": "dependent-kernel chain (launch + sync latency)" if chain_ms > max(hbm_ms, nvl_ms)
                             else ("hbm" if hbm_ms >= nvl_ms else "nvlink")}
    del sess, tensors
    gc.collect()
    torch.cuda.empty_cache()
    return res, ids


def tp_parity(cx, name, tp_ids):
    """rank 0: the same request on a tp = 1 engine of the same weights; ids must agree up to the first step at which the
    tp = 1 top-2 logit margin is below the logit tolerance (a near-tie may legitimately break differently: the
    all-reduce sums partials in a different order than one GEMV does)."""
    torch = cx.torch
    import numpy as np
    mode, int8_kv, B, in_len, out_len, _ = WORKLOADS[name]
    sess, tensors = build_session(cx, mode, int8_kv, B, in_len, out_len, 1, 0, graph=True, peer_ar=False)
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(3, LLAMA7B["vocab"], (B, in_len), generator=g, dtype=torch.int32)
    lens = torch.full((B,), in_len, dtype=torch.int32)
    margins, toks = [], []
    lg = sess.context(ids, lens)
    for s in range(out_len):
        top2 = torch.topk(lg, 2, dim=-1).values
        margins.append((top2[:, 0] - top2[:, 1]).cpu().numpy())
        toks.append(lg.argmax(-1).cpu().numpy())
        if s + 1 < out_len:
            lg = sess.step()
    scale = float(lg.abs().max())
    one = sess.output_ids(out_len).cpu().numpy()
    del sess, tensors
    gc.collect()
    torch.cuda.empty_cache()
    tol = 2e-2 * max(1.0, scale)
    agree, first_diff, ok = 0, None, True
    for s in range(out_len):
        if np.array_equal(one[:, s], tp_ids[:, s]):
            agree += 1
            continue
        first_diff = s
        rows = one[:, s] != tp_ids[:, s]
        ok = bool((np.stack(margins, 1)[rows, s] <= tol).all())
        break
    return {"ok": ok, "ids_equal_steps": agree, "of": out_len, "first_difference_step": first_diff,
            "rule": f"ids identical until a step whose tp=1 top-2 margin <= {tol:.3g} (2e-2 x |logits|max)"}


def tp_parity_forced(cx, name, n_steps=32):
    """COLLECTIVE (every rank): the tensor-parallel engine and — on rank 0 — a tp = 1 engine of the same weights are stepped
    along the SAME token path (tbrt_force_ids: the tp = 1 arg-max is broadcast and forced into both), and rank 0 compares
    their fp32 logits after the context phase and after every one of ``n_steps`` generation steps.  Unlike the id
    comparison this does not end at the first near-tie, so the sharding, both all-reduces per layer, the vocabulary-parallel
    head and the decode path the engine selects under tensor parallelism (the fused step kernel) are checked on every
    step.  Tolerance: 2e-2 x max(1, |logits|max), the engine-vs-oracle logit tolerance of tests/test_engine_gpu.py."""
    torch, dist = cx.torch, cx.dist
    mode, int8_kv, B, in_len, out_len, _ = WORKLOADS[name]
    n_steps = min(n_steps, out_len - 1)
    sess, tensors = build_session(cx, mode, int8_kv, B, in_len, out_len, cx.world, cx.rank, graph=not cx.args.no_graph,
                                  peer_ar=not cx.args.nccl_only)
    # Every rank runs the SAME sequence of collectives whatever happens to rank 0's private tp = 1 engine: a failure there is
    # recorded (the check then reports ok = false) and the tensor-parallel engine is stepped along its own arg-max path.
    one = one_t = None
    one_error = None
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(3, LLAMA7B["vocab"], (B, in_len), generator=g, dtype=torch.int32)
    lens = torch.full((B,), in_len, dtype=torch.int32)
    lg_1 = None
    if cx.rank == 0:
        try:
            one, one_t = build_session(cx, mode, int8_kv, B, in_len, out_len, 1, 0, graph=True, peer_ar=False)
            lg_1 = one.context(ids, lens)
        except Exception as ex:  # noqa: BLE001
            one, lg_1, one_error = None, None, f"{type(ex).__name__}: {ex}"
    lg_tp = sess.context(ids, lens)
    worst, scale, argmax_equal = 0.0, 1.0, 0
    for s in range(n_steps + 1):
        tok = torch.zeros(B, dtype=torch.int32, device="cuda")
        if cx.rank == 0:
            if lg_1 is not None:
                worst = max(worst, float((lg_tp - lg_1).abs().max()))
                scale = max(scale, float(lg_1.abs().max()))
                argmax_equal += int(torch.equal(lg_tp.argmax(-1), lg_1.argmax(-1)))
                tok = lg_1.argmax(-1).to(torch.int32)
            else:
                tok = lg_tp.argmax(-1).to(torch.int32)
        dist.broadcast(tok, 0)
        if s == n_steps:
            break
        sess.force_ids(tok)
        lg_tp = sess.step()
        if one is not None:
            try:
                one.force_ids(tok)
                lg_1 = one.step()
            except Exception as ex:  # noqa: BLE001
                one, lg_1, one_error = None, None, f"{type(ex).__name__}: {ex}"
    path = "fused step kernel" if cx.lib.tbrt_last_launches(sess._e) == 1 else "per-operator plugin schedule"
    del sess, tensors, one, one_t
    gc.collect()
    torch.cuda.empty_cache()
    tol = 2e-2 * scale
    if one_error is not None:
        return {"ok": False, "error": "tp = 1 engine on rank 0: " + one_error, "decode_path": path}
    return {"ok": worst <= tol, "steps_compared": n_steps + 1, "max_abs_logit_diff": round(worst, 5), "tolerance": round(tol, 5),
"argmax_equal_steps": argmax_equal, "decode_path": path, "rule": "teacher-forced: both engines follow the tp=1 arg-max path; fp32 logits "
            "compared after the context phase and after every generation step"}


def int8_peak(cx):
    """tcgen05 kind::i8 tensor-pipe peak of this GPU at its sustained clock: back-to-back MMAs on resident operands,
    one CTA per SM, CUDA events (csrc/peak.cu).  Also kind::f16 the same way, for calibration against cuBLAS bf16."""
    import ctypes as C
    torch, lib = cx.torch, cx.lib
    sink = torch.zeros(4, dtype=torch.int32, device="cuda")
    out = {}
    for kind, key in ((0, "int8_tops"), (1, "f16_tflops")):
        ops = C.c_double(0)
        best = None
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.tb_mma_peak(kind, 20000, 148, sink.data_ptr(), C.byref(ops), torch.cuda.current_stream().cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0, f"tb_mma_peak failed: {rc}"
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        out[key] = round(ops.value / (best * 1e-3) / 1e12, 1)
    out["how"] = ("148 CTAs x 20000 x 4 back-to-back tcgen05.mma 128x256x(32 int8 | 16 fp16) on resident smem tiles, "
                  "best of 6, CUDA events (csrc/peak.cu)")
    return out


def run_prefill_workload(cx, steps, warmup):
    """BASELINE configs[3]: SmoothQuant per-token int8, batch 8, prefill 2048 (M = 16384 rows) through tbrt_context."""
    torch, lib = cx.torch, cx.lib
    B, S = 8, 2048
    sess, tensors = build_session(cx, "sq", True, B, S, 8, 1, 0)
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(3, LLAMA7B["vocab"], (B, S), generator=g, dtype=torch.int32).cuda()
    lens = torch.full((B,), S, dtype=torch.int32).cuda()
    st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    for _ in range(max(warmup, 3)):
        lib.tbrt_context(sess._e, ids.data_ptr(), lens.data_ptr(), B, S, st())
    torch.cuda.synchronize()
    clocks = ClockSampler(cx.local).start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        lib.tbrt_context(sess._e, ids.data_ptr(), lens.data_ptr(), B, S, st())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = int(sess.last_launches)
    # the dominant kernel class alone: the four projection GEMMs of every layer (distinct weights), one pass
    from trtllm_llama_b200 import ops
    M = B * S
    xq = torch.randint(-127, 128, (M, LLAMA7B["inter"]), device="cuda", dtype=torch.int8)
    sr = torch.rand(M, 1, device="cuda") * 0.01 + 1e-3
    calls = []
    for i in range(LLAMA7B["layers"]):
        for nm in ("attention.qkv", "attention.dense", "mlp.fc_gate", "mlp.proj"):
            wt = tensors[f"layers.{i}.{nm}.weight"]
            calls.append((wt, tensors[f"layers.{i}.{nm}.per_channel_scale"].view(1, -1), wt.shape[1]))
    xs = {K: xq[:, :K].contiguous() for K in {c[2] for c in calls}}
    def run_gemms(n):
        for wt, sc, K in calls[:n]:
            ops.gemm_tc(ops.KIND_A8W8, xs[K], wt, sc=sc, sr=sr)
    run_gemms(8)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    run_gemms(len(calls))
    g1.record()
    torch.cuda.synchronize()
    gemm_ms = g0.elapsed_time(g1)
    clk = clocks.stop()
    pk = int8_peak(cx)
    gemm_ops = 2.0 * M * GEMM_WEIGHT_ELEMS
    attn_flops = 2.0 * S * S * 4096 * 32 * B            # causal QK^T + PV, fp16
    tops = gemm_ops / (gemm_ms * 1e-3) / 1e12
    res = {"workload": "LLaMA-7B SmoothQuant per-token int8 (sq GEMM + RmsnormQuant), batch=8, prefill 2048",
           "prefill_ms": round(ms, 3), "value": round(B * S / (ms * 1e-3), 1), "unit": "prefill tokens/s", "steps": steps,
           "warmup": max(warmup, 3), "kernels": launches, "dtype": "int8",
           "int8_ops": gemm_ops, "attention_fp16_flops": attn_flops,
           "whole_context_int8_tops": round(gemm_ops / (ms * 1e-3) / 1e12, 1),
           "roofline": {"bound": "tensor", "kernel": "SmoothQuant projection GEMMs at M = 16384 (gemm_tc2_kernel, tcgen05 "
                                                     "kind::i8 cta_group::2), all 128 launches of the model, timed alone",
                        "achieved": round(tops, 1), "peak": pk["int8_tops"], "unit": "TOP/s",
                        "frac": round(tops / pk["int8_tops"], 4),
                        "peak_source": "measured in this run: " + pk["how"],
                        "frac_of_2x_bf16_sustained": round(tops / (2 * float(cx.pk.get("bf16_tflops_sustained", 1400.0))), 4),
                        "frac_of_datasheet_4500": round(tops / 4500.0, 4), "ms": round(gemm_ms, 3), "traffic": None},
           "roofline_floor_ms": {"gemm_at_measured_int8_peak": round(gemm_ops / (pk["int8_tops"] * 1e12) * 1e3, 2),
                                 "attention_at_bf16_sustained": round(attn_flops / (float(cx.pk.get("bf16_tflops_sustained", 1400.0)) * 1e12) * 1e3, 2)},
           "measured_peaks": pk, "clocks": clk}
    del sess, tensors, xq, xs
    gc.collect()
    torch.cuda.empty_cache()
    return res


def run_reference(args):
    """--impl reference: the reference's run_hf.py path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mode, int8_kv, B, in_len, out_len, desc = WORKLOADS[args.workload]
    vals, last = [], None
    t_all = time.perf_counter()
    # the model build dominates; build once, time `steps` bounded samples (warm-up samples included in W)
    import torch
    from oracle import hf_baseline as hb
    torch.set_num_threads(os.cpu_count() or 1)
    model = hb.build_hf_llama(**{k: LLAMA7B[k] for k in ("hidden", "inter", "layers", "heads", "vocab")})
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(3, LLAMA7B["vocab"], (B, in_len), generator=g)
    new_tokens = 4

    def sample():
        with torch.no_grad():
            t0 = time.perf_counter()
            out = model(input_ids=ids, use_cache=True)
            tp = time.perf_counter() - t0
            past, tok = out.past_key_values, out.logits[:, -1].argmax(-1, keepdim=True)
            t1 = time.perf_counter()
            for _ in range(new_tokens):
                out = model(input_ids=tok, past_key_values=past, use_cache=True)
                past, tok = out.past_key_values, out.logits[:, -1].argmax(-1, keepdim=True)
            ts = (time.perf_counter() - t1) / new_tokens
        return tp, ts
    for _ in range(max(1, min(args.warmup, 1))):
        sample()
    for _ in range(args.steps):
        last = sample()
        vals.append(B * out_len / (last[0] + out_len * last[1]))
    value = statistics.mean(vals)
    ms_per_step = 1e3 * B * out_len / value
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "decode_tokens_per_sec", "value": round(value, 3), "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            # the same `config` object as the b200 arm prints for this N (the driver compares the two)
            "config": {"workload": desc, "batch": B, "in_len": in_len, "out_len": out_len, "parallelism": f"tp{args.gpus}",
                       "l2": "inputs larger than L2: every step streams %.1f GB of weights (L2 = 126 MB)"
                             % (int(step_bytes_of(mode, int8_kv, B, in_len, out_len, args.gpus)) / 1e9),
                       "step_definition": "one request = context phase + out_len-1 generation steps (CUDA-graph replays)"},
            "cpu_baseline": {"value": round(value, 3), "unit": "tokens/s", "cores": cores, "kind": "port",
                             "sample": f"run_hf.py path (HF LlamaForCausalLM.forward greedy, fp32, random-init 7B) on {cores} host "
                                       f"threads: {in_len}-token prefill + {new_tokens} greedy steps per step, rate extrapolated "
                                       f"to {out_len} new tokens as out/(t_prefill + out*t_step)"},
            "e2e": {"value": round(value, 3), "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(time.perf_counter() - t_all, 1)}
    print(json.dumps(line), flush=True)


def _final_rendezvous(cx, seconds=120.0):
    """Last barrier + process-group teardown of a torchrun job.  The result line is already printed: if a peer never arrives
    (it died, or an earlier mismatch left it elsewhere) this process exits cleanly after `seconds` instead of hanging the job."""
    import threading
    watchdog = threading.Timer(seconds, lambda: os._exit(0))
    watchdog.daemon = True
    watchdog.start()
    cx.dist.barrier()
    cx.dist.destroy_process_group()
    watchdog.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--only-headline", action="store_true", help="skip the `workloads` / `reference_kernels` blocks")
    ap.add_argument("--side", default=None, help="comma list of side workloads (default: the BASELINE set for this N)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-tp-parity", action="store_true")
    ap.add_argument("--nccl-only", action="store_true", help="tensor parallel: use the NCCL AllReduce plugin on the decode path too")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 (sm_100a): the product path has no CPU fallback")
    cx = Ctx(args)
    rank, world = cx.rank, cx.world
    name = args.workload
    mode, int8_kv, B, in_len, out_len, desc = WORKLOADS[name]

    head, ids = run_decode_workload(cx, name, args.steps, args.warmup, return_ids=True)
    parity = None
    if world > 1 and not args.no_tp_parity:
        if rank == 0:
            try:
                parity = tp_parity(cx, name, ids)
            except Exception as ex:  # noqa: BLE001  (rank 0 only: must reach the barrier below like every other rank)
                parity = {"ok": False, "error": f"{type(ex).__name__}: {ex}"}
        cx.barrier()
        forced = tp_parity_forced(cx, name)
        cx.barrier()
        if rank == 0:
            parity["logits"] = forced
            parity["ok"] = bool(parity["ok"] and forced["ok"])

    side = {}
    if not args.only_headline:
        names = args.side.split(",") if args.side else (SIDE_N1 if world == 1 else SIDE_TP)
        for nm in [n for n in names if n and n != name]:
            if nm == "cfg4_prefill":
                continue
            try:
                r, _ = run_decode_workload(cx, nm, steps=max(3, min(args.steps, 5)), warmup=3, with_e2e=False)
            except Exception as ex:  # noqa: BLE001  (one side workload must not take the headline down)
                r = {"error": f"{type(ex).__name__}: {ex}"}
            if rank == 0:
                side[nm] = r
        if world == 1 and (args.side is None or "cfg4_prefill" in args.side):
            try:
                side["cfg4_prefill"] = run_prefill_workload(cx, steps=max(3, min(args.steps, 5)), warmup=3)
            except Exception as ex:  # noqa: BLE001
                side["cfg4_prefill"] = {"error": f"{type(ex).__name__}: {ex}"}

    if rank != 0:
        if world > 1:
            _final_rendezvous(cx)
        return
    # rank 0 alone from here to the final barrier: whatever happens while the line is assembled, the barrier is reached (the other
    # ranks are waiting in it) and a line is printed
    line = None
    try:
        line = {"metric": "decode_tokens_per_sec", "value": head["value"], "unit": "tokens/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_request"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE[mode], "data": "synthetic",
                "config": {"workload": desc, "batch": B, "in_len": in_len, "out_len": out_len, "parallelism": f"tp{world}",
                           "l2": "inputs larger than L2: every step streams %.1f GB of weights (L2 = 126 MB)"
                                 % (head["decode_step"]["algorithmic_bytes"] / 1e9),
                           "step_definition": "one request = context phase + out_len-1 generation steps (CUDA-graph replays)"},
                "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "decode_step": head["decode_step"],
                "context_ms": head["context_ms"], "roofline": head["roofline"], "clocks": head["clocks"]}
        if "decode_paths" in head:
            line["decode_paths"] = head["decode_paths"]
        if "limits" in head:
            line["limits"] = head["limits"]
        if parity is not None:
            line["tp_parity"] = parity["ok"]
            line["tp_parity_detail"] = parity
        if side:
            line["workloads"] = side
        if not args.only_headline and world == 1:
            try:
                from tools.ref_kernel_bench import reference_kernels
                line["reference_kernels"] = reference_kernels()
            except Exception as ex:  # noqa: BLE001
                line["reference_kernels"] = {"error": f"{type(ex).__name__}: {ex}"}
        if not args.no_cpu_baseline and world == 1:
            try:
                from oracle.hf_baseline import time_hf_cpu
                r = time_hf_cpu(batch=B, in_len=in_len, out_len=out_len, new_tokens=6,
                                **{k: LLAMA7B[k] for k in ("hidden", "inter", "layers", "heads", "vocab")})
                line["cpu_baseline"] = {"value": round(r["value"], 3), "unit": "tokens/s", "cores": r["cores"], "kind": "port",
                                        "sample": r["sample"], "in_len": in_len, "t_prefill_s": round(r["t_prefill_s"], 2),
                                        "t_step_s": round(r["t_step_s"], 3)}
            except Exception as ex:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {type(ex).__name__}: {ex}"}
    except Exception as ex:  # noqa: BLE001
        if line is None:
            line = {"metric": "decode_tokens_per_sec", "value": head["value"], "unit": "tokens/s", "n_gpus": world}
        line["assembly_error"] = f"{type(ex).__name__}: {ex}"
    print(json.dumps(line), flush=True)
    if world > 1:
        _final_rendezvous(cx)


if __name__ == "__main__":
    main()
