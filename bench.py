#!/usr/bin/env python
"""Headline benchmark (driver contract): LLaMA-7B decode tokens/s on B200 through the plugin engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg3|cfg3_int8kv|cfg5]

Workload (BASELINE.json configs[1], the N=1 default): LLaMA-7B, fp16 weights, GPTAttention plugin + int8 KV cache,
batch 1, 128-token prompt, 128 generated tokens, synthetic prompt ids and seeded random-init weights.
A "step" is one whole request (context phase + 127 generation steps).  Metric = batch * out_len / latency, the
reference's own definition (T/benchmarks/gpt_benchmark.py:339; LQ/run.py:117-198 times setup+decode).
  value : requests driven with device-resident prompt ids, timed with CUDA events on the launching stream
  e2e   : GenerationSession.decode() with pinned HOST buffers (H2D prompt, D2H output ids inside the timed region)
  roofline : the weight-streaming GEMV (dominant kernel: > 90 % of a decode step) timed alone with CUDA events over
             all 32 layers' projections (13 GB of distinct weights, far beyond L2), algorithmic bytes = weight bytes
  cpu_baseline / --impl reference : the reference's run_hf.py path (HF fp32 generate on the host cores), bounded sample
N > 1 (torchrun): tensor parallel over N GPUs (strong scaling: one request sharded across ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LLAMA7B = dict(hidden=4096, heads=32, inter=11008, layers=32, vocab=32000)
WORKLOADS = {
    # name: (mode, int8_kv, batch, in_len, out_len, description)
    "cfg2": ("fp16", True, 1, 128, 128, "LLaMA-7B fp16 GPTAttentionPlugin + int8 KV-cache, batch=1, 128-in/128-out"),
    "cfg3": ("w8", False, 8, 1920, 128, "LLaMA-7B weight-only int8, batch=8, 2048-ctx decode (1920-in/128-out), fp16 KV"),
    "cfg3_int8kv": ("w8", True, 8, 1920, 128, "LLaMA-7B weight-only int8 + int8 KV, batch=8, 2048-ctx decode"),
    "cfg5": ("w4", True, 1, 128, 128, "LLaMA-7B int4 weight-only + int8 KV-cache, batch=1, 128-in/128-out"),
    "w8_b1": ("w8", True, 1, 128, 128, "LLaMA-7B weight-only int8 + int8 KV-cache, batch=1, 128-in/128-out"),
    "sq": ("sq", True, 1, 128, 128, "LLaMA-7B SmoothQuant per-token/per-channel int8 + int8 KV, batch=1, 128-in/128-out"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


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


def ncu_traffic_per_launch():
    """dram__bytes_read + dram__bytes_write per GEMV launch from the committed `ncu --set full` capture
    (profiles/r01_gemv_full.txt): one launch of each of the four projection shapes of a layer, averaged."""
    import re
    try:
        txt = open(os.path.join(ROOT, "profiles", "r01_gemv_full.txt")).read()
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


def gemv_roofline(torch, sess_tensors, cfg, mode, hbm_peak, which, rows=1):
    """Time the dominant kernel class alone: one weight-streaming decode projection launch per projection of every
    layer (`rows` token rows), replayed as ONE CUDA graph (no host launch gaps), CUDA events on the launching stream;
    every launch reads weights no other launch touched (13 GB of distinct weights for fp16: far beyond L2)."""
    from trtllm_llama_b200 import ops
    kind = {"fp16": ops.KIND_F16, "w8": ops.KIND_W8, "w4": ops.KIND_W4, "sq": ops.KIND_A8W8}[mode]
    bpw = {"fp16": 2.0, "w8": 1.0, "w4": 0.5, "sq": 1.0}[mode]
    hid = cfg["hidden"]
    x16 = (torch.randn(rows, max(hid, cfg["inter"]), device="cuda") * 0.1).half()
    x8 = torch.randint(-127, 127, (rows, max(hid, cfg["inter"])), device="cuda", dtype=torch.int8)
    st = torch.ones(rows, 1, device="cuda", dtype=torch.float32)
    calls, bytes_total = [], 0
    for i in range(cfg["layers"]):
        for name in ("attention.qkv", "attention.dense", "mlp.fc_gate", "mlp.proj"):
            wt = sess_tensors[f"layers.{i}.{name}.weight"]
            sc = sess_tensors.get(f"layers.{i}.{name}.per_channel_scale")
            N = wt.shape[0]
            K = int(round(wt.numel() * wt.element_size() / bpw / N))
            calls.append((wt, sc, N, K, name == "mlp.fc_gate"))
            bytes_total += int(N * K * bpw)
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
    return {"bound": "hbm", "kernel": f"decode projection GEMV ({'gemv_mma_kernel' if rows > 4 or mode == 'w4' else 'gemv_kernel'}, "
                                      f"{rows} token row{'s' if rows > 1 else ''})",
            "achieved": round(achieved, 1), "peak": hbm_peak, "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
            "unit": "GB/s", "frac": round(achieved / hbm_peak, 4), "bytes_per_launch": bytes_total // n_launch,
            "us_per_launch": round(ms * 1e3 / n_launch, 2), "launches_timed": n_launch * reps,
            "traffic": ncu_traffic_per_launch() if (mode == "fp16" and rows == 1) else None,
            "timing": "one CUDA graph of the %d launches, replayed %d times" % (n_launch, reps) if graphed else "eager launches"}


def run_reference(args):
    """--impl reference: the reference's run_hf.py path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.hf_baseline import time_hf_cpu
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
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = sample()
        vals.append(B * out_len / (last[0] + out_len * last[1]))
    value = statistics.mean(vals)
    ms_per_step = 1e3 * B * out_len / value
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "decode_tokens_per_sec", "value": round(value, 3), "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": desc, "batch": B, "in_len": in_len, "out_len": out_len},
            "cpu_baseline": {"value": round(value, 3), "unit": "tokens/s", "cores": cores, "kind": "port",
                             "sample": f"run_hf.py path (HF LlamaForCausalLM.forward greedy, fp32, random-init 7B) on {cores} host "
                                       f"threads: {in_len}-token prefill + {new_tokens} greedy steps per step, rate extrapolated "
                                       f"to {out_len} new tokens as out/(t_prefill + out*t_step)"},
            "e2e": {"value": round(value, 3), "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(time.perf_counter() - t_all, 1)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--nccl-only", action="store_true", help="tensor parallel: use the NCCL AllReduce plugin on the decode path too")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 (sm_100a): the product path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import trtllm_llama_b200  # noqa: F401
    from trtllm_llama_b200 import runtime as rt
    from trtllm_llama_b200._lib import lib
    from trtllm_llama_b200.quantization import QuantMode
    assert lib.tb_check_device() == 0

    mode, int8_kv, B, in_len, out_len, desc = WORKLOADS[args.workload]
    tp = world
    if tp > 1:
        import ctypes as C
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            assert lib.tb_comm_unique_id(idbuf.data_ptr()) == 0
        idd = idbuf.cuda()
        dist.broadcast(idd, 0)
        idbuf = idd.cpu()
        group = (C.c_int32 * tp)(*range(tp))
        rc = lib.tb_comm_init(idbuf.data_ptr(), group, tp, rank)
        assert rc == 0, f"tb_comm_init failed: {rc}"

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
    sess = rt.GenerationSession(mc, tensors, use_cuda_graph=not args.no_graph)
    if tp > 1 and not args.nccl_only:
        sess.enable_peer_allreduce()
    sess.setup(B, in_len, out_len)

    g = torch.Generator().manual_seed(1234)
    host_ids = torch.randint(3, LLAMA7B["vocab"], (B, in_len), generator=g, dtype=torch.int32).pin_memory()
    host_lens = torch.full((B,), in_len, dtype=torch.int32).pin_memory()
    host_out = torch.empty((B, out_len), dtype=torch.int32).pin_memory()
    dev_ids, dev_lens = host_ids.cuda(), host_lens.cuda()

    def request_device():
        st = torch.cuda.current_stream().cuda_stream
        if lib.tbrt_context(sess._e, dev_ids.data_ptr(), dev_lens.data_ptr(), B, in_len, st):
            raise RuntimeError(lib.tbrt_last_error().decode())
        for _ in range(out_len - 1):
            if lib.tbrt_step(sess._e, st):
                raise RuntimeError(lib.tbrt_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs, CUDA events, max over ranks --------------------------------
    for _ in range(max(args.warmup, 3)):
        request_device()
    launches = 0
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        request_device()
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    # ---- decode-only step time (graph replays), for the step-level roofline ---------------------------
    lib.tbrt_context(sess._e, dev_ids.data_ptr(), dev_lens.data_ptr(), B, in_len, torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        lib.tbrt_step(sess._e, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    n_dec = out_len - 4
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for _ in range(n_dec):
        lib.tbrt_step(sess._e, torch.cuda.current_stream().cuda_stream)
    d1.record()
    torch.cuda.synchronize()
    step_ms = d0.elapsed_time(d1) / n_dec
    step_launches = int(sess.last_launches)
    # ---- e2e: public API with pinned host buffers ----------------------------------------------------------
    for _ in range(2):
        sess.decode(host_ids, host_lens, out=host_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sess.decode(host_ids, host_lens, out=host_out)
    launches = int(sess.last_launches)
    barrier()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop()

    t = torch.tensor([dev_ms, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()
    value = B * out_len * args.steps / (dev_ms * 1e-3)
    e2e_value = B * out_len * args.steps / (e2e_ms * 1e-3)

    pk, which = peaks()
    hbm = float(pk["hbm_gbs"])
    bpw = {"fp16": 2.0, "w8": 1.0, "w4": 0.5, "sq": 1.0}[mode]
    L_mid = in_len + out_len // 2
    step_bytes = (6476005376 * bpw + 262144000) / tp + 2 * 32 * B * L_mid * 4096 * (1 if int8_kv else 2) / tp
    roof = gemv_roofline(torch, tensors, LLAMA7B, mode, hbm, which, rows=min(B, 8)) if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    line = {"metric": "decode_tokens_per_sec", "value": round(value, 2), "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(dev_ms / args.steps, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": {"fp16": "fp16", "w8": "fp16 x int8",
            "w4": "fp16 x int4", "sq": "int8"}[mode], "data": "synthetic",
            "config": {"workload": desc, "batch": B, "in_len": in_len, "out_len": out_len, "parallelism": f"tp{tp}",
                       "l2": "inputs larger than L2: every step streams %.1f GB of weights (L2 = 126 MB)" % (step_bytes / 1e9),
                       "step_definition": "one request = context phase + out_len-1 generation steps (CUDA-graph replays)"},
            "e2e": {"value": round(e2e_value, 2), "unit": "tokens/s", "h2d_bytes_per_step": int(B * in_len * 4 + B * 4),
                    "d2h_bytes_per_step": int(B * out_len * 4)},
            "gpu_launches": launches * args.steps,
            "decode_step": {"ms": round(step_ms, 4), "tokens_per_sec": round(B / (step_ms * 1e-3), 1), "kernels": step_launches,
                            "algorithmic_bytes": int(step_bytes), "achieved_gbs": round(step_bytes / (step_ms * 1e-3) / 1e9, 1),
                            "frac_of_hbm_peak": round(step_bytes / (step_ms * 1e-3) / 1e9 / hbm, 4)},
            "roofline": roof, "clocks": clk}
    if not args.no_cpu_baseline and world == 1:
        try:
            from oracle.hf_baseline import time_hf_cpu
            r = time_hf_cpu(batch=B, in_len=min(in_len, 128), out_len=out_len, new_tokens=6,
                            **{k: LLAMA7B[k] for k in ("hidden", "inter", "layers", "heads", "vocab")})
            line["cpu_baseline"] = {"value": round(r["value"], 3), "unit": "tokens/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"], "t_prefill_s": round(r["t_prefill_s"], 2),
                                    "t_step_s": round(r["t_step_s"], 3)}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {type(ex).__name__}: {ex}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
