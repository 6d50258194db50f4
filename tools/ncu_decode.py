#!/usr/bin/env python
"""Profiling target for ncu: one LLaMA-7B engine (workload of bench.py), the context phase and a few generation steps;
the steps (or, with --prefill, one context phase) to be profiled run inside the NVTX range "profile", eagerly (no CUDA
graph), so `ncu --nvtx --nvtx-include "profile/"` sees exactly those launches.

    ncu --nvtx --nvtx-include "profile/" --metrics gpu__time_duration.sum --clock-control none --csv \\
        --log-file gpurun_out/r02_cfg2_launches.csv python tools/ncu_decode.py --workload cfg2 --steps 2
    ncu --nvtx --nvtx-include "profile/" --set full --clock-control none --import-source on -k regex:gemv_kernel -c 5 \\
        -o gpurun_out/r02_gemv_fp16 python tools/ncu_decode.py --workload cfg2 --steps 1
    ... --fused 1 -k regex:decode_step -c 1       # the persistent whole-step kernel
    ... --prefill --layers 2                      # BASELINE configs[3]: SmoothQuant prefill, batch 8 x 2048"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--fused", type=int, default=0)
    ap.add_argument("--prefill", action="store_true")
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--no-graph", action="store_true", default=True)
    ap.add_argument("--nccl-only", action="store_true")
    args = ap.parse_args()
    import torch
    import bench
    bench.LLAMA7B = dict(bench.LLAMA7B, layers=args.layers)
    cx = bench.Ctx(args)
    if args.prefill:
        mode, int8_kv, B, in_len, out_len = "sq", True, 8, 2048, 8
    else:
        mode, int8_kv, B, in_len, out_len, _ = bench.WORKLOADS[args.workload]
    sess, tensors = bench.build_session(cx, mode, int8_kv, B, in_len, out_len, 1, 0, graph=False)
    sess.set_decode_mode(bool(args.fused))
    lib = cx.lib
    ids = torch.randint(3, 32000, (B, in_len), dtype=torch.int32).cuda()
    lens = torch.full((B,), in_len, dtype=torch.int32).cuda()
    st = torch.cuda.current_stream().cuda_stream
    assert lib.tbrt_context(sess._e, ids.data_ptr(), lens.data_ptr(), B, in_len, st) == 0
    if args.prefill:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("profile")
        assert lib.tbrt_context(sess._e, ids.data_ptr(), lens.data_ptr(), B, in_len, st) == 0
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        return
    for _ in range(3):
        assert lib.tbrt_step(sess._e, st) == 0
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("profile")
    for _ in range(args.steps):
        assert lib.tbrt_step(sess._e, st) == 0
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print("launches in the last step:", sess.last_launches)


if __name__ == "__main__":
    main()
