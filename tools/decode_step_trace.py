#!/usr/bin/env python
"""Where a fused decode step spends its time: %globaltimer stamps recorded by every CTA of tb_decode_step (diagnostics hook
tb_decode_step_trace) for one step of a LLaMA-7B engine, summarised per phase type: barrier wait, activation staging,
weight stages, epilogue — mean over layers, for the fastest / median / slowest CTA.

    python tools/decode_step_trace.py [--workload cfg2|sq|cfg5|w8_b1]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--nccl-only", action="store_true")
    args = ap.parse_args()
    import torch
    import bench
    cx = bench.Ctx(args)
    mode, int8_kv, B, in_len, out_len, _ = bench.WORKLOADS[args.workload]
    sess, tensors = bench.build_session(cx, mode, int8_kv, B, in_len, out_len, cx.world, cx.rank)
    lib = cx.lib
    ids = torch.randint(3, 32000, (B, in_len), dtype=torch.int32).cuda()
    lens = torch.full((B,), in_len, dtype=torch.int32).cuda()
    st = torch.cuda.current_stream().cuda_stream
    lib.tbrt_context(sess._e, ids.data_ptr(), lens.data_ptr(), B, in_len, st)
    for _ in range(8):
        lib.tbrt_step(sess._e, st)
    torch.cuda.synchronize()
    h = lib.tbrt_decode_step_handle(sess._e)
    assert h, "fused step not available"
    G, SL = 148, 2048
    grid = C.c_int(0)
    lib.tb_decode_step_info(h, B, None, None, C.byref(grid))
    G = grid.value
    assert lib.tb_decode_step_trace(h, 1, None) == 0
    lib.tbrt_step(sess._e, st)
    torch.cuda.synchronize()
    out = np.zeros((G, SL), np.uint64)
    assert lib.tb_decode_step_trace(h, 0, out.ctypes.data_as(C.c_void_p)) == 0
    if cx.rank != 0:
        cx.barrier()
        return
    L = 32
    t = out.astype(np.int64)
    t0 = t[:, 0].min()
    # per layer: qkv 4 stamps, attn 2, dense 4, fc 4, proj 4 = 18 stamps; stamp 0 is the kernel start
    per = 18
    names = ["qkv", "dense", "fc_gate", "proj"]
    offs = {"qkv": 0, "dense": 6, "fc_gate": 10, "proj": 14}
    print(f"step total (first stamp -> last stamp of layer 31): {(t[:, 1 + per * L - 1].max() - t0) / 1e3:.1f} us, {G} CTAs")
    prev_end = None
    for nm in names:
        w, sx, rp, ep = [], [], [], []
        for li in range(L):
            b = 1 + li * per + offs[nm]
            s = t[:, b:b + 4]
            # barrier wait: from this CTA's previous stamp (end of its previous phase) to the post-barrier stamp
            prev = t[:, b - 1]
            w.append(s[:, 0] - prev)
            sx.append(s[:, 1] - s[:, 0]); rp.append(s[:, 2] - s[:, 1]); ep.append(s[:, 3] - s[:, 2])
        f = lambda a: np.mean(np.stack(a), 0) / 1e3   # noqa: E731  per CTA, mean over layers, us
        for label, a in (("wait (incl. barrier)", w), ("stage_x", sx), ("stages", rp), ("epilogue", ep)):
            v = f(a)
            print(f"{nm:8s} {label:22s} min {v.min():6.2f}  med {np.median(v):6.2f}  max {v.max():6.2f} us")
        tot = f(w) + f(sx) + f(rp) + f(ep)
        print(f"{nm:8s} {'phase total':22s} med {np.median(tot):6.2f} us")
    aw, ab = [], []
    for li in range(L):
        b = 1 + li * per + 4
        aw.append(t[:, b] - t[:, b - 1]); ab.append(t[:, b + 1] - t[:, b])
    v, u = np.mean(np.stack(aw), 0) / 1e3, np.mean(np.stack(ab), 0) / 1e3
    print(f"attention wait min {v.min():6.2f} med {np.median(v):6.2f} max {v.max():6.2f}; body min {u.min():6.2f} med {np.median(u):6.2f} max {u.max():6.2f} us")
    # skew: spread of the post-barrier stamps across CTAs (how simultaneously CTAs leave a barrier)
    sk = [np.ptp(t[:, 1 + li * per]) for li in range(1, L)]
    print(f"spread of barrier exit across CTAs: mean {np.mean(sk) / 1e3:.2f} us")
    # time from the LAST CTA finishing a phase's epilogue to the FIRST / median CTA leaving the next barrier
    lat = []
    for li in range(L):
        b = 1 + li * per
        lat.append(np.median(t[:, b + 4]) - t[:, b + 3].max())   # qkv epilogue end -> attention barrier exit
    print(f"barrier latency (last arrival -> median exit): mean {np.mean(lat) / 1e3:.2f} us")
    lag = []
    for li in range(L):
        b = 1 + li * per
        lag.append(t[:, b + 3].max() - np.median(t[:, b + 3]))
    print(f"straggler lag at the qkv epilogue (slowest - median CTA): mean {np.mean(lag) / 1e3:.2f} us")
    if cx.world > 1:
        cx.barrier()


if __name__ == "__main__":
    main()
