#!/usr/bin/env python
"""examples/llama_quant/run.py of the reference, on the B200 plugin engine (LQ/run.py:29-205): loads config.json + the
rank's engine, generates greedily and reports the mean latency of iterations 5..54 of 55 (LQ/run.py:117-198).
Without tokenizer.model (none offline) use --input_tokens (CSV / .npy of token ids), as the reference allows."""
import argparse
import csv
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))

EOS_TOKEN = 2
PAD_TOKEN = 2


def parse_arguments():
    p = argparse.ArgumentParser()
    p.add_argument('--max_output_len', type=int, required=True)
    p.add_argument('--log_level', type=str, default='error')
    p.add_argument('--engine_dir', type=str, default='llama_outputs')
    p.add_argument('--tokenizer_dir', type=str, default=".", help="Directory containing the tokenizer.model.")
    p.add_argument('--input_text', type=str, default='Born in north-east France, Soyer trained as a')
    p.add_argument('--input_tokens', dest='input_file', type=str, default=None,
                   help='CSV or Numpy file containing tokenized input. Alternative to text input.')
    p.add_argument('--output_csv', type=str, default=None, help='CSV file where the tokenized output is stored.')
    p.add_argument('--output_npy', type=str, default=None, help='Numpy file where the tokenized output is stored.')
    p.add_argument('--num_beams', type=int, default=1, help="Use beam search if num_beams >1")
    p.add_argument('--iterations', type=int, default=55, help="timing loop length (reference: 55, first 5 dropped)")
    return p.parse_args()


def read_input_ids(args):
    if args.input_file is not None:
        if args.input_file.endswith('.csv'):
            with open(args.input_file) as f:
                rows = [np.array(r, dtype='int32') for r in csv.reader(f, delimiter=',')]
            return rows
        if args.input_file.endswith('.npy'):
            a = np.load(args.input_file).astype('int32')
            return [r for r in (a if a.ndim == 2 else a[None])]
        raise SystemExit('Input format not supported.')
    tok_path = os.path.join(args.tokenizer_dir, "tokenizer.model")
    if not os.path.exists(tok_path):
        raise SystemExit(f"{tok_path} not found: pass --input_tokens (no tokenizer is available offline)")
    from transformers import LlamaTokenizer
    tok = LlamaTokenizer.from_pretrained(args.tokenizer_dir, legacy=False)
    return [np.array(tok.encode(args.input_text), dtype='int32')]


def setup_tp(lib, world, rank):
    """communicator bootstrap through torch.distributed (replaces the MPI exchange of allreducePlugin.cpp:128-167)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank % torch.cuda.device_count()))
    idbuf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        assert lib.tb_comm_unique_id(idbuf.data_ptr()) == 0
    idd = idbuf.cuda()
    dist.broadcast(idd, 0)
    idbuf = idd.cpu()
    group = (C.c_int32 * world)(*range(world))
    assert lib.tb_comm_init(idbuf.data_ptr(), group, world, rank) == 0


def generate(args):
    import torch
    from trtllm_llama_b200 import builder as B
    from trtllm_llama_b200 import runtime as rt
    from trtllm_llama_b200._lib import lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    mc = B.model_config_from_json(os.path.join(args.engine_dir, "config.json"), rank)
    assert world == mc.tp_size, f'Engine world size ({mc.tp_size}) != Runtime world size ({world})'
    torch.cuda.set_device(rank % torch.cuda.device_count())
    if world > 1:
        setup_tp(lib, world, rank)
    tensors = B.deserialize_engine(os.path.join(args.engine_dir, B.get_engine_name("llama", "float16", world, rank)))
    rows = read_input_ids(args)
    max_in = max(len(r) for r in rows)
    ids = np.full((len(rows), max_in), PAD_TOKEN, dtype=np.int32)
    for i, r in enumerate(rows):
        ids[i, :len(r)] = r
    lens = np.array([len(r) for r in rows], dtype=np.int32)
    session = rt.GenerationSession(mc, tensors)
    if world > 1:
        session.enable_peer_allreduce()
    session.setup(len(rows), max_in, args.max_output_len, beam_width=args.num_beams)
    host_ids, host_lens = torch.from_numpy(ids).pin_memory(), torch.from_numpy(lens).pin_memory()
    sampling = rt.SamplingConfig(end_id=EOS_TOKEN, pad_id=PAD_TOKEN, num_beams=args.num_beams)
    lat = []
    for _ in range(args.iterations):
        t0 = time.time()
        out = session.decode(host_ids, host_lens, sampling)
        torch.cuda.synchronize()
        lat.append(time.time() - t0)
    if rank == 0:
        out = out.numpy()
        for b in range(len(rows)):
            print(f'Input ids: {rows[b].tolist()}')
            print(f'Output ids: {out[b].tolist()}')      # num_beams > 1: [num_beams, output_len], best beam first
        if args.output_csv:
            with open(args.output_csv, 'w') as f:
                csv.writer(f, delimiter=',').writerows(out.reshape(-1, out.shape[-1]).tolist())
        if args.output_npy:
            np.save(args.output_npy, out)
        drop = 5 if len(lat) > 5 else 0
        mean = float(np.mean(lat[drop:]))
        print(f'TensorRT-LLM mean latency: {mean:.5f} sec   ({len(rows) * args.max_output_len / mean:.1f} tokens/s)')


if __name__ == '__main__':
    generate(parse_arguments())
