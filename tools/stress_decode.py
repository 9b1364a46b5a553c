"""Graph-replay stress of the decode step: build Llama-3-8B-geometry layers, prefill, capture the whole greedy decode
step as one CUDA graph (GreedyDecodeLoop, exactly what bench.py times) and replay it `--replays` times, checking
that the sampled-token stream is identical across two passes from the same state (a race that corrupts data shows
up as a different stream; one that dead-locks a pipeline trips the device-side hang watchdog, whose record is
printed).  Run in its own process: a trapped kernel kills the CUDA context.

    python tools/stress_decode.py --batch 1 --layers 4 --replays 2000 --pdl 1
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--quant", default="int4")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--prompt-len", type=int, default=128)
    ap.add_argument("--replays", type=int, default=2000)
    ap.add_argument("--pdl", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    os.environ["B2L_PDL"] = str(a.pdl)

    from lite_llama_b200 import _lib
    from lite_llama_b200.executor import ModelRunner
    from lite_llama_b200.executor.model_runner import GreedyDecodeLoop

    torch.cuda.set_device(a.device)
    dev = f"cuda:{a.device}"
    B, P = a.batch, a.prompt_len
    # the context wraps: every `span` replays the loop state is reset, so that KV rows stay inside the pool
    span = 256
    bucket = 512
    max_seq_len = bucket
    out = {"batch": B, "layers": a.layers, "pdl": a.pdl, "replays": a.replays, "quant": a.quant, "seed": a.seed}
    t0 = time.time()
    try:
        runner = ModelRunner.build(a.model, max_seq_len=max_seq_len, max_gpu_num_blocks=B * max_seq_len, device=dev,
                                   quantization=None if a.quant == "fp16" else a.quant, seed=a.seed,
                                   num_layers=a.layers)
        cfg = runner.config
        gen = torch.Generator(device=dev).manual_seed(77 + a.seed)
        prompts = torch.randint(0, cfg.vocab_size, (B, P), generator=gen, device=dev)
        lens = torch.full((B,), P, dtype=torch.int64, device=dev)
        b_req_idx = torch.arange(B, dtype=torch.int64, device=dev)
        positions = torch.arange(P, device=dev).repeat(B, 1)
        runner.kv_cache_manager.free_all()
        runner.prefill_alloc_kv_cache(P, lens.clone(), b_req_idx)
        logits = runner.forward(prompts, positions, last_token_index=lens - 1)
        first = logits[:, -1].argmax(-1)
        torch.cuda.synchronize()
        loop = GreedyDecodeLoop(runner, B, bucket)
        loop.begin(first, lens, b_req_idx, max_steps=span)       # reserves span * B consecutive KV rows
        first_row = loop._reserved[0]
        loop.capture()
        out["launches_per_step"] = loop.launches_per_step

        def run_pass(n):
            loop.begin(first, lens, b_req_idx, first_row)
            toks = []
            for _ in range(n):
                loop.step()
                toks.append(loop.next_tokens.clone())
            torch.cuda.synchronize()
            return torch.stack(toks)

        ref = run_pass(span)
        done, mism = 0, 0
        while done < a.replays:
            n = min(span, a.replays - done)
            got = run_pass(n)
            if not torch.equal(got, ref[:n]):
                mism += 1
            done += n
        out["mismatching_passes"] = mism
        out["ok"] = mism == 0
    except Exception as ex:  # a trapped kernel surfaces here as a CUDA error
        out["ok"] = False
        out["error"] = repr(ex)[:400]
        try:
            out["hang"] = _lib.hang_report()
        except Exception as ex2:
            out["hang_error"] = repr(ex2)[:200]
    out["seconds"] = round(time.time() - t0, 1)
    print(json.dumps(out), flush=True)
    sys.stdout.flush()
    # a dead context can block interpreter teardown: leave at once
    os._exit(0 if out.get("ok") else 1)


if __name__ == "__main__":
    main()
