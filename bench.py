"""Headline benchmark: greedy decode throughput of Llama-3-8B W4A16 at batch 32, 128 -> 128 tokens.

    python bench.py --gpus N --steps K --warmup W            # this framework (B200 kernels)
    python bench.py --impl reference --gpus N --steps K ...  # CPU port of the reference path (host cores)

One "step" = one decode step of the whole batch (every sequence gains one token). The model is
random-initialised with the named architecture's shapes (no checkpoints offline), prompts are
synthetic token ids; a 128-token prefill fills the KV cache, then W warm-up and K timed decode steps
run with contexts growing from 129. Every rank holds a full replica (data parallel, no collective:
the path shards by independent sequences), so scaling is weak and `value` is the sum over ranks.

JSON keys follow the driver contract; see DESIGN.md "Measurement" for what each figure means.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "decode_tokens_per_s"
UNIT = "tok/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--quant", default="int4", choices=["fp16", "int4", "int8", "fp8", "smoothquant"])
    ap.add_argument("--max-seq-len", type=int, default=None, help="KV slot length per request (default: fits the run)")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--prompt-len", type=int, default=128)
    ap.add_argument("--layers", type=int, default=None, help="override the layer count (debugging only)")
    ap.add_argument("--parallel", default="tp", choices=["tp", "dp"],
                    help="N > 1: tp = ONE model sharded over the N GPUs (column/row-parallel projections, one exchange per "
                         "row-parallel projection; strong scaling, same global batch); dp = N independent replicas "
                         "(weak scaling, no collective)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hf-baseline", action="store_true", help="skip the HF-transformers-on-CPU row (configs[0])")
    ap.add_argument("--watchdog-s", type=int, default=600, help="give up (exit 3) after this many seconds; 0 = never")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--no-dp-line", action="store_true", help="N > 1, tp: skip the extra replica (weak-scaling) measurement")
    ap.add_argument("--cpu-layers", type=int, default=2, help="decoder layers in the CPU sample")
    return ap.parse_args()


def workload_config(args, world):
    tp = world > 1 and getattr(args, "parallel", "tp") == "tp"
    return {
        "workload": f"{args.model} {'W4A16 g128' if args.quant == 'int4' else args.quant} greedy decode, "
                    f"batch {args.batch}{'' if tp else '/GPU'}, prompt {args.prompt_len} -> {args.steps} new tokens "
                    + ("(BASELINE.json configs[2]; the metric's named config)"
                       if (args.model, args.quant) == ("llama-3-8b", "int4") else "(a BASELINE.json side configuration)"),
        "model": args.model, "quant": args.quant, "per_gpu_batch": args.batch,
        "global_batch": args.batch if tp else args.batch * world,
        "prompt_len": args.prompt_len, "parallelism": (f"tp{world}" if tp else f"dp{world}") if world > 1 else "single",
        "l2_policy": "inputs_larger_than_l2 (4.98 GB of weights streamed per step vs 126 MB L2)",
        "cuda_graph": "one captured graph per decode step",
        "weight_layout": ("int4 tensors repacked once at load into the kernel-native layout (bit-exact nibble / table "
                          "permutation, lite_llama_b200.kernels.w4a16_pack); B2L_W4_LAYOUT=reference keeps the "
                          "reference's tensors") if args.quant == "int4" else "reference",
    }


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/ncu_traffic.json), or
    None when no capture of that kernel at this shape is on file."""
    f = ROOT / "profiles" / "ncu_traffic.json"
    try:
        return json.loads(f.read_text())[kernel]["bytes"]
    except Exception:
        return None


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), "measured"
    return 6650.0, 1400.0, "fallback"


# --------------------------------------------------------------------------- #
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------- #
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index, self.samples, self._stop = index, [], threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self) -> None:
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no_samples"]}
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 3 + i and s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------- #
# CPU baseline: the oracle's decode step on the host cores (bounded sample)
# --------------------------------------------------------------------------- #
def cpu_decode_sample(args, steps: int, warmup: int, threads: int | None = None) -> dict:
    from lite_llama_b200.executor.model_config import PRESETS
    from oracle import model as omodel

    if threads:
        torch.set_num_threads(threads)
    cfg = PRESETS[args.model]
    n_layers = max(1, min(args.cpu_layers, cfg.num_layers))
    B, P = args.batch, args.prompt_len
    ctx = P + 1
    t0 = time.time()
    # random weights drawn directly (no 128k x 4096 embedding draw needed: the gather is not timed work)
    gen = torch.Generator().manual_seed(0)

    def draw(*shape):
        return torch.randn(*shape, generator=gen, dtype=torch.float32) * 0.02

    small_vocab_cfg = cfg  # lm_head keeps its real shape: it is 21 % of the step's bytes
    embed, lm_head, layers = omodel.random_cpu_layers(small_vocab_cfg, args.quant, n_layers, draw=draw)
    dec = omodel.CpuDecoder(cfg, embed, lm_head, layers, max_tokens=B * (ctx + steps + warmup + 2),
                            max_pos=ctx + steps + warmup + 2, method=args.quant)
    for kv in dec.kv:
        kv.normal_(0, 1)
    build_s = time.time() - t0
    width = ctx + steps + warmup + 2
    table = torch.arange(B * width, dtype=torch.int32).view(B, width)
    b_req = torch.arange(B)
    ids = torch.randint(0, cfg.vocab_size, (B, 1), generator=gen)

    def one_step(i):
        seq = torch.full((B,), ctx + i, dtype=torch.int64)
        pos = (seq - 1).view(B, 1)
        sel = table[torch.arange(B), seq - 1]
        logits = dec.forward(ids, pos, sel, table, b_req, seq)
        return logits[:, -1].argmax(-1)

    for i in range(warmup):
        one_step(i)
    t = time.time()
    for i in range(steps):
        one_step(warmup + i)
    per_step_sample = (time.time() - t) / steps
    # time the lm_head share separately so the layer part can be scaled to the full depth
    x = torch.randn(B, cfg.hidden_size).half()
    t = time.time()
    for _ in range(max(1, steps)):
        omodel._lin(x, lm_head.float())
    lm_s = (time.time() - t) / max(1, steps)
    layers_s = max(per_step_sample - lm_s, 0.0)
    full_step = layers_s * (cfg.num_layers / n_layers) + lm_s
    return {
        "value": B / full_step, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
        "sample": f"{steps} decode steps of {n_layers}/{cfg.num_layers} layers + full lm_head at batch {B}, ctx {ctx}+; "
                  f"layer time scaled x{cfg.num_layers / n_layers:.0f}; weights pre-dequantised to fp32 "
                  f"(oracle/model.py); host has {os.cpu_count()} logical cores; build {build_s:.0f}s",
        "ms_per_step_full_model": full_step * 1e3,
    }


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm must use the host's cores whatever launched it
    threads = os.cpu_count() or 1
    steps = max(8, min(args.steps, 12))       # every timed step is a bounded sample; 8-12 of them keep the run short
    warm = max(1, min(args.warmup, 2))
    base = cpu_decode_sample(args, steps, warm, threads=threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "requested_steps": args.steps, "requested_warmup": args.warmup,
        "ms_per_step": base["ms_per_step_full_model"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),     # the same workload description as this framework's arm at N GPUs
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": f"CPU port of the reference decode path; each of the {steps} timed steps is a bounded sample "
                f"(see cpu_baseline.sample); world={world}, rank 0 only",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- #
# per-kernel breakdown (graph replays over all layers => working set >> L2)
# --------------------------------------------------------------------------- #
def time_graph(fn, reps: int = 5) -> float:
    """Median device milliseconds of one replay of ``fn`` captured as a CUDA graph."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def kernel_breakdown(runner, batch: int, ctx: int, hbm_peak: float) -> list[dict]:
    import lite_llama_b200.kernels as K
    from lite_llama_b200.quantization import apply_linear

    dec = runner.model
    cfg = runner.config
    dev = runner.device
    L = len(dec.w.layers)
    x_h = torch.randn(batch, cfg.hidden_size, device=dev, dtype=torch.float16)
    x_q = torch.randn(batch, dec.q_size, device=dev, dtype=torch.float16)
    x_i = torch.randn(batch, max(dec.inter, 8), device=dev, dtype=torch.float16)
    out = []

    def add(name, fn, bytes_per_launch, launches):
        ms = time_graph(fn)
        per = ms / launches
        gbs = bytes_per_launch / per / 1e6
        out.append({"kernel": name, "us_per_launch": per * 1e3, "launches_per_step": launches,
                    "algorithmic_bytes_per_launch": int(bytes_per_launch), "achieved_gbs": gbs,
                    "frac_hbm": gbs / hbm_peak, "us_per_step": ms * 1e3})

    lw0 = dec.w.layers[0]
    add("qkv_proj", lambda: [apply_linear(x_h, lw.qkv) for lw in dec.w.layers], lw0.qkv.nbytes(), L)
    add("o_proj", lambda: [apply_linear(x_q, lw.o) for lw in dec.w.layers], lw0.o.nbytes(), L)
    if lw0.moe is None:
        add("gate_up_proj", lambda: [apply_linear(x_h, lw.gate_up) for lw in dec.w.layers], lw0.gate_up.nbytes(), L)
        add("down_proj", lambda: [apply_linear(x_i, lw.down) for lw in dec.w.layers], lw0.down.nbytes(), L)
    else:
        m0 = lw0.moe
        touched = m0.w1.shape[0] * (1.0 - (1.0 - m0.top_k / m0.w1.shape[0]) ** batch)
        add("moe_block(router+fused_moe)", lambda: [dec._moe(x_h, lw.moe) for lw in dec.w.layers],
            touched * dec.w.moe_expert_bytes() + m0.router.numel() * 2, L)
    add("lm_head", lambda: K.linear(x_h, dec.w.lm_head), dec.w.lm_head.numel() * 2, 1)
    # flash decoding over every layer's pool at the mean timed context
    info = runner.atten_info
    q = torch.randn(batch, dec.num_heads, dec.head_dim, device=dev, dtype=torch.float16)
    seq = torch.full((batch,), ctx, dtype=torch.int32, device=dev)
    req = torch.arange(batch, dtype=torch.int32, device=dev)
    bucket = 256 if ctx <= 256 else 1 << (ctx - 1).bit_length()
    kv_bytes = batch * ctx * 2 * dec.num_kv_heads * dec.head_dim * 2
    add("flash_decoding",
        lambda: [K.flash_decoding(q, kv[:, : dec.num_kv_heads], kv[:, dec.num_kv_heads:], dec.scale,
                                  info.b_req_tokens_table, req, seq, bucket) for kv in info.kv_buffer],
        kv_bytes, L)
    res = torch.randn(batch, 1, cfg.hidden_size, device=dev, dtype=torch.float16)
    add("skip_rmsnorm", lambda: [K.skip_rmsnorm(x_h.view(batch, 1, -1), res, lw.input_norm, 1e-5) for lw in dec.w.layers],
        batch * cfg.hidden_size * 2 * 5, L)
    return out


# --------------------------------------------------------------------------- #
def arm_watchdog(seconds: int) -> None:
    """A hung kernel must not hold the GPU box until the caller's limit kills it: after `seconds` print what is
    known and leave."""
    def fire():
        print(json.dumps({"error": f"bench.py watchdog: no result after {seconds} s (device hang?)"}), flush=True)
        os._exit(3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()


def main() -> None:
    args = parse_args()
    if args.watchdog_s > 0:
        arm_watchdog(args.watchdog_s)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    tp = world > 1 and args.parallel == "tp"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if tp:
            # one model over all ranks: NCCL group + the NVLink peer-memory communicator of the fused exchange kernel
            from lite_llama_b200.distributed import init_tensor_parallel

            init_tensor_parallel()
        else:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    # TP ranks hold shards of ONE model and step the SAME batch: identical seeds; DP replicas differ
    seed = 0 if tp else rank

    from lite_llama_b200 import _lib
    from lite_llama_b200.executor import ModelRunner
    from lite_llama_b200.executor.model_runner import GreedyDecodeLoop
    from lite_llama_b200.kernels import argmax_rows

    lib = _lib.load()
    B, P, K, W = args.batch, args.prompt_len, args.steps, args.warmup
    total_len = P + 2 * (K + W) + 8            # device-resident pass + end-to-end pass
    bucket = 256
    while bucket < P + K + W + 1:
        bucket *= 2
    max_seq_len = max(bucket, total_len)
    t_build = time.time()
    runner = ModelRunner.build(args.model, max_seq_len=max_seq_len, max_gpu_num_blocks=B * max_seq_len, device=dev,
                               quantization=None if args.quant == "fp16" else args.quant, seed=seed,
                               num_layers=args.layers)
    # capture the public-API decode graph before any KV is live (its warm-up writes dummy rows)
    runner.enable_cuda_graph(batch_sizes=(B,), seq_len_buckets=(bucket,))
    torch.cuda.synchronize()
    build_s = time.time() - t_build
    cfg = runner.config
    weight_bytes = runner.model.w.nbytes_per_step()

    gen = torch.Generator(device=dev).manual_seed(1234 + seed)
    prompts = torch.randint(0, cfg.vocab_size, (B, P), generator=gen, device=dev)
    lens = torch.full((B,), P, dtype=torch.int64, device=dev)
    b_req_idx = torch.arange(B, dtype=torch.int64, device=dev)
    positions = torch.arange(P, device=dev).repeat(B, 1)

    def prefill():
        runner.kv_cache_manager.free_all()
        runner.prefill_alloc_kv_cache(P, lens.clone(), b_req_idx)
        logits = runner.forward(prompts, positions, last_token_index=lens - 1)
        return logits[:, -1].argmax(-1)

    # ---- prefill (timed separately: TTFT and TFLOP/s) ----
    first = prefill()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    first = prefill()
    e.record()
    torch.cuda.synchronize()
    prefill_ms = s.elapsed_time(e)
    def _ffn_params(lw):
        if lw.moe is not None:      # active experts only
            return lw.moe.top_k * (lw.moe.w1.shape[1] * lw.moe.w1.shape[2] + lw.moe.w2.shape[1] * lw.moe.w2.shape[2])
        return lw.gate_up.out_features * lw.gate_up.in_features + lw.down.out_features * lw.down.in_features

    params_active = sum(lw.qkv.out_features * lw.qkv.in_features + lw.o.out_features * lw.o.in_features + _ffn_params(lw)
                        for lw in runner.model.w.layers)
    prefill_flops = 2.0 * B * P * params_active + 2.0 * B * cfg.vocab_size * cfg.hidden_size \
        + 4.0 * B * cfg.num_heads * P * P * cfg.head_dim / 2 * cfg.num_layers
    prefill_tflops = prefill_flops / prefill_ms / 1e9

    # ---- device-resident decode: one captured graph per step, zero host work ----
    loop = GreedyDecodeLoop(runner, B, bucket)
    loop.begin(first, lens, b_req_idx, max_steps=K + W + 1)      # reserves its KV rows through the allocator
    loop.capture()
    for _ in range(W):
        loop.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = lib.b2l_launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        s.record()
        for _ in range(K):
            loop.step()
        e.record()
        torch.cuda.synchronize()
    dev_ms = s.elapsed_time(e)
    if world > 1:
        dist.barrier()
        t = torch.tensor([dev_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    tokens_dev = loop.next_tokens.clone()
    gpu_launches = loop.launches_per_step * K   # replayed graph nodes of this library (torch glue not counted)

    # ---- end to end through the public API with host buffers ----
    first = prefill()
    host_in = torch.empty((B, 1), dtype=torch.int64).pin_memory()
    host_out = torch.empty((B,), dtype=torch.int64).pin_memory()
    host_in.copy_(first.view(B, 1).cpu())
    pos = lens.clone().view(B, 1)
    ids_dev = torch.empty((B, 1), dtype=torch.int64, device=dev)
    runner.atten_info.b_seq_len = lens.clone()
    runner.atten_info.max_actual_seq_len = P

    def e2e_step():
        ids_dev.copy_(host_in, non_blocking=True)                      # H2D: this step's input token ids
        runner.decode_alloc_kv_cache(B)                                # reserve rows, bump lengths, update the table
        logits = runner.forward(ids_dev, pos)                          # public API (captured decode graph)
        nxt = argmax_rows(logits[:, -1, :])                            # greedy sampling (library kernel, first-index ties like torch)
        host_out.copy_(nxt, non_blocking=True)                         # D2H: sampled tokens
        torch.cuda.current_stream().synchronize()
        host_in.copy_(host_out.view(B, 1))                             # the host feeds the next step
        pos.add_(1)

    for _ in range(W):
        e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())

    # ---- tensor parallel: what one exchange costs (all ranks take part) ----
    tp_info = None
    if tp:
        from lite_llama_b200.distributed import all_reduce_tp, get_tp_comm
        from lite_llama_b200.kernels import skip_rmsnorm
        from lite_llama_b200.kernels.tp_fused import allreduce_skip_rmsnorm

        xx = torch.randn(B, cfg.hidden_size, device=dev, dtype=torch.float16)
        rr = torch.randn(B, cfg.hidden_size, device=dev, dtype=torch.float16)
        ww = torch.ones(cfg.hidden_size, device=dev, dtype=torch.float16)
        calls = 64
        dist.barrier()
        us_fused = time_graph(lambda: [allreduce_skip_rmsnorm(xx, rr, ww) for _ in range(calls)]) * 1e3 / calls
        dist.barrier()
        us_nccl = time_graph(lambda: [skip_rmsnorm(all_reduce_tp(xx), rr, ww) for _ in range(calls)]) * 1e3 / calls
        t = torch.tensor([us_fused, us_nccl], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_ex = 2 * cfg.num_layers
        msg = B * cfg.hidden_size * 2
        tp_info = {
            "exchange": ("one kernel per row-parallel projection: push partial rows to every peer over NVLink, sum in rank "
                         "order, residual add + RMSNorm (csrc/tp_comm.cu)") if get_tp_comm() is not None
                        else "NCCL all-reduce + skip_rmsnorm (NVLink communicator unavailable)",
            "nvlink_comm": get_tp_comm() is not None, "exchanges_per_step": n_ex, "message_bytes": msg,
            "us_per_exchange_fused": float(t[0].item()), "us_per_nccl_allreduce_plus_rmsnorm": float(t[1].item()),
            "exchange_share_of_step": float(t[0].item()) * n_ex / (dev_ms / K * 1e3),
            "limiting_collective": f"{n_ex} latency-bound {msg // 1024} KiB exchanges per step",
        }

    # ---- the same GPUs as N independent replicas (no collective): a second, weak-scaling line beside the TP one ----
    dp_info = None
    if tp and not args.no_dp_line:
        try:
            dp_info = dp_replica_line(args, dev, rank, world, bucket, max_seq_len, K, W)
        except Exception as ex:  # the TP headline must survive
            dp_info = {"error": repr(ex)[:300]}

    if rank != 0:
        finish(world)
        return

    hbm_peak, tf_peak, peak_src = peaks()
    replicas = 1 if tp else world
    value = replicas * B * K / (dev_ms / 1e3)
    e2e_value = replicas * B * K / (e2e_ms / 1e3)
    mean_ctx = P + W + (K + 1) / 2
    kv_bytes_step = B * mean_ctx * 2 * runner.num_kv_heads * cfg.head_dim * 2 * cfg.num_layers
    step_bytes = weight_bytes + kv_bytes_step
    if cfg.num_experts > 0:
        # routed experts: expected number of DISTINCT experts a batch of B tokens touches per layer
        touched = cfg.num_experts * (1.0 - (1.0 - cfg.num_experts_per_tok / cfg.num_experts) ** B)
        step_bytes += touched * runner.model.w.moe_expert_bytes() * cfg.num_layers
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong" if tp else "weak", "vs_baseline": None,
        "dtype": {"int4": "f16 activations, int4 weights (fp32 accumulate)", "fp16": "f16",
                  "fp8": "f16 activations, fp8-e4m3 weights (fp32 accumulate)", "int8": "f16 activations, int8 weights",
                  "smoothquant": "int8 activations x int8 weights (int32 accumulate), f16 elsewhere"}[args.quant],
        "data": "synthetic (random-init weights of the named shapes, random token ids)",
        "config": workload_config(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 8, "d2h_bytes_per_step": B * 8,
                "ms_per_step": e2e_ms / K},
        "gpu_launches": int(gpu_launches),
        "clocks": clk.summary(),
        "step_roofline": {"bound": "hbm", "algorithmic_bytes_per_step": int(step_bytes),
                          "achieved_gbs": step_bytes / (dev_ms / K) / 1e6,
                          "frac": step_bytes / (dev_ms / K) / 1e6 / hbm_peak, "peak_gbs": hbm_peak, "peak_source": peak_src},
        "prefill": {"ms": prefill_ms, "tflops": prefill_tflops, "frac_tensor_peak": prefill_tflops / tf_peak,
                    "tokens": B * P},
        "build_s": build_s,
        "sampled_token_checksum": int(tokens_dev.sum().item()),
    }
    if tp_info is not None:
        line["tensor_parallel"] = tp_info
    if dp_info is not None:
        line["dp_replicas"] = dp_info
    if not args.no_breakdown:
        try:
            kb = kernel_breakdown(runner, B, int(mean_ctx), hbm_peak)
            line["kernels"] = kb
            top = max(kb, key=lambda r: r["us_per_step"])
            line["roofline"] = {
                "bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved_gbs"], "peak": hbm_peak,
                "unit": "GB/s", "frac": top["frac_hbm"], "traffic": ncu_traffic(top["kernel"]), "peak_source": peak_src,
                "share_of_step": top["us_per_step"] / (dev_ms / K * 1e3),
            }
        except Exception as ex:  # the headline number must survive a breakdown failure
            line["kernels_error"] = repr(ex)[:300]
    if "roofline" not in line:
        line["roofline"] = {"bound": "hbm", "achieved": line["step_roofline"]["achieved_gbs"], "peak": hbm_peak,
                            "unit": "GB/s", "frac": line["step_roofline"]["frac"], "traffic": None}
    if not args.no_cpu_baseline and world == 1:      # reported baselines: rank 0 at N = 1 only
        try:
            base = cpu_decode_sample(args, steps=8, warmup=1, threads=os.cpu_count() or 1)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:
            line["cpu_baseline"] = {"error": repr(ex)[:300]}
    if not args.no_cpu_baseline and not args.no_hf_baseline and world == 1:
        # BASELINE.json configs[0]: HF transformers on the host cores (Qwen2.5-0.5B, greedy, one prompt) -- a reported
        # baseline row in the same run (tools/bench_hf_baseline.py), not what the GPU number is compared against
        try:
            sys.path.insert(0, str(ROOT / "tools"))
            import bench_hf_baseline

            line["hf_cpu_baseline"] = bench_hf_baseline.measure()
        except Exception as ex:
            line["hf_cpu_baseline"] = {"error": repr(ex)[:300]}
    print(json.dumps(line), flush=True)
    finish(world)


def dp_replica_line(args, dev, rank, world, bucket, max_seq_len, K, W) -> dict:
    """Device-timed decode throughput of `world` independent full replicas (one per GPU, different seeds, no
    collective) on the GPUs that just ran the tensor-parallel measurement: the weak-scaling reading of the same
    hardware. Built while the tensor-parallel state is masked, so the captured graphs contain no exchange."""
    import torch.distributed as dist

    from lite_llama_b200.distributed import parallel_state
    from lite_llama_b200.executor import ModelRunner
    from lite_llama_b200.executor.model_runner import GreedyDecodeLoop

    B, P = args.batch, args.prompt_len
    saved = (parallel_state._TP_SIZE, parallel_state._TP_RANK, parallel_state._TP_COMM)
    parallel_state._TP_SIZE, parallel_state._TP_RANK, parallel_state._TP_COMM = 1, 0, None
    try:
        runner = ModelRunner.build(args.model, max_seq_len=max_seq_len, max_gpu_num_blocks=B * max_seq_len, device=dev,
                                   quantization=None if args.quant == "fp16" else args.quant, seed=100 + rank,
                                   num_layers=args.layers)
        gen = torch.Generator(device=dev).manual_seed(4321 + rank)
        prompts = torch.randint(0, runner.config.vocab_size, (B, P), generator=gen, device=dev)
        lens = torch.full((B,), P, dtype=torch.int64, device=dev)
        req = torch.arange(B, dtype=torch.int64, device=dev)
        runner.prefill_alloc_kv_cache(P, lens.clone(), req)
        first = runner.forward(prompts, torch.arange(P, device=dev).repeat(B, 1), last_token_index=lens - 1)[:, -1].argmax(-1)
        loop = GreedyDecodeLoop(runner, B, bucket)
        loop.begin(first, lens, req, max_steps=K + W + 1)
        loop.capture()
        for _ in range(W):
            loop.step()
        torch.cuda.synchronize()
    finally:
        parallel_state._TP_SIZE, parallel_state._TP_RANK, parallel_state._TP_COMM = saved
    dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(K):
        loop.step()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"value": world * B * K / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / K, "scaling": "weak",
            "global_batch": world * B, "parallelism": f"dp{world}", "note": "independent replicas, no collective"}


def finish(world: int) -> None:
    """Multi-rank exit: meet once more, then leave without tearing NCCL down -- destroying a process group while
    captured graphs still hold its kernels can block for ever, and a bench that printed its line must not hang."""
    if world <= 1:
        return
    import torch.distributed as dist

    torch.cuda.synchronize()
    try:
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
