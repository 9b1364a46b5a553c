"""The UNMODIFIED reference engine (baseline/_ref = offline pip install of /root/reference) on a synthetic checkpoint,
three ways, same B200, same weights, same prompts, temperature 0:

  A  reference as shipped: its Triton kernels, its ModelRunner / CUDAGraphManager, LLMEngine._DecodeSession
  B  same engine, same model classes, kernel names re-bound to libb200lite.so (lite_llama_b200.integration.bind,
     INTEGRATION.md s2) -- "the kernel layer drops in"
  C  the reference's LLMEngine._DecodeSession and ContinuousBatchingEngine driving THIS package's ModelRunner
     (executor surface: prefill/decode_alloc_kv_cache, forward, kv_cache_manager.release_ref, spec,
     enable_slot_kv_cache, graph_batch_size) -- "the executor drops in"; weights are read from arm A's model

and reports decode tok/s of each (wall clock around the engine's own loop, after a warm-up generation) plus whether
the generated token grids are identical (token-for-token, north_star).  No checkpoint or tokenizer exists offline, so
the script writes one: config.json of the named geometry, random fp16 safetensors (randn * 0.02), and a WordLevel
tokenizer with one token per id.

    python tools/ref_engine_ab.py --model llama-3-8b --layers 32 --quant int4 --batch 32 --prompt-len 128 --gen 64
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import shutil
import sys
import time
import traceback
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "baseline" / "_ref"))

GEOM = {
    "llama-3-8b": dict(model_type="llama", architectures=["LlamaForCausalLM"], hidden_size=4096, num_hidden_layers=32,
                       num_attention_heads=32, num_key_value_heads=8, intermediate_size=14336, vocab_size=128256,
                       rope_theta=500000.0, rms_norm_eps=1e-5, max_position_embeddings=8192, head_dim=128,
                       tie_word_embeddings=False, hidden_act="silu", torch_dtype="float16"),
    "qwen2.5-7b": dict(model_type="qwen2", architectures=["Qwen2ForCausalLM"], hidden_size=3584, num_hidden_layers=28,
                       num_attention_heads=28, num_key_value_heads=4, intermediate_size=18944, vocab_size=152064,
                       rope_theta=1000000.0, rms_norm_eps=1e-6, max_position_embeddings=32768,
                       tie_word_embeddings=False, hidden_act="silu", torch_dtype="float16"),
    "tiny-llama": dict(model_type="llama", architectures=["LlamaForCausalLM"], hidden_size=256, num_hidden_layers=2,
                       num_attention_heads=8, num_key_value_heads=2, intermediate_size=512, vocab_size=1000,
                       rope_theta=10000.0, rms_norm_eps=1e-5, max_position_embeddings=2048, head_dim=32,
                       tie_word_embeddings=False, hidden_act="silu", torch_dtype="float16"),
}


def write_checkpoint(path: Path, geom: dict, layers: int, seed: int = 0, device: str = "cuda") -> None:
    from safetensors.torch import save_file
    from tokenizers import Tokenizer
    from tokenizers.models import WordLevel
    from tokenizers.pre_tokenizers import WhitespaceSplit
    from transformers import PreTrainedTokenizerFast

    path.mkdir(parents=True, exist_ok=True)
    cfg = dict(geom, num_hidden_layers=layers)
    (path / "config.json").write_text(json.dumps(cfg))
    V, h, inter = cfg["vocab_size"], cfg["hidden_size"], cfg["intermediate_size"]
    hd = cfg.get("head_dim", h // cfg["num_attention_heads"])
    qs, kvs = cfg["num_attention_heads"] * hd, cfg["num_key_value_heads"] * hd
    g = torch.Generator(device=device).manual_seed(seed)

    def draw(*shape):
        return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * 0.02).half().cpu()

    save_file({"model.embed_tokens.weight": draw(V, h), "lm_head.weight": draw(V, h),
               "model.norm.weight": torch.ones(h, dtype=torch.float16)}, str(path / "model-head.safetensors"))
    bias = cfg["model_type"] == "qwen2"
    for i in range(layers):
        p = f"model.layers.{i}."
        t = {p + "self_attn.q_proj.weight": draw(qs, h), p + "self_attn.k_proj.weight": draw(kvs, h),
             p + "self_attn.v_proj.weight": draw(kvs, h), p + "self_attn.o_proj.weight": draw(h, qs),
             p + "mlp.gate_proj.weight": draw(inter, h), p + "mlp.up_proj.weight": draw(inter, h),
             p + "mlp.down_proj.weight": draw(h, inter),
             p + "input_layernorm.weight": torch.ones(h, dtype=torch.float16),
             p + "post_attention_layernorm.weight": torch.ones(h, dtype=torch.float16)}
        if bias:
            t.update({p + "self_attn.q_proj.bias": draw(qs), p + "self_attn.k_proj.bias": draw(kvs),
                      p + "self_attn.v_proj.bias": draw(kvs)})
        save_file(t, str(path / f"model-layer{i:03d}.safetensors"))
    vocab = {f"t{i}": i for i in range(V - 2)}
    vocab["<eos>"] = V - 2
    vocab["<unk>"] = V - 1
    tok = Tokenizer(WordLevel(vocab, unk_token="<unk>"))
    tok.pre_tokenizer = WhitespaceSplit()
    PreTrainedTokenizerFast(tokenizer_object=tok, eos_token="<eos>", unk_token="<unk>", pad_token="<eos>").save_pretrained(str(path))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--quant", default="int4")
    ap.add_argument("--batch", default="32,1")
    ap.add_argument("--prompt-len", type=int, default=128)
    ap.add_argument("--gen", type=int, default=64)
    ap.add_argument("--arms", default="A,B,C")
    ap.add_argument("--continuous", type=int, default=0, help="also run ContinuousBatchingEngine with this many requests (arms B, C)")
    ap.add_argument("--out", default="gpurun_out/r2_ref_engine_ab.json")
    ap.add_argument("--workdir", default=None)
    a = ap.parse_args()
    quant = None if a.quant in ("fp16", "none") else a.quant
    work = Path(a.workdir or ("/dev/shm/b2l_ckpt" if os.path.isdir("/dev/shm") else "/tmp/b2l_ckpt"))
    shutil.rmtree(work, ignore_errors=True)
    result = {"model": a.model, "layers": a.layers, "quant": a.quant, "prompt_len": a.prompt_len, "gen": a.gen, "arms": {}}
    t0 = time.time()
    write_checkpoint(work, GEOM[a.model], a.layers)
    result["checkpoint_write_s"] = round(time.time() - t0, 1)

    from lite_llama.engine.llm_engine import LLMEngine, _DecodeSession
    from lite_llama.engine.sampler import SamplingParams

    import lite_llama_b200.integration as integ

    params = SamplingParams(temperature=0.0, max_gen_len=a.gen, repetition_penalty=1.0, stop_on_repeat=False)
    V = GEOM[a.model]["vocab_size"]
    batches = [int(b) for b in a.batch.split(",")]
    max_seq_len = 512
    while max_seq_len < a.prompt_len + a.gen + 8:
        max_seq_len *= 2
    blocks = max(batches + [a.continuous + 1]) * max_seq_len + max_seq_len

    def prompts_for(B):
        g = torch.Generator().manual_seed(1234 + B)
        return torch.randint(0, V - 2, (B, a.prompt_len), generator=g).tolist()

    def run_session(engine, B):
        """(token grid [B, P+gen], decode tok/s, ttft-ish prefill seconds) through the reference's own decode session."""
        engine.stop_token_ids = set()          # random weights: never stop on EOS (harness config, not a code change)
        ids = prompts_for(B)
        warm = SamplingParams(temperature=0.0, max_gen_len=8, repetition_penalty=1.0, stop_on_repeat=False)
        s = _DecodeSession(engine, ids, warm, None, None)
        for _ in s.run(stream=False):
            pass
        torch.cuda.synchronize()
        s = _DecodeSession(engine, ids, params, None, None)
        t = time.perf_counter()
        for _ in s.run(stream=False):
            pass
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        return s._tokens.clone().cpu(), B * a.gen / dt, dt

    grids = {}

    def arm(name, build):
        if name not in a.arms.split(","):
            return None
        info = {}
        try:
            t = time.time()
            engine = build()
            info["build_s"] = round(time.time() - t, 1)
            for B in batches:
                grid, tps, dt = run_session(engine, B)
                grids[(name, B)] = grid
                info[f"bs{B}"] = {"tok_per_s_incl_prefill": round(tps, 1), "seconds": round(dt, 3)}
            result["arms"][name] = info
            return engine
        except Exception as ex:
            info["error"] = repr(ex)[:500]
            info["trace"] = traceback.format_exc()[-2000:]
            result["arms"][name] = info
            return None

    # ---- A: the reference as shipped ----
    def build_ref():
        return LLMEngine(str(work), max_seq_len=max_seq_len, max_gpu_num_blocks=blocks, use_cuda_graph=True, quantization=quant)

    eng_a = arm("A", build_ref)

    # ---- C needs arm A's (or a fresh reference) model to copy the weights from ----
    def build_ours_under_ref_engine():
        from lite_llama_b200.executor import ModelRunner
        from lite_llama_b200.executor.decoder import DecoderWeights, LayerWeights
        from lite_llama_b200.executor.model_config import ModelConfig
        from lite_llama_b200.quantization import make_linear_weights
        from safetensors.torch import load_file

        geom = GEOM[a.model]
        hd = geom.get("head_dim", geom["hidden_size"] // geom["num_attention_heads"])
        cfg = ModelConfig(a.model, geom["hidden_size"], a.layers, geom["num_attention_heads"], geom["num_key_value_heads"],
                          hd, geom["intermediate_size"], geom["vocab_size"], rope_theta=geom["rope_theta"],
                          rms_norm_eps=geom["rms_norm_eps"], qkv_bias=geom["model_type"] == "qwen2", max_seq_len=max_seq_len)
        head = load_file(str(work / "model-head.safetensors"), device="cuda")
        w = DecoderWeights(embed=head["model.embed_tokens.weight"], lm_head=head["lm_head.weight"],
                           final_norm=head["model.norm.weight"])
        method = quant or "fp16"
        for i in range(a.layers):
            t = load_file(str(work / f"model-layer{i:03d}.safetensors"), device="cuda")
            p = f"model.layers.{i}."
            bias = None
            if cfg.qkv_bias:
                bias = torch.cat([t[p + "self_attn.q_proj.bias"], t[p + "self_attn.k_proj.bias"], t[p + "self_attn.v_proj.bias"]])
            qkv = torch.cat([t[p + "self_attn.q_proj.weight"], t[p + "self_attn.k_proj.weight"], t[p + "self_attn.v_proj.weight"]])
            gu = torch.cat([t[p + "mlp.gate_proj.weight"], t[p + "mlp.up_proj.weight"]])

            def pack(x, b=None):
                return make_linear_weights(x if method == "fp16" else x.float(), method, b, 128)

            w.layers.append(LayerWeights(input_norm=t[p + "input_layernorm.weight"], post_norm=t[p + "post_attention_layernorm.weight"],
                                         qkv=pack(qkv, bias), o=pack(t[p + "self_attn.o_proj.weight"]), gate_up=pack(gu),
                                         down=pack(t[p + "mlp.down_proj.weight"])))
            del t
        runner = ModelRunner(cfg, w, blocks, device="cuda")
        runner.enable_cuda_graph(batch_sizes=tuple(sorted(set(batches + ([runner.graph_batch_size(1)] if False else [])))),
                                 seq_len_buckets=tuple(b for b in (256, 512, 1024, 2048) if b <= max_seq_len))
        # the reference's LLMEngine constructor insists on building its own runner: assemble the same object by hand
        from lite_llama.engine.sampler import Sampler
        from lite_llama.engine.stop_criteria import load_stop_token_ids

        eng = object.__new__(LLMEngine)
        eng.device = "cuda"
        eng.model_path = str(work)
        eng.tensor_parallel_size = 1
        eng.model_runner = runner
        eng.tokenizer = LLMEngine._load_tokenizer(str(work))
        eng.sampler = Sampler()
        eng.max_seq_len = runner.max_seq_len
        eng.stop_token_ids = load_stop_token_ids(str(work), eng.tokenizer)
        eng.last_stop_reasons = None
        return eng

    if eng_a is not None:
        del eng_a
    gc.collect()
    torch.cuda.empty_cache()

    # ---- B: same engine, kernels re-bound ----
    bound = integ.bind()
    result["bound_names"] = bound
    eng_b = arm("B", build_ref)
    if eng_b is not None and a.continuous > 0:
        pass
    del eng_b
    integ.unbind()
    gc.collect()
    torch.cuda.empty_cache()

    eng_c = arm("C", build_ours_under_ref_engine)
    if eng_c is not None and a.continuous > 0:
        try:
            from lite_llama.engine.continuous_engine import ContinuousBatchingEngine
            from lite_llama.engine.scheduler import SchedulerConfig

            eng_c.model_runner.kv_cache_manager.free_all()
            cb = ContinuousBatchingEngine(eng_c, SchedulerConfig(max_seq_len=max_seq_len, max_num_seqs=a.continuous))
            cb.stop_token_ids = set()
            g = torch.Generator().manual_seed(7)
            lens = torch.randint(min(32, max(1, a.prompt_len // 2)), a.prompt_len + 1, (a.continuous,), generator=g).tolist()
            reqs = [cb.add_request("", params, prompt_token_ids=torch.randint(0, V - 2, (n,), generator=g).tolist()) for n in lens]
            t = time.perf_counter()
            steps = 0
            while cb.has_unfinished_requests():
                cb.step()
                steps += 1
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            gen = sum(len(r.output_token_ids) for r in reqs)
            result["arms"]["C"]["continuous"] = {"requests": a.continuous, "engine_steps": steps, "generated_tokens": gen,
                                                 "tok_per_s": round(gen / dt, 1), "seconds": round(dt, 3),
                                                 "finish_reasons": sorted({r.finish_reason for r in reqs})}
        except Exception as ex:
            result["arms"]["C"]["continuous"] = {"error": repr(ex)[:500], "trace": traceback.format_exc()[-2000:]}

    # ---- token-for-token comparison ----
    cmp = {}
    for B in batches:
        ref = grids.get(("A", B))
        for other in ("B", "C"):
            got = grids.get((other, B))
            if ref is None or got is None:
                continue
            same = got == ref
            first_div = [int((~row).nonzero()[0]) if (~row).any() else -1 for row in same]
            cmp[f"{other}_vs_A_bs{B}"] = {
                "identical": bool(same.all()), "rows_identical": int(sum(1 for d in first_div if d < 0)), "rows": B,
                "first_divergence_positions": sorted(d for d in first_div if d >= 0)[:8],
                "tokens_equal_frac": round(float(same[:, a.prompt_len:].float().mean()), 4)}
        gb, gc_ = grids.get(("B", B)), grids.get(("C", B))
        if gb is not None and gc_ is not None:
            cmp[f"C_vs_B_bs{B}"] = {"identical": bool((gb == gc_).all()),
                                    "tokens_equal_frac": round(float((gb == gc_)[:, a.prompt_len:].float().mean()), 4)}
    result["token_parity"] = cmp
    for B in batches:
        ta = result["arms"].get("A", {}).get(f"bs{B}")
        for other in ("B", "C"):
            to = result["arms"].get(other, {}).get(f"bs{B}")
            if ta and to:
                result.setdefault("speedup_vs_reference_triton", {})[f"{other}_bs{B}"] = round(
                    to["tok_per_s_incl_prefill"] / ta["tok_per_s_incl_prefill"], 2)
    shutil.rmtree(work, ignore_errors=True)
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(result, indent=1))
    print(json.dumps(result), flush=True)


if __name__ == "__main__":
    main()
