"""HF transformers on the host CPU cores: BASELINE.json configs[0] (Qwen2.5-0.5B, greedy, single prompt).

Restates the measurement of the reference's ``benchmarks/bench_hf_baseline.py`` / ``benchmarks/common.py:169-206``
(``HFBackend.measure``: greedy, ``min_new_tokens == max_new_tokens``, TPS = generated tokens / wall time) with what
exists offline: a random-init model of the public Qwen2.5-0.5B geometry, token ids instead of text, ``device="cpu"``
and fp32 (the reference's script hard-codes ``.cuda()``).  A reported baseline, not an optimisation target.

    python tools/bench_hf_baseline.py [--prompt-len 32] [--new-tokens 16]
"""

from __future__ import annotations

import argparse
import json
import os
import time


def measure(prompt_len: int = 32, new_tokens: int = 16, seed: int = 0) -> dict:
    import torch
    import transformers

    torch.manual_seed(seed)
    cfg = transformers.Qwen2Config(
        vocab_size=151936, hidden_size=896, intermediate_size=4864, num_hidden_layers=24, num_attention_heads=14,
        num_key_value_heads=2, rms_norm_eps=1e-6, rope_theta=1e6, max_position_embeddings=4096, tie_word_embeddings=True)
    t0 = time.time()
    model = transformers.Qwen2ForCausalLM(cfg).float().eval()
    build_s = time.time() - t0
    ids = torch.randint(0, cfg.vocab_size, (1, prompt_len))
    gen = dict(do_sample=False, max_new_tokens=new_tokens, min_new_tokens=new_tokens, use_cache=True, pad_token_id=0)
    with torch.no_grad():
        model.generate(ids, **{**gen, "max_new_tokens": 4, "min_new_tokens": 4})      # warm-up (common.py:183-190)
        t0 = time.time()
        out = model.generate(ids, **gen)
        wall = time.time() - t0
    n = int(out.shape[1] - prompt_len)
    return {"metric": "decode_tokens_per_s", "value": n / wall, "unit": "tok/s", "impl": "hf-transformers-cpu",
            "model": "qwen2.5-0.5b (random init)", "dtype": "f32", "batch": 1, "prompt_len": prompt_len,
            "new_tokens": n, "wall_s": wall, "threads": torch.get_num_threads(), "logical_cores": os.cpu_count(),
            "build_s": build_s, "transformers": transformers.__version__}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--prompt-len", type=int, default=32)
    ap.add_argument("--new-tokens", type=int, default=16)
    a = ap.parse_args()
    print(json.dumps(measure(a.prompt_len, a.new_tokens)), flush=True)
