#!/usr/bin/env python
"""
GPU comparator of BASELINE.md §4 row 2 — NOT a pytest module and NOT part of the product or of bench.py's contract.

Times the "reference HF / flash_attn bf16 path" on the GPU box: the torch-eager restatement under oracle/ (pure-torch ViTs +
projector + the container's transformers.LlamaForCausalLM with attn_implementation="flash_attention_2") driven by HF
`GenerationMixin.generate(do_sample=False)` from `inputs_embeds`, i.e. what PrismaticForConditionalGeneration.generate does
(/root/reference/prismatic/extern/hf/modeling_prismatic.py:362-415, :519), on the workload of bench.py (BASELINE.json
configs[1]: 224x224 image, 40-id prompt, 512 new tokens, same seeded weights). This is the denominator of the north_star's
">= 15x the reference flash_attn bf16 generate_actions throughput" target. Lives under tests/ because only tests/ may import
oracle/ besides bench.py's CPU legs.

  python tests/gpu_comparator.py [--steps 3] [--warmup 2] [--new 512] [--attn flash_attention_2|sdpa|eager]
prints one JSON line (actions/s, ms per token p50, achieved decode GB/s from the same algorithmic byte count as bench.py).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--new", type=int, default=bench.N_NEW)
    ap.add_argument("--attn", default="flash_attention_2")
    args = ap.parse_args()
    from emmax_b200 import PrismaticImageProcessor
    from oracle.model import OracleVLA

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg, tok, sd, script = bench.build_weights(dev)
    oracle = OracleVLA.from_state_dict(cfg, sd, device=dev, dtype=torch.bfloat16, attn_implementation=args.attn)
    del sd
    image, ids = bench.synthetic_request(0)
    pv = PrismaticImageProcessor()(image, return_tensors="pt")["pixel_values"].to(torch.bfloat16).pin_memory()
    ids = ids.pin_memory()
    lm = oracle.language_model
    how = "transformers GenerationMixin.generate(inputs_embeds, do_sample=False)"

    def request():
        d_ids, d_pv = ids.to(dev, non_blocking=True), pv.to(dev, non_blocking=True)
        x, _ = oracle.multimodal_embeddings(d_ids, d_pv)
        mask = torch.ones(x.shape[:2], dtype=torch.long, device=dev)
        with torch.inference_mode():
            out = lm.generate(inputs_embeds=x, attention_mask=mask, do_sample=False, max_new_tokens=args.new, min_new_tokens=args.new,
                              pad_token_id=cfg.text_config.pad_token_id)  # fmt: skip
        return out[0].cpu().tolist()

    try:
        new = request()
    except Exception as e:  # HF generate refuses something in this transformers version: use the oracle's own greedy loop
        how = f"oracle greedy loop (HF generate failed: {type(e).__name__})"

        def request():  # noqa: F811
            d_ids, d_pv = ids.to(dev, non_blocking=True), pv.to(dev, non_blocking=True)
            out = oracle.generate(d_ids, d_pv, args.new, eos_token_id=None)
            return out[0, d_ids.shape[1] :].cpu().tolist()

        new = request()
    want = script[: args.new]
    agree = sum(int(a == b) for a, b in zip(new, want))
    for _ in range(max(args.warmup - 1, 0)):
        request()
    torch.cuda.synchronize()
    times = []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        request()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = statistics.median(times)
    S = ids.shape[1] + cfg.num_patches
    avg_bytes = sum(bench.decode_bytes(cfg, S + j) for j in range(args.new - 1)) / max(args.new - 1, 1)
    line = {
        "impl": "reference-gpu", "what": f"oracle restatement on cuda:0, bf16, attn={args.attn}, {how}",
        "metric": "actions/sec (7-DoF)", "value": 1e3 / ms, "unit": "actions/s", "ms_per_step": ms, "steps": args.steps,
        "warmup": args.warmup, "new_tokens": args.new, "ms_per_token_incl_prefill": ms / args.new,
        "decode_gbs_upper_bound": avg_bytes / (ms / args.new * 1e-3) / 1e9,
        "token_agreement_with_script": f"{agree}/{len(want)}", "gpu": torch.cuda.get_device_name(0),
        "cpu_threads": torch.get_num_threads(),
    }  # fmt: skip
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
