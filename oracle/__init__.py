"""
oracle/ — TEST INFRASTRUCTURE ONLY.

CPU/GPU torch-eager restatement of the reference `generate_actions` path, used as the checker by `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`. Nothing under
`emmax_b200/` imports it; the product path has no CPU fallback and raises when `libemmax.so` is missing.

What it restates (all paths under /root/reference):
  * prismatic/extern/hf/modeling_prismatic.py:63-158   vision-backbone wiring, fused-GELU projector
  * prismatic/extern/hf/modeling_prismatic.py:362-415  multimodal sequence assembly; :325-341 cached step
  * prismatic/extern/hf/modeling_prismatic.py:495-537  de-tokenise + un-normalise
  * prismatic/extern/hf/processing_prismatic.py:128-145 image transform
  * prismatic/vla/action_tokenizer.py:28-68, prismatic/vla/solver.py:42-137

Third-party arithmetic that is NOT in /root/reference (pins from requirements-min.txt:1-5, README.md:76):
  * timm==0.9.10 `VisionTransformer` — restated from its published semantics in `oracle/vit.py`
  * transformers==4.40.1 `LlamaForCausalLM` — the container's transformers 5.5.0 class is used directly (same math)
  * flash-attn==2.5.5 — container has 2.8.3; used on the GPU box via `attn_implementation="flash_attention_2"`

PARITY PINNING: the reference has no tests, fixtures or golden vectors for the neural path and the `prismatic` package
cannot be imported here (no timm/draccus/tensorflow; gated tokenizer). Its HF model FILE can be executed, though: `gen_golden_hf_model.py`
loads prismatic/extern/hf/modeling_prismatic.py by path and runs the reference's own `OpenVLAForActionPrediction` (CPU, fp32, toy widths;
timm stood in by the ViT restatement, two transformers-5.x shims) and freezes its outputs in tests/golden/hf_model_golden.npz — the oracle's
wiring (state-dict names, backbone split/concat, projector, sequence assembly, cached step, predict_action) is pinned against that.
What stays **parity unpinned**: timm's ViT block internals, transformers 4.40.1 vs 5.5.0 Llama, flash-attn 2.5.5 vs 2.8.3, full size / bf16.
The ViT restatement is additionally cross-checked against an independent implementation of the same architectures
(`transformers.Dinov2WithRegistersModel`, `transformers.SiglipVisionModel`; tests/test_oracle_vit_crosscheck.py).
The image processor (`gen_golden_processor.py`), the prompt builder (`gen_golden_prompts.py`) and the SimplerEnv policy post-processing
(`gen_golden_simpler.py`) are pinned the same way as the integer/fp64 de-tokeniser and the Solver text parser: `oracle/gen_golden.py` imports
`prismatic/vla/action_tokenizer.py` and executes the `Solver` class source of `prismatic/vla/solver.py` from
/root/reference and freezes their outputs in `tests/golden/detok_golden.json`.
"""
