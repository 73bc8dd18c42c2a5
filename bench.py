#!/usr/bin/env python
"""
bench.py — headline benchmark of the `generate_actions` hot path (BASELINE.json: actions/sec at bs=1).

A "step" is one complete action request: one synthetic 224x224 image + a fixed 40-id prompt -> DINOv2+SigLIP ViT
-> projector -> Llama-2-7B prefill -> greedy decode of 512 reasoning+action tokens -> action de-tokenise
(BASELINE.json configs[1]; full Emma-X architecture, seeded random-init bf16 weights, EOS planted at token 512).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (N>1: torchrun, one replica per GPU + one NCCL
                                                               all-gather of the 7 action tokens per step)
  python bench.py --impl reference [...]                       the reference's own CPU path (torch-eager oracle port),
                                                               bounded sample, host cores only
  python bench.py --impl reference-gpu [...]                   the reference HF path on THIS GPU: torch-eager restatement +
                                                               transformers Llama with flash-attn, HF generate
  python bench.py --config c5 [--gpus N]                       BASELINE.json configs[4]: 8 sequences per GPU in ONE batched decode
                                                               (emx_decode_batch_step), half 128 / half 512 new tokens
  python bench.py --config c3                                  BASELINE.json configs[2]: bs=32 ViT + prompt prefill, 1 new token
                                                               (tensor-core roofline probe)
The default run (configs[1], the headline) also carries short c5 / c3 / GPU-comparator legs under "extras" (N=1 only).

`value`  : actions/s with inputs already resident in HBM (engine.generate on device tensors).
`e2e`    : actions/s through the public API (`model.generate_actions(inputs, tokenizer, ...)`) from PINNED HOST
           buffers: H2D of the raw uint8 frame + ids, GPU image transform, D2H of the generated ids, host text decode +
           Solver parse, every step.
`roofline`: the decode-step kernel (99 % of a step): algorithmic bytes per launch / measured launch duration vs the
           measured HBM peak in MEASURED_PEAKS.json.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_NEW = 512
PROMPT_LEN = 40
WORKLOAD = "bs=1 single-GPU bf16, 224x224 image, max_new_tokens=512 grounded-CoT decode (BASELINE.json configs[1])"


def synthetic_request(seed: int):
    """image = default_rng(seed).integers(0,256,(224,224,3)); prompt = [1] + rng(1234).integers(3,31744,39)."""
    from PIL import Image

    image = Image.fromarray(np.random.default_rng(seed).integers(0, 256, (224, 224, 3), dtype=np.uint8))
    ids = [1] + np.random.default_rng(1234).integers(3, 31744, PROMPT_LEN - 1).tolist()
    return image, torch.tensor([ids], dtype=torch.long)


def build_weights(device):
    from emmax_b200 import SyntheticLlamaTokenizer, emma_x_config
    from emmax_b200.synthetic import default_script, make_state_dict

    cfg = emma_x_config()
    tok = SyntheticLlamaTokenizer()
    _, ids = synthetic_request(0)
    script = default_script(tok, N_NEW, seed=0)
    sd = make_state_dict(cfg, seed=0, device=device, script=script, script_prev=int(ids[0, -1]))
    return cfg, tok, sd, script


def decode_bytes(cfg, ctx: int) -> int:
    """Algorithmic bytes of one decode step at context `ctx` (tokens already cached): SURVEY.md §8d / BASELINE.md §3."""
    t = cfg.text_config
    H, I, L, V = t.hidden_size, t.intermediate_size, t.num_hidden_layers, t.vocab_size
    weights = 2 * (L * (4 * H * H + 3 * H * I) + V * H)
    kv_row = 2 * 2 * L * H  # K and V, bf16, all layers
    return weights + kv_row * ctx + kv_row


def c2_config(world: int) -> dict:
    """The headline configuration block, identical in our arm and in the reference arms (the driver compares them)."""
    return {"workload": WORKLOAD, "prompt_ids": PROMPT_LEN, "prefill_positions": PROMPT_LEN + 256, "new_tokens": N_NEW,
            "weights": "seeded random-init, full Emma-X architecture (DINOv2-L/14-reg4 + SigLIP-so400m/14 + Llama-2-7B)",
            "parallelism": f"replicas x{world}" + (" + 1 NCCL all-gather of action tokens per step" if world > 1 else ""),
            "l2": "per-step inputs (13.2 GB of weights streamed per token) exceed the 126 MB L2; no flush needed"}  # fmt: skip


def batch_decode_bytes(cfg, ctxs) -> int:
    """One batched decode launch over the sequences with cached contexts `ctxs`: the weights once + every sequence's KV (SURVEY.md §8d:
    13,214,679,040 + sum_i 524,288 ctx_i + b 524,288)."""
    return decode_bytes(cfg, 0) - 2 * 2 * cfg.text_config.num_hidden_layers * cfg.text_config.hidden_size + sum(
        2 * 2 * cfg.text_config.num_hidden_layers * cfg.text_config.hidden_size * (c + 1) for c in ctxs)


def prefill_flops(cfg, S: int) -> dict:
    """Algorithmic FLOPs of vision + projector + LLM prefill for ONE image and S prefill positions (405.2 GFLOP/image
    for the blocks whose output is used; 2 * 6,476,005,376 * S + 2 * S^2 * 4096 * 32 for the LLM, + lm_head on the last row)."""
    t = cfg.text_config
    P = cfg.num_patches

    def vit(v):
        D, T, M, hd, nh = v.embed_dim, v.num_tokens, v.mlp_dim, v.head_dim, v.num_heads
        return v.used_depth * (2 * T * D * 3 * D + 2 * T * D * D + 4 * T * D * M + 4 * T * T * hd * nh) + 2 * P * D * 3 * v.patch_size * v.patch_size

    vd, H, L, I, V = cfg.vision_embed_dim, t.hidden_size, t.num_hidden_layers, t.intermediate_size, t.vocab_size
    out = {"vision": sum(vit(v) for v in cfg.vision_dims), "projector": 2 * P * (vd * 4 * vd + 4 * vd * H + H * H),
           "llm": 2 * S * L * (4 * H * H + 3 * H * I) + 2 * S * S * H * L + 2 * V * H}  # fmt: skip
    out["total"] = out["vision"] + out["projector"] + out["llm"]
    return out


C5_LIMITS = [128, 512, 128, 512, 128, 512, 128, 512]  # BASELINE.json configs[4]: 8 sequences per GPU, mixed 128/512 max_new_tokens
C5_WORKLOAD = "bs=64 across 8xB200 = 8 sequences per GPU in one batched decode, mixed 128/512 max_new_tokens (BASELINE.json configs[4])"
C3_WORKLOAD = "bs=32 single-GPU bf16 prefill-heavy (ViT+prompt only, 1 new token) - tensor-core roofline probe (BASELINE.json configs[2])"


def read_peaks():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        d = json.load(open(peaks_path))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", 1388.5)), "MEASURED_PEAKS.json"
    return 6650.0, 1390.0, "B200_PROFILING.md fallback"


class ClockSampler(threading.Thread):
    QUERY = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int) -> None:
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()

    def run(self) -> None:
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")  # fmt: skip
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self) -> dict:
        self._halt.set()
        self.join(timeout=6)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}  # fmt: skip


# =====================================================================================================================
# CPU reference arm (oracle port), bounded sample
# =====================================================================================================================
def cpu_reference_sample(cfg, sd, n_decode: int = 4):
    """One bounded sample of the workload on the host cores with the torch-eager oracle (bf16 weights, sdpa):
    vision + projector + prefill (S = 296) once, then `n_decode` cached decode steps; actions/s is extrapolated to the
    512-token request as 1 / (t_prefill + 511 * t_token)."""
    from oracle.model import OracleVLA

    oracle = OracleVLA.from_state_dict(cfg, sd, device="cpu", dtype=torch.bfloat16, attn_implementation="sdpa")
    image, ids = synthetic_request(0)
    from emmax_b200 import PrismaticImageProcessor

    pv = PrismaticImageProcessor()(image, return_tensors="pt")["pixel_values"].to(torch.bfloat16)
    t0 = time.perf_counter()
    logits, past = oracle.prefill(ids, pv)
    tok = int(torch.argmax(logits[:, -1], dim=-1)[0])
    t1 = time.perf_counter()
    for _ in range(n_decode):
        logits, past = oracle.step(torch.tensor([[tok]]), past)
        tok = int(torch.argmax(logits[:, -1], dim=-1)[0])
    t2 = time.perf_counter()
    t_prefill, t_token = t1 - t0, (t2 - t1) / n_decode
    return oracle, {"t_prefill_s": t_prefill, "t_token_s": t_token, "actions_per_s": 1.0 / (t_prefill + (N_NEW - 1) * t_token)}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 for every rank; the reference arm is ONE process that may use every host core
    torch.set_num_threads(os.cpu_count() or 1)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    cfg, tok, sd, _ = build_weights(dev)
    sd = {k: v.cpu() for k, v in sd.items()}
    n_decode = 4
    oracle, first = cpu_reference_sample(cfg, sd, n_decode)
    # keep the whole run bounded: each step is one sample (prefill + n_decode tokens)
    per_step = first["t_prefill_s"] + n_decode * first["t_token_s"]
    budget_s = 240.0
    steps = max(1, min(args.steps, int(budget_s / per_step) - 1))
    warm = max(0, min(args.warmup, int(budget_s / per_step) - 1 - steps))
    from emmax_b200 import PrismaticImageProcessor

    image, ids = synthetic_request(0)
    pv = PrismaticImageProcessor()(image, return_tensors="pt")["pixel_values"].to(torch.bfloat16)
    samples = []
    for i in range(warm + steps):
        t0 = time.perf_counter()
        logits, past = oracle.prefill(ids, pv)
        tk = int(torch.argmax(logits[:, -1], dim=-1)[0])
        t1 = time.perf_counter()
        for _ in range(n_decode):
            logits, past = oracle.step(torch.tensor([[tk]]), past)
            tk = int(torch.argmax(logits[:, -1], dim=-1)[0])
        t2 = time.perf_counter()
        if i >= warm:
            samples.append((t1 - t0, (t2 - t1) / n_decode))
    if not samples:
        samples = [(first["t_prefill_s"], first["t_token_s"])]
    tp = statistics.median(s[0] for s in samples)
    tt = statistics.median(s[1] for s in samples)
    step_s = tp + (N_NEW - 1) * tt
    value = 1.0 / step_s
    cores = torch.get_num_threads()
    sample = (f"torch-eager oracle port on CPU, bf16 weights, sdpa: vision+prefill(S=296) + {n_decode} cached decode tokens per step, "
              f"{len(samples)} timed steps ({warm} warm-up; requested {args.steps}/{args.warmup}); extrapolated to the 512-token request: "
              f"t_prefill={tp:.2f}s, t_token={tt * 1e3:.0f}ms; os.cpu_count()={os.cpu_count()}")  # fmt: skip
    line = {
        "impl": "reference", "metric": "actions/sec (7-DoF)", "value": value, "unit": "actions/s", "n_gpus": args.gpus,
        "steps": len(samples), "warmup": warm, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": c2_config(args.gpus), "device": "host CPU",
        "cpu_baseline": {"value": value, "unit": "actions/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "actions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


# =====================================================================================================================
# parity gate ("parity gate before any timing counts")
# =====================================================================================================================
def parity_gate(cfg, sd, script, eng, dev, with_oracle: bool, n_forced: int = 8) -> dict:
    """Runs BEFORE anything is timed and raises if the CUDA path is wrong:
    (1) the full 512-token greedy continuation of the bench request must equal the planted script, id for id;
    (2) (rank 0) prefill + the first `n_forced` teacher-forced decode steps: every step's logits within 1e-2 * max|logit| of the
        torch-eager oracle (transformers Llama + flash-attn / sdpa) run on this GPU with the same weights, and the same argmax.
    The oracle is the checker only; it is freed before the timed region."""
    image, ids = synthetic_request(0)
    from emmax_b200 import PrismaticImageProcessor

    pv = PrismaticImageProcessor()(image, return_tensors="pt")["pixel_values"].to(dev, torch.bfloat16)
    ids = ids.to(dev)
    new, _ = eng.generate(ids, pv, N_NEW, eos_token_id=2)
    got = new.cpu().tolist()
    if got != list(script):
        bad = next(i for i, (a, b) in enumerate(zip(got, script)) if a != b) if len(got) == len(script) else min(len(got), len(script))
        raise SystemExit(f"PARITY GATE FAILED: greedy ids differ from the planted script at token {bad} ({len(got)} generated)")
    out = {"ids_equal_script": True, "n_ids": len(got), "oracle": None}
    if with_oracle:
        from oracle.model import OracleVLA

        try:
            import flash_attn  # noqa: F401

            attn = "flash_attention_2"
        except Exception:
            attn = "sdpa"
        forced = list(script[:n_forced])
        _, logits = eng.generate(ids, pv, n_forced, eos_token_id=None, return_logits=True, forced=forced)
        oracle = OracleVLA.from_state_dict(cfg, sd, device=dev, dtype=torch.bfloat16, attn_implementation=attn)
        _, logits_o = oracle.generate(ids, pv, n_forced, eos_token_id=None, return_logits=True, forced=forced)
        del oracle
        lo = logits_o.float()
        err = float(((logits.cpu().float() - lo).abs().max() / lo.abs().max()).item())
        same_argmax = bool((logits.cpu().argmax(-1) == lo.argmax(-1)).all())
        if not (err < 1e-2 and same_argmax):
            raise SystemExit(f"PARITY GATE FAILED: teacher-forced logits vs the oracle: rel err {err:.3g} (tolerance 1e-2), argmax equal: {same_argmax}")
        out["oracle"] = {"kind": f"torch-eager oracle on this GPU ({attn})", "steps": n_forced, "logits_rel_err": err, "tolerance": 1e-2,
                         "argmax_equal": same_argmax}  # fmt: skip
    return out


# =====================================================================================================================
# GPU comparator: the reference HF / flash-attn bf16 path on the same GPU (north_star's >= 15x denominator)
# =====================================================================================================================
def gpu_comparator_sample(cfg, sd, script, dev, steps: int = 2, warmup: int = 1, n_new: int = N_NEW, attn: str = "flash_attention_2") -> dict:
    """The torch-eager restatement under oracle/ (pure-torch ViTs + projector + the container's transformers.LlamaForCausalLM with
    flash-attn) driven by HF `GenerationMixin.generate(do_sample=False)` from `inputs_embeds`, i.e. what
    PrismaticForConditionalGeneration.generate does (modeling_prismatic.py:362-415, :519), on this bench's request and weights, from
    pinned host inputs. None of this repo's kernels run here. Returns actions/s and ms per token."""
    from emmax_b200 import PrismaticImageProcessor
    from oracle.model import OracleVLA

    try:
        import flash_attn  # noqa: F401
    except Exception:
        attn = "sdpa"
    oracle = OracleVLA.from_state_dict(cfg, sd, device=dev, dtype=torch.bfloat16, attn_implementation=attn)
    image, ids = synthetic_request(0)
    pv = PrismaticImageProcessor()(image, return_tensors="pt")["pixel_values"].to(torch.bfloat16).pin_memory()
    ids = ids.pin_memory()
    lm = oracle.language_model
    how = "transformers GenerationMixin.generate(inputs_embeds, do_sample=False)"

    def request_hf():
        d_ids, d_pv = ids.to(dev, non_blocking=True), pv.to(dev, non_blocking=True)
        x, _ = oracle.multimodal_embeddings(d_ids, d_pv)
        mask = torch.ones(x.shape[:2], dtype=torch.long, device=dev)
        with torch.inference_mode():
            out = lm.generate(inputs_embeds=x, attention_mask=mask, do_sample=False, max_new_tokens=n_new, min_new_tokens=n_new,
                              pad_token_id=cfg.text_config.pad_token_id)  # fmt: skip
        return out[0].cpu().tolist()

    def request_loop():
        d_ids, d_pv = ids.to(dev, non_blocking=True), pv.to(dev, non_blocking=True)
        out = oracle.generate(d_ids, d_pv, n_new, eos_token_id=None)
        return out[0, d_ids.shape[1] :].cpu().tolist()

    request = request_hf
    try:
        new = request()
    except Exception as e:  # HF generate refuses something in this transformers version: the oracle's own greedy loop
        how = f"oracle greedy loop (HF generate failed: {type(e).__name__})"
        request = request_loop
        new = request()
    want = list(script[:n_new])
    agree = sum(int(a == b) for a, b in zip(new, want))
    for _ in range(max(warmup - 1, 0)):
        request()
    torch.cuda.synchronize()
    times = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        request()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    del oracle
    torch.cuda.empty_cache()
    ms = statistics.median(times)
    return {"what": f"oracle restatement of the reference HF path on this GPU, bf16, attn={attn}, {how}; none of this repo's kernels",
            "value": 1e3 / ms, "unit": "actions/s", "ms_per_step": ms, "steps": steps, "warmup": warmup, "new_tokens": n_new,
            "ms_per_token_incl_prefill": ms / n_new, "token_agreement_with_script": f"{agree}/{len(want)}"}  # fmt: skip


def run_reference_gpu(args) -> None:
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    cfg, tok, sd, script = build_weights(dev)
    sampler = ClockSampler(dev.index)
    sampler.start()
    r = gpu_comparator_sample(cfg, sd, script, dev, steps=max(1, min(args.steps, 5)), warmup=max(1, min(args.warmup, 2)))
    line = {"impl": "reference-gpu", "metric": "actions/sec (7-DoF)", "value": r["value"], "unit": "actions/s", "n_gpus": 1,
            "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": c2_config(args.gpus), "device": torch.cuda.get_device_name(dev),
            "clocks": sampler.stop(), "gpu_comparator": r, "gpu_launches": 0,
            "e2e": {"value": r["value"], "unit": "actions/s", "h2d_bytes_per_step": 6 * 224 * 224 * 2 + PROMPT_LEN * 8, "d2h_bytes_per_step": (PROMPT_LEN + N_NEW) * 8}}  # fmt: skip
    print(json.dumps(line), flush=True)


# =====================================================================================================================
# BASELINE.json configs[4] (8 sequences per GPU, one batched decode) and configs[2] (bs=32 prefill probe)
# =====================================================================================================================
def c5_requests(proc, rank: int, world: int, step: int):
    """8 frames of the synthetic stream for this rank and step (same fixed prompt): raw uint8 frames [8, 224, 224, 3] + ids [8, 40]."""
    base = (step * world + rank) * len(C5_LIMITS)
    frames, ids = [], None
    for j in range(len(C5_LIMITS)):
        image, ids = synthetic_request(base + j)
        frames.append(torch.from_numpy(np.asarray(image).copy()))
    return torch.stack(frames), ids.repeat(len(C5_LIMITS), 1)


def measure_c5(cfg, model, proc, script, dev, world: int, rank: int, K: int, W: int, timed, tick_gather) -> dict:
    """One step = 8 complete requests per GPU served by ONE batched decode (one pass over the weights per token for all of them); the
    reference needs 8 sequential bs=1 generations (its cached branch asserts bs == 1). Parity first: every sequence's ids must equal
    the planted script up to its own limit."""
    from emmax_b200 import _lib

    eng = model.engine
    B = len(C5_LIMITS)
    reqs = [c5_requests(proc, rank, world, i) for i in range(W + K)]
    d_reqs = [(proc.image_processor.preprocess_device(fr.to(dev)), ids.to(dev)) for fr, ids in reqs]
    h_reqs = [(fr.pin_memory(), ids.pin_memory()) for fr, ids in reqs]
    new, _ = eng.generate_batch(d_reqs[0][1], d_reqs[0][0], C5_LIMITS, eos_token_id=2)
    for b, lim in enumerate(C5_LIMITS):
        if new[b].cpu().tolist() != list(script[:lim]):
            raise SystemExit(f"PARITY GATE FAILED (c5): sequence {b} (limit {lim}) differs from the planted script")
    acc = {"ms": 0.0, "launches": 0, "bytes": 0}

    def step_device(i: int) -> None:
        pv, ids = d_reqs[i]
        new, _ = eng.generate_batch(ids, pv, C5_LIMITS, eos_token_id=2)
        tick_gather(new)
        ev0, ev1, n, S, limits = eng.last_decode_batch
        if i >= W:
            acc["ms"] += ev0.elapsed_time(ev1)
            acc["launches"] += n
            # launch j (0-based) runs the sequences that still have tokens to produce: n_generated = j + 1 < limit
            acc["bytes"] += sum(batch_decode_bytes(cfg, [S[b] + j for b in range(B) if j + 1 < limits[b]]) for j in range(n))

    total_ms, launches = timed(step_device, W, K)
    value = world * K * B / (total_ms / 1e3)
    last = [None]

    def step_e2e(i: int) -> None:
        frames, ids = h_reqs[i]
        pv = proc.image_processor.preprocess_device(frames.to(dev, non_blocking=True))
        d_ids = ids.to(dev, non_blocking=True)
        inputs = [{"input_ids": d_ids[b : b + 1], "pixel_values": pv[b : b + 1]} for b in range(B)]
        last[0] = model.generate_actions_batch(inputs, proc.tokenizer, max_new_tokens=C5_LIMITS)

    e2e_ms, _ = timed(step_e2e, W, K)
    hbm_peak, _, peak_src = read_peaks()
    btpath = os.path.join(ROOT, "profiles", "decode_batch_traffic.json")
    btraffic = json.load(open(btpath)) if os.path.exists(btpath) else {}
    avg_ms = acc["ms"] / max(acc["launches"], 1)
    per_launch = acc["bytes"] / max(acc["launches"], 1)
    achieved = per_launch / (avg_ms * 1e-3) / 1e9

    # ---- continuous batching: the SAME K x 8 requests as one stream through the 8 sequence slots (Engine.serve): a slot whose sequence
    # has produced its 128 tokens is refilled with the next request instead of idling until the 512-token sequences of its batch finish
    stream = [(d_reqs[i][1][b : b + 1], d_reqs[i][0][b : b + 1], C5_LIMITS[b]) for i in range(W, W + K) for b in range(B)]
    warm = [(d_reqs[0][1][b : b + 1], d_reqs[0][0][b : b + 1], 8 if b % 2 else 4) for b in range(B)] * 2  # short: touches every code path once
    outs = [None]

    def step_stream(i: int) -> None:
        outs[0] = eng.serve(stream if i >= 1 else warm, eos_token_id=2, use_graph=False)
        if i >= 1:
            for g0 in range(0, len(stream), B):
                tick_gather(outs[0][g0 : g0 + B])

    stream_ms, stream_launches = timed(step_stream, 1, 1)
    for r, new_r in enumerate(outs[0]):
        if new_r.cpu().tolist() != list(script[: stream[r][2]]):
            raise SystemExit(f"PARITY GATE FAILED (c5, continuous batching): request {r} (limit {stream[r][2]}) differs from the planted script")
    ls = eng.last_serve
    kv_tok = 2 * 2 * cfg.text_config.num_hidden_layers * cfg.text_config.hidden_size
    s_bytes = ls["launches"] * (decode_bytes(cfg, 0) - kv_tok) + kv_tok * (ls["kv_reads"] + ls["kv_writes"])
    s_dec_ms = sum(e0.elapsed_time(e1) for e0, e1 in ls["decode_events"])
    s_achieved = s_bytes / max(s_dec_ms, 1e-9) / 1e6
    # ---- the same stream admitted longest-first (token limits are known up front in an offline batch): the tail of the stream is then
    # made of 128-token requests and the slots drain together (emmax_b200.engine.admission_order)
    def step_stream_lpt(i: int) -> None:
        outs[0] = eng.serve(stream if i >= 1 else warm, eos_token_id=2, use_graph=False, order="longest_first")
        if i >= 1:
            for g0 in range(0, len(stream), B):
                tick_gather(outs[0][g0 : g0 + B])

    lpt_ms, _ = timed(step_stream_lpt, 1, 1)
    for r, new_r in enumerate(outs[0]):
        if new_r.cpu().tolist() != list(script[: stream[r][2]]):
            raise SystemExit(f"PARITY GATE FAILED (c5, continuous batching, longest first): request {r} differs from the planted script")
    lpt = {"value": world * len(stream) / (lpt_ms / 1e3), "unit": "actions/s", "total_ms": lpt_ms, "decode_launches": eng.last_serve["launches"],
           "over_static_batches": (world * len(stream) / (lpt_ms / 1e3)) / value}  # fmt: skip
    continuous = {
        "what": "the same K x 8 requests served as ONE stream through the 8 sequence slots (Engine.serve, continuous batching: a finished slot is "
                "refilled with the next request at once); inputs resident in HBM; every request's ids checked against the planted script",
        "value": world * len(stream) / (stream_ms / 1e3), "unit": "actions/s", "requests": len(stream), "total_ms": stream_ms,
        "decode_launches": ls["launches"], "avg_launch_ms": s_dec_ms / max(ls["launches"], 1), "gpu_launches": stream_launches,
        "roofline": {"bound": "hbm", "kernel": "decode_batch_kernel", "achieved": s_achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": s_achieved / hbm_peak, "bytes_per_launch": s_bytes / max(ls["launches"], 1), "share_of_stream": s_dec_ms / stream_ms},
        "over_static_batches": (world * len(stream) / (stream_ms / 1e3)) / value,
        "longest_first": lpt,
    }  # fmt: skip
    return {
        "metric": "actions/sec (7-DoF)", "value": value, "unit": "actions/s", "ms_per_step": total_ms / K, "steps": K, "warmup": W,
        "continuous_batching": continuous,
        "config": {"workload": C5_WORKLOAD, "sequences_per_gpu": B, "max_new_tokens": C5_LIMITS, "prompt_ids": PROMPT_LEN,
                   "prefill_positions": PROMPT_LEN + 256, "parallelism": f"replicas x{world}, 8 sequences per replica in one batched decode kernel",
                   "l2": "13.2 GB of weights + up to 3.4 GB of KV streamed per token exceed the 126 MB L2; no flush needed"},
        "parity": {"ids_equal_script_per_sequence": True, "sequences": B},
        "e2e": {"value": world * K * B / (e2e_ms / 1e3), "unit": "actions/s", "ms_per_step": e2e_ms / K,
                "h2d_bytes_per_step": int(h_reqs[0][0].numel() + h_reqs[0][1].numel() * 8), "d2h_bytes_per_step": 4 * sum(C5_LIMITS) + 4 * B},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "decode_batch_kernel", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": btraffic.get("dram_bytes_per_launch"),
                     "traffic_source": f"profiles/decode_batch_traffic.json: ncu --set full capture of this kernel at batch {btraffic.get('batch')}, context {btraffic.get('context')} "
                                       f"(algorithmic there: {btraffic.get('algorithmic_bytes_at_that_point')} B, ratio {btraffic.get('ratio_to_algorithmic')})",
                     "peak_source": peak_src + " hbm_gbs", "bytes_per_launch": per_launch,
                     "avg_launch_ms": avg_ms, "launches_timed": acc["launches"], "share_of_step": acc["ms"] / total_ms},
        "action_of_sequence_0": [round(float(a), 6) for a in last[0][0][0]],
    }  # fmt: skip


def measure_c3(cfg, model, proc, dev, K: int, W: int, timed) -> dict:
    """One step = vision towers + projector + Llama prefill (S = 296) for 32 images at once, 1 new token each (the first greedy id)."""
    eng = model.engine
    B = 32
    if eng.max_batch < B:
        raise RuntimeError(f"engine max_batch {eng.max_batch} < {B}")
    frames = torch.stack([torch.from_numpy(np.asarray(synthetic_request(100 + j)[0]).copy()) for j in range(B)])
    _, ids1 = synthetic_request(0)
    ids = ids1.repeat(B, 1)
    h_frames, h_ids = frames.pin_memory(), ids.pin_memory()
    d_pv, d_ids = proc.image_processor.preprocess_device(frames.to(dev)), ids.to(dev)

    def step_device(i: int) -> None:
        eng.prefill(d_ids, d_pv)

    total_ms, launches = timed(step_device, W, K)
    first = [None]

    def step_e2e(i: int) -> None:
        pv = proc.image_processor.preprocess_device(h_frames.to(dev, non_blocking=True))
        ws = eng.prefill(h_ids.to(dev, non_blocking=True), pv)
        first[0] = ws["first"][:B].cpu()

    e2e_ms, _ = timed(step_e2e, W, K)
    S = PROMPT_LEN + cfg.num_patches
    fl = prefill_flops(cfg, S)
    _, tc_peak, peak_src = read_peaks()
    achieved = B * fl["total"] / (total_ms / K * 1e-3) / 1e12
    return {
        "metric": "prefill requests/sec (ViT + projector + Llama prefill, 1 new token)", "value": K * B / (total_ms / 1e3), "unit": "requests/s",
        "ms_per_step": total_ms / K, "steps": K, "warmup": W,
        "config": {"workload": C3_WORKLOAD, "batch": B, "prefill_positions": S, "new_tokens": 1,
                   "l2": "activations of 32 x 296 positions (0.7 GB workspace) + 15 GB of weights exceed the 126 MB L2; no flush needed"},
        "e2e": {"value": K * B / (e2e_ms / 1e3), "unit": "requests/s", "ms_per_step": e2e_ms / K,
                "h2d_bytes_per_step": int(h_frames.numel() + h_ids.numel() * 8), "d2h_bytes_per_step": 4 * B},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "gemm_tn_kernel (whole prefill graph)", "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s",
                     "frac": achieved / tc_peak, "traffic": None, "peak_source": peak_src + " bf16_tflops_sustained",
                     "flop_per_step": B * fl["total"], "gflop_per_image": {k: round(v / 1e9, 1) for k, v in fl.items()}},
        "first_tokens": first[0].tolist()[:8],
    }  # fmt: skip


def predict_action_latency(cfg, model, proc, sd, script, dev, n: int = 20, warm: int = 3) -> dict:
    """Latency of the OpenVLA-style entry point `predict_action` (modeling_prismatic.py:506-537: 7 action tokens after the prompt, the call
    experiments/robot/openvla_utils.py:169 and the SimplerEnv policy make every control step): wall clock per call through the public API
    from a pinned host frame (H2D, GPU image transform, prefill graph, 7 decode launches, device de-tokeniser, D2H of the action), next to
    the same 7-token request through the HF-generate + flash-attn restatement on the same GPU."""
    import time

    image, ids = synthetic_request(0)
    frame, h_ids = torch.from_numpy(np.asarray(image).copy()).pin_memory(), ids.pin_memory()

    def call():
        pv = proc.image_processor.preprocess_device(frame.to(dev, non_blocking=True))
        return model.predict_action(input_ids=h_ids.to(dev, non_blocking=True), pixel_values=pv, unnorm_key=None, do_sample=False)

    for _ in range(warm):
        call()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        a = call()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    out = {"what": "predict_action (7 action tokens) per call, wall clock, public API from a pinned host frame", "calls": n,
           "ms_p50": ts[n // 2], "ms_p10": ts[n // 10], "ms_p90": ts[(9 * n) // 10], "hz_p50": 1e3 / ts[n // 2],
           "action": [round(float(x), 6) for x in a]}  # fmt: skip
    try:
        comp = gpu_comparator_sample(cfg, sd, script, dev, steps=5, warmup=2, n_new=7)
        out["gpu_comparator_ms"] = comp["ms_per_step"]
        out["ours_over_comparator"] = comp["ms_per_step"] / ts[n // 2]
    except Exception as e:  # noqa: BLE001
        out["gpu_comparator_ms"] = f"failed: {type(e).__name__}: {e}"
    return out


# =====================================================================================================================
# our arm
# =====================================================================================================================
def run_ours(args) -> None:
    import torch.distributed as dist

    from emmax_b200 import AutoProcessor, OpenVLAForActionPrediction, _lib
    from emmax_b200.replicas import gather_action_tokens, pack_action_tokens

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":  # keeps the version banner off stdout (one JSON line only)
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)

    cfg, tok, sd, script = build_weights(dev)
    cpu_sd = {k: v.cpu() for k, v in sd.items()} if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    extras_on = args.config == "c2" and world == 1 and not args.no_extras
    max_batch = 32 if (args.config == "c3" or extras_on) else 8
    model = OpenVLAForActionPrediction(cfg, dict(sd), max_context=1024, max_batch=max_batch).to(dev)
    eng = model.engine
    proc = AutoProcessor.from_pretrained(None)
    parity = parity_gate(cfg, sd, script, eng, dev, with_oracle=(rank == 0 and not args.no_parity_oracle))
    if not extras_on:
        del sd
    torch.cuda.empty_cache()

    # request stream: rank r serves frames r, r+N, ...; same fixed prompt
    def request(i: int):
        image, ids = synthetic_request(rank + i * world)
        return proc.image_processor(image, return_tensors="pt")["pixel_values"].to(torch.bfloat16), ids, torch.from_numpy(np.asarray(image).copy())

    reqs = [request(i) for i in range(W + K)]
    d_reqs = [(pv.to(dev), ids.to(dev)) for pv, ids, _ in reqs]
    # e2e inputs: the RAW uint8 frame (224x224x3, as the robot loop delivers it) + prompt ids, in pinned host memory
    h_reqs = [(frame.pin_memory(), ids.pin_memory()) for _, ids, frame in reqs]
    gathered = torch.zeros((world, 8), dtype=torch.int32, device=dev)
    act_lo = script.index(tok.key_id("POLICIES:")) + 2  # first policy's 7 action tokens

    def tick_gather(new_tokens: torch.Tensor) -> None:
        """the ONE collective of the path: all-gather of each replica's action tokens, enqueued on the decode stream"""
        if world > 1:
            gather_action_tokens(pack_action_tokens(new_tokens[act_lo : act_lo + 7]), out=gathered)

    def sync_all() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n_warm: int, n: int):
        for i in range(n_warm):
            fn(i)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count
        e0.record()
        for i in range(n_warm, n_warm + n):
            fn(i)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), _lib.launch_count - l0

    base = {"n_gpus": world, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic"}
    if args.config in ("c5", "c3"):
        gathered8 = torch.zeros((world, 8 * len(C5_LIMITS)), dtype=torch.int32, device=dev)

        def tick_gather8(new_list) -> None:
            """all-gather of the 7 action tokens of each of this replica's 8 sequences (8 x 8 int32), on the decode stream"""
            if world > 1:
                mine = torch.zeros(8 * len(C5_LIMITS), dtype=torch.int32, device=dev)
                for b, ids_b in enumerate(new_list):
                    if ids_b.numel() >= act_lo + 7:
                        mine[8 * b : 8 * b + 7] = ids_b[act_lo : act_lo + 7]
                dist.all_gather_into_tensor(gathered8.view(-1), mine)

        sampler = ClockSampler(local)
        sampler.start()
        if args.config == "c5":
            r = measure_c5(cfg, model, proc, script, dev, world, rank, K, W, timed, tick_gather8)
        else:
            if world > 1:
                raise SystemExit("--config c3 is a single-GPU probe")
            r = measure_c3(cfg, model, proc, dev, K, W, timed)
        r["clocks"] = sampler.stop()
        if args.config == "c5" and world > 1:  # sequences with limit 512 carry the planted action tokens; check every replica's rows
            want = torch.tensor(script[act_lo : act_lo + 7], dtype=torch.int32, device=dev)
            rows = gathered8.view(world, len(C5_LIMITS), 8)[:, [b for b, lim in enumerate(C5_LIMITS) if lim >= act_lo + 7], :7]
            assert bool((rows == want).all()), "all-gathered action tokens differ from the script"
            r["parity"]["gathered_rows_checked"] = int(rows.shape[0] * rows.shape[1])
        if rank == 0:
            print(json.dumps({**base, **r}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- device-resident throughput (`value`) + roofline of the decode kernel ------------------------------------
    decode_ms, decode_launches, decode_bytes_total = [0.0], [0], [0]

    def step_device(i: int) -> None:
        pv, ids = d_reqs[i]
        new, _ = eng.generate(ids, pv, N_NEW, eos_token_id=2)
        tick_gather(new)
        ev0, ev1, n, S = eng.last_decode
        if i >= W:
            decode_ms[0] += ev0.elapsed_time(ev1)
            decode_launches[0] += n
            decode_bytes_total[0] += sum(decode_bytes(cfg, S + j) for j in range(n))
        assert new.numel() == N_NEW, f"expected {N_NEW} tokens, got {new.numel()}"

    sampler = ClockSampler(local)
    sampler.start()
    total_ms, launches = timed(step_device, W, K)
    clocks = sampler.stop()
    value = world * K / (total_ms / 1e3)

    # ---- end to end through the public API from pinned host memory (`e2e`) ------------------------------------------
    last_action = [None]

    def step_e2e(i: int) -> None:
        frame, ids = h_reqs[i]
        # H2D of the raw frame + ids, image transform on the GPU (emx_preprocess_u8: bit-exact twin of the host processor)
        inputs = {"input_ids": ids.to(dev, non_blocking=True),
                  "pixel_values": proc.image_processor.preprocess_device(frame.to(dev, non_blocking=True))}
        action, text = model.generate_actions(inputs, proc.tokenizer, do_sample=False, max_new_tokens=N_NEW)
        if world > 1:
            tick_gather(eng.d_out_tokens)
        last_action[0] = action

    e2e_ms, _ = timed(step_e2e, W, K)
    if world > 1:  # the collective's result is checked too: every replica's 7 action tokens, as planted
        want = torch.tensor(script[act_lo : act_lo + 7], dtype=torch.int32, device=dev)
        assert bool((gathered[:, :7] == want[None]).all()), f"all-gathered action tokens differ from the script: {gathered.tolist()}"
        parity["gathered_rows_checked"] = world
    e2e_value = world * K / (e2e_ms / 1e3)
    h2d = reqs[0][2].numel() + reqs[0][1].numel() * 8
    d2h = N_NEW * 4 + 4

    # ---- per-token latency distribution (untimed extra pass, events around every launch) ---------------------------
    import ctypes as C

    pv, ids = d_reqs[0]
    eng.generate(ids, pv, 2, eos_token_id=None)
    p = eng._decode_params(0)
    lib = _lib.load()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(129)]
    evs[0].record()
    for j in range(128):
        _lib.check(lib.emx_decode_step(C.byref(p), _lib.stream()))
        evs[j + 1].record()
    torch.cuda.synchronize()
    tok_ms = sorted(evs[j].elapsed_time(evs[j + 1]) for j in range(128))

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
        avg_launch_ms = decode_ms[0] / max(decode_launches[0], 1)
        achieved = decode_bytes_total[0] / max(decode_launches[0], 1) / (avg_launch_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "decode_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        line = {
            "metric": "actions/sec (7-DoF)", "value": value, "unit": "actions/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": c2_config(world),
            "clocks": clocks, "parity": parity,
            "e2e": {"value": e2e_value, "unit": "actions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / K},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "decode_step_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": "profiles/decode_traffic.json (ncu --set full capture of this kernel, dram read+write per launch)",
                         "peak_source": peak_src,
                         "bytes_per_launch": decode_bytes_total[0] / max(decode_launches[0], 1), "avg_launch_ms": avg_launch_ms,
                         "launches_timed": decode_launches[0], "share_of_step": decode_ms[0] / total_ms},
            "decode_ms_per_token": {"p50": tok_ms[64], "p10": tok_ms[12], "p90": tok_ms[115], "context": "296..424"},
            "action": [round(float(a), 6) for a in last_action[0]],
        }  # fmt: skip
        if extras_on:
            # short legs of the other BASELINE configs + the GPU comparator, so that the driver's one default run carries them too
            extras = {}
            for name, fn in (("c5", lambda: measure_c5(cfg, model, proc, script, dev, 1, 0, 4, 3, timed, lambda new: None)),
                             ("c3", lambda: measure_c3(cfg, model, proc, dev, 3, 3, timed)),
                             ("gpu_comparator", lambda: gpu_comparator_sample(cfg, sd, script, dev, steps=2, warmup=1)),
                             ("predict_action", lambda: predict_action_latency(cfg, model, proc, sd, script, dev))):  # fmt: skip
                try:
                    extras[name] = fn()
                except Exception as e:  # the headline line must survive a problem in an extra leg
                    extras[name] = {"failed": f"{type(e).__name__}: {e}"}
            if "value" in extras.get("gpu_comparator", {}):
                extras["gpu_comparator"]["ours_over_comparator_e2e"] = e2e_value / extras["gpu_comparator"]["value"]
                if "e2e" in extras.get("c5", {}):
                    extras["gpu_comparator"]["ours_c5_per_gpu_over_comparator_e2e"] = extras["c5"]["e2e"]["value"] / extras["gpu_comparator"]["value"]
            line["extras"] = extras
        if cpu_sd is not None:
            try:
                _, s = cpu_reference_sample(cfg, cpu_sd, 4)
                line["cpu_baseline"] = {
                    "value": s["actions_per_s"], "unit": "actions/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": (f"torch-eager oracle port on CPU, bf16 weights: vision+prefill(S=296) once ({s['t_prefill_s']:.2f}s) + 4 cached decode "
                               f"tokens ({s['t_token_s'] * 1e3:.0f} ms/token), extrapolated to the 512-token request"),
                }  # fmt: skip
            except Exception as e:  # the headline number must survive a CPU-side problem
                line["cpu_baseline"] = {"value": None, "unit": "actions/s", "cores": torch.get_num_threads(), "kind": "port",
                                        "sample": f"failed: {type(e).__name__}: {e}"}  # fmt: skip
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c5"], help="BASELINE.json configs[1] (headline, default), [2] or [4]")
    ap.add_argument("--no-extras", action="store_true", help="default config only: skip the short c5 / c3 / GPU-comparator legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-oracle", action="store_true", help="skip the live-oracle leg of the parity gate (the id check always runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
