"""
Seeded synthetic checkpoints with the reference's state-dict names.

No Emma-X / OpenVLA weights, no Llama-2 tokenizer and no network exist in the build or GPU images,
so parity tests, `smoke()` and `bench.py` run on random-init weights of the exact reference architecture. The names
follow the HF export contract of /root/reference/vla-scripts/extern/convert_openvla_weights_to_hf.py:84-116:

    vision_backbone.featurizer.*        timm DINOv2 names, LayerScale as `ls{1,2}.scale_factor`
    vision_backbone.fused_featurizer.*  timm SigLIP names
    projector.fc{1,2,3}.{weight,bias}
    language_model.model.* / language_model.lm_head.weight

"Scripted" heads: greedy token ids of a random-init LM are decided by near-ties that flip under any change of
reduction order, so bit-exact id parity would be meaningless noise. `script=[...]` therefore plants a known answer:
`lm_head[script[i+1]] = gain * embed[script[i]]`, which makes the greedy continuation of a prompt ending in
`script[0]`'s predecessor follow `script` with top-2 margins far above bf16 noise, while every layer still
contributes O(1) to the hidden state (so logits remain a sensitive numerical probe).
"""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .configuration import OpenVLAConfig, ViTDims


def _randn(shape, std, gen, device, dtype, mean=0.0):
    t = torch.empty(shape, dtype=torch.float32, device=device)
    t.normal_(mean, std, generator=gen)
    return t.to(dtype)


def _vit_state(prefix: str, v: ViTDims, gen, device, dtype, sd: Dict[str, torch.Tensor]) -> None:
    D, P = v.embed_dim, v.patch_size
    sd[f"{prefix}patch_embed.proj.weight"] = _randn((D, 3, P, P), 0.02, gen, device, dtype)
    sd[f"{prefix}patch_embed.proj.bias"] = _randn((D,), 0.02, gen, device, dtype)
    sd[f"{prefix}pos_embed"] = _randn((1, v.num_patches, D), 0.02, gen, device, dtype)
    if v.num_prefix_tokens > 0:
        sd[f"{prefix}cls_token"] = _randn((1, 1, D), 0.02, gen, device, dtype)
        sd[f"{prefix}reg_token"] = _randn((1, v.num_prefix_tokens - 1, D), 0.02, gen, device, dtype)
    for i in range(v.depth):
        b = f"{prefix}blocks.{i}."
        sd[b + "norm1.weight"] = _randn((D,), 0.02, gen, device, dtype, mean=1.0)
        sd[b + "norm1.bias"] = _randn((D,), 0.02, gen, device, dtype)
        sd[b + "attn.qkv.weight"] = _randn((3 * D, D), 0.02, gen, device, dtype)
        sd[b + "attn.qkv.bias"] = _randn((3 * D,), 0.02, gen, device, dtype)
        sd[b + "attn.proj.weight"] = _randn((D, D), 0.02, gen, device, dtype)
        sd[b + "attn.proj.bias"] = _randn((D,), 0.02, gen, device, dtype)
        sd[b + "norm2.weight"] = _randn((D,), 0.02, gen, device, dtype, mean=1.0)
        sd[b + "norm2.bias"] = _randn((D,), 0.02, gen, device, dtype)
        sd[b + "mlp.fc1.weight"] = _randn((v.mlp_dim, D), 0.02, gen, device, dtype)
        sd[b + "mlp.fc1.bias"] = _randn((v.mlp_dim,), 0.02, gen, device, dtype)
        sd[b + "mlp.fc2.weight"] = _randn((D, v.mlp_dim), 0.02, gen, device, dtype)
        sd[b + "mlp.fc2.bias"] = _randn((D,), 0.02, gen, device, dtype)
        if v.layerscale:
            # not timm's 1e-5 init (blocks would be near-identity): U(0.05, 1) keeps every block numerically relevant
            for ls in ("ls1", "ls2"):
                t = torch.empty((D,), dtype=torch.float32, device=device).uniform_(0.05, 1.0, generator=gen)
                sd[b + f"{ls}.scale_factor"] = t.to(dtype)
    # final norm exists in the timm state dict but is unused by `get_intermediate_layers` (no norm applied)
    sd[f"{prefix}norm.weight"] = _randn((D,), 0.02, gen, device, dtype, mean=1.0)
    sd[f"{prefix}norm.bias"] = _randn((D,), 0.02, gen, device, dtype)


def make_state_dict(
    config: OpenVLAConfig,
    seed: int = 0,
    device: str | torch.device = "cpu",
    dtype: torch.dtype = torch.bfloat16,
    script: Optional[Sequence[int]] = None,
    script_prev: Optional[int] = None,
    extra_chains: Optional[Sequence[Tuple[int, Sequence[int]]]] = None,
    head_gain: float = 16.0,
    resid_scale: float = 1.0,
) -> Dict[str, torch.Tensor]:
    """Build a full state dict (reference names). `script`/`script_prev`: see module docstring."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    _vit_state("vision_backbone.featurizer.", config.vision_dims[0], gen, device, dtype, sd)
    _vit_state("vision_backbone.fused_featurizer.", config.vision_dims[1], gen, device, dtype, sd)

    t = config.text_config
    vd, H = config.vision_embed_dim, t.hidden_size
    sd["projector.fc1.weight"] = _randn((4 * vd, vd), 0.02, gen, device, dtype)
    sd["projector.fc1.bias"] = _randn((4 * vd,), 0.02, gen, device, dtype)
    sd["projector.fc2.weight"] = _randn((H, 4 * vd), 0.02, gen, device, dtype)
    sd["projector.fc2.bias"] = _randn((H,), 0.02, gen, device, dtype)
    sd["projector.fc3.weight"] = _randn((H, H), 0.05, gen, device, dtype)
    sd["projector.fc3.bias"] = _randn((H,), 0.02, gen, device, dtype)

    # Token embeddings at unit scale. Projections are scaled so that every Linear output is O(1) at any width
    # (q/k/v/gate/up: std 1/sqrt(H)), and the residual branches (o_proj / down_proj) so that the SUM of all 2L
    # branch outputs has rms ~1, i.e. comparable to the embedding: the final hidden state keeps cos ~0.7 with the
    # current token's embedding (what the scripted head keys on) while every layer still matters numerically.
    p = "language_model.model."
    emb32 = torch.empty((t.vocab_size, H), dtype=torch.float32, device=device).normal_(0.0, 1.0, generator=gen)
    sd[p + "embed_tokens.weight"] = emb32.to(dtype)
    I, L = t.intermediate_size, t.num_hidden_layers
    a = resid_scale * 1.47 / L**0.5
    for i in range(L):
        b = f"{p}layers.{i}."
        sd[b + "input_layernorm.weight"] = _randn((H,), 0.02, gen, device, dtype, mean=1.0)
        for n in ("q_proj", "k_proj", "v_proj"):
            sd[b + f"self_attn.{n}.weight"] = _randn((H, H), H**-0.5, gen, device, dtype)
        sd[b + "self_attn.o_proj.weight"] = _randn((H, H), a * H**-0.5, gen, device, dtype)
        sd[b + "post_attention_layernorm.weight"] = _randn((H,), 0.02, gen, device, dtype, mean=1.0)
        sd[b + "mlp.gate_proj.weight"] = _randn((I, H), H**-0.5, gen, device, dtype)
        sd[b + "mlp.up_proj.weight"] = _randn((I, H), H**-0.5, gen, device, dtype)
        sd[b + "mlp.down_proj.weight"] = _randn((H, I), a * I**-0.5, gen, device, dtype)
    sd[p + "norm.weight"] = _randn((H,), 0.02, gen, device, dtype, mean=1.0)

    # un-scripted rows give ~N(0,1) logits; a scripted row peaks at ~head_gain * cos(hidden, embed[prev])
    head = torch.empty((t.vocab_size, H), dtype=torch.float32, device=device).normal_(0.0, H**-0.5, generator=gen)
    chains: List[Tuple[int, List[int]]] = []
    if script is not None:
        assert script_prev is not None
        chains.append((int(script_prev), list(script)))
    for prev0, ids in extra_chains or []:
        chains.append((int(prev0), list(ids)))
    if script is not None and not extra_chains:
        chains.append(predict_action_chain(list(script)))
    all_next = [i for _, ids in chains for i in ids]
    assert len(set(all_next)) == len(all_next), "scripted ids must be unique (successor is a function of the current id)"
    for prev0, ids in chains:
        assert prev0 not in all_next
        prev = [prev0] + ids[:-1]
        idx_next = torch.tensor(ids, device=device)
        idx_prev = torch.tensor(prev, device=device)
        head[idx_next] = (head_gain / H) * sd[p + "embed_tokens.weight"][idx_prev].float()
    sd["language_model.lm_head.weight"] = head.to(dtype)
    return sd


# === scripted continuation used by bench / smoke / parity ============================================================
def predict_action_chain(main_script: Sequence[int], action_dim: int = 7) -> Tuple[int, List[int]]:
    """Second planted chain for `predict_action` (which appends id 29871 and decodes `action_dim` tokens,
    /root/reference/prismatic/extern/hf/modeling_prismatic.py:513-519): 29871 -> 7 action ids unused by the main script."""
    used = set(main_script)
    free = [i for i in range(31999, 31744, -1) if i not in used]
    return 29871, free[3 : 3 + 5 * action_dim : 5]


def default_script(tokenizer, n_new: int, seed: int = 0, n_policies: int = 2) -> List[int]:
    """A unique-id script of `n_new` tokens shaped like the grounded-CoT output grammar
    (/root/reference/prismatic/vla/datasets/datasets.py:483-581): filler "reasoning", then
    `MOVEMENT:\\n<7 act>\\nPOLICIES:\\n<7 act>;<7 act>\\n`, then EOS as the last token."""
    import numpy as np

    rng = np.random.default_rng(seed)
    act = rng.permutation(np.arange(tokenizer.action_id_lo + 1, tokenizer.action_id_hi + 1))  # 31745..31999
    act = [int(a) for a in act[: 7 * (n_policies + 1)]]
    tail: List[int] = [tokenizer.key_id("MOVEMENT:"), tokenizer.newline_id(0)]
    tail += act[:7] + [tokenizer.newline_id(1)]
    tail += [tokenizer.key_id("POLICIES:"), tokenizer.newline_id(2)]
    for p in range(n_policies):
        if p > 0:
            tail.append(tokenizer.semicolon_id(p - 1))
        tail += act[7 * (p + 1) : 7 * (p + 2)]
    tail += [tokenizer.newline_id(3), tokenizer.eos_token_id]
    n_fill = n_new - len(tail)
    assert n_fill >= 0, f"n_new={n_new} too short for the scripted tail ({len(tail)})"
    filler = rng.permutation(np.arange(tokenizer.filler_lo, tokenizer.filler_hi))[:n_fill]
    return [int(x) for x in filler] + tail
