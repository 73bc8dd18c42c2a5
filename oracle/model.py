"""
Torch-eager restatement of the reference HF path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows /root/reference/prismatic/extern/hf/modeling_prismatic.py:
  :114-123  PrismaticVisionBackbone.forward   split 3|3 channels -> two ViTs -> cat on the feature dim
  :146-158  PrismaticProjector.forward        fc1 -> GELU -> fc2 -> GELU -> fc3 (fused-backbone branch)
  :362-415  multimodal forward                embed ids; cat [BOS | patches | rest]; LLM on inputs_embeds, positions 0..S-1
  :325-341  cached single-token step          bs == 1, attention_mask=None
  :506-537  predict_action                    append 29871, generate action_dim tokens, de-tokenise, un-normalise
and the greedy loop of transformers `GenerationMixin` (argmax of the last position's fp32 logits; stop on EOS).
The Llama stack is the container's `transformers.LlamaForCausalLM` itself (the third-party class the reference calls at
modeling_prismatic.py:248-250), not a re-implementation.
"""

from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .vit import OracleViT


class _VisionBackbone(nn.Module):
    def __init__(self, vision_dims) -> None:
        super().__init__()
        self.featurizer = OracleViT(vision_dims[0])
        self.fused_featurizer = OracleViT(vision_dims[1])

    def forward(self, pixel_values: torch.Tensor) -> torch.Tensor:
        img, img_fused = torch.split(pixel_values, [3, 3], dim=1)
        return torch.cat([self.featurizer(img), self.fused_featurizer(img_fused)], dim=2)


class _Projector(nn.Module):
    def __init__(self, vision_dim: int, llm_dim: int) -> None:
        super().__init__()
        self.fc1 = nn.Linear(vision_dim, 4 * vision_dim)
        self.fc2 = nn.Linear(4 * vision_dim, llm_dim)
        self.fc3 = nn.Linear(llm_dim, llm_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc3(F.gelu(self.fc2(F.gelu(self.fc1(x)))))


def _llama(text, attn_implementation: str):
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(
        vocab_size=text.vocab_size, hidden_size=text.hidden_size, intermediate_size=text.intermediate_size,
        num_hidden_layers=text.num_hidden_layers, num_attention_heads=text.num_attention_heads,
        num_key_value_heads=text.num_key_value_heads, rms_norm_eps=text.rms_norm_eps, rope_theta=text.rope_theta,
        max_position_embeddings=text.max_position_embeddings, pad_token_id=text.pad_token_id,
        bos_token_id=text.bos_token_id, eos_token_id=text.eos_token_id, tie_word_embeddings=False,
        attention_bias=False, mlp_bias=False, hidden_act="silu",
    )  # fmt: skip
    cfg._attn_implementation = attn_implementation
    return LlamaForCausalLM(cfg)


class OracleVLA(nn.Module):
    def __init__(self, config, attn_implementation: str = "sdpa") -> None:
        super().__init__()
        self.config = config
        self.vision_backbone = _VisionBackbone(config.vision_dims)
        self.projector = _Projector(config.vision_embed_dim, config.text_config.hidden_size)
        self.language_model = _llama(config.text_config, attn_implementation)
        self.norm_stats = config.norm_stats
        self.bins = np.linspace(-1, 1, config.n_action_bins)
        self.bin_centers = (self.bins[:-1] + self.bins[1:]) / 2.0
        self.vocab_size = config.text_config.vocab_size - config.pad_to_multiple_of
        self.eval()

    @classmethod
    def from_state_dict(cls, config, sd: Dict[str, torch.Tensor], device="cpu", dtype=torch.bfloat16,
                        attn_implementation: str = "sdpa") -> "OracleVLA":
        with torch.device("meta"):
            m = cls(config, attn_implementation)
        m = m.to_empty(device=device).to(dtype)
        missing, unexpected = m.load_state_dict({k: v.to(device=device, dtype=dtype) for k, v in sd.items()}, strict=False)
        # rotary inv_freq is a non-persistent buffer: rebuild after to_empty()
        assert not unexpected, unexpected
        assert all("rotary_emb" in k for k in missing), missing
        rot = m.language_model.model.rotary_emb
        inv_freq, rot.attention_scaling = rot.compute_default_rope_parameters(rot.config, device)
        rot.inv_freq = inv_freq
        rot.original_inv_freq = inv_freq.clone()
        return m.eval()

    # ---------------------------------------------------------------------------------------------------------------
    @torch.inference_mode()
    def multimodal_embeddings(self, input_ids: torch.Tensor, pixel_values: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        patches = self.projector(self.vision_backbone(pixel_values))
        emb = self.language_model.get_input_embeddings()(input_ids)
        return torch.cat([emb[:, :1], patches, emb[:, 1:]], dim=1), patches

    @torch.inference_mode()
    def prefill(self, input_ids: torch.Tensor, pixel_values: torch.Tensor):
        x, _ = self.multimodal_embeddings(input_ids, pixel_values)
        mask = torch.ones(x.shape[:2], dtype=torch.long, device=x.device)
        out = self.language_model(inputs_embeds=x, attention_mask=mask, use_cache=True)
        return out.logits.float(), out.past_key_values

    @torch.inference_mode()
    def step(self, token: torch.Tensor, past):
        assert token.shape == (1, 1), "Generation is only currently supported for batch size of 1!"
        out = self.language_model(input_ids=token, attention_mask=None, past_key_values=past, use_cache=True)
        return out.logits.float(), out.past_key_values

    @torch.inference_mode()
    def generate(self, input_ids: torch.Tensor, pixel_values: torch.Tensor, max_new_tokens: int,
                 eos_token_id: Optional[int] = 2, return_logits: bool = False, forced: Optional[List[int]] = None):
        """Greedy decode. `forced`: teacher-forcing ids fed back instead of the argmax (logits still returned)."""
        if input_ids.shape[0] != 1:
            raise ValueError("Generation with batch size > 1 is not currently supported!")
        logits, past = self.prefill(input_ids, pixel_values)
        new: List[int] = []
        trace: List[torch.Tensor] = []
        last = logits[:, -1]
        for t in range(max_new_tokens):
            if return_logits:
                trace.append(last[0].cpu())
            nxt = int(torch.argmax(last, dim=-1)[0])
            fed = nxt if forced is None else int(forced[t])
            new.append(fed if forced is not None else nxt)
            if forced is None and eos_token_id is not None and nxt == eos_token_id:
                break
            if t + 1 < max_new_tokens:
                logits, past = self.step(torch.tensor([[fed]], device=input_ids.device), past)
                last = logits[:, -1]
        ids = torch.cat([input_ids, torch.tensor([new], device=input_ids.device, dtype=input_ids.dtype)], dim=1)
        return (ids, torch.stack(trace)) if return_logits else ids

    @torch.inference_mode()
    def full_sequence_logits(self, input_ids: torch.Tensor, pixel_values: torch.Tensor) -> torch.Tensor:
        """One causal pass over prompt+continuation; row P-1+256+i predicts continuation token i."""
        logits, _ = self.prefill(input_ids, pixel_values)
        return logits

    def predict_action(self, input_ids: torch.Tensor, pixel_values: torch.Tensor, unnorm_key: Optional[str] = None) -> np.ndarray:
        if not torch.all(input_ids[:, -1] == 29871):
            input_ids = torch.cat((input_ids, torch.tensor([[29871]], device=input_ids.device)), dim=1)
        key = unnorm_key if unnorm_key is not None else next(iter(self.norm_stats.keys()))
        stats = self.norm_stats[key]["action"]
        n = len(stats["q01"])
        ids = self.generate(input_ids, pixel_values, max_new_tokens=n, eos_token_id=None)
        tok = ids[0, -n:].cpu().numpy()
        from .detok import decode_token_ids_to_actions, unnormalize_actions

        return unnormalize_actions(decode_token_ids_to_actions(tok, self.vocab_size), stats)
