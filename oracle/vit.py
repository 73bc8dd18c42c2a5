"""
Pure-torch restatement of timm==0.9.10 `VisionTransformer` as the reference drives it.

timm is not vendored in /root/reference and not installed here; the semantics below are the published ones of
`timm/models/vision_transformer.py` (0.9.10) for the two models the reference creates at
/root/reference/prismatic/extern/hf/modeling_prismatic.py:78-101 and calls through
`get_intermediate_layers(n={depth-2})` (:85-87, :99-101):

  x = Conv2d(3, D, k=14, s=14)(img).flatten(2).transpose(1, 2)          # PatchEmbed
  DINOv2-reg4 (no_embed_class=True): x = x + pos_embed; x = cat([cls_token, reg_token, x], 1)
  SigLIP (class_token=False):        x = x + pos_embed
  for blk: x = x + ls1(attn(norm1(x))); x = x + ls2(mlp(norm2(x)))       # LayerNorm eps 1e-6, exact GELU
  take the output of block index depth-2, strip prefix tokens, NO final norm.

Parameter names are timm's (with LayerScale `gamma` renamed `scale_factor`, modeling_prismatic.py:52-59) so the
reference's state-dict contract is exercised by `load_state_dict(strict=True)`.
"""

from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _LayerScale(nn.Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.scale_factor = nn.Parameter(torch.ones(dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x * self.scale_factor


class _Attention(nn.Module):
    def __init__(self, dim: int, heads: int) -> None:
        super().__init__()
        self.num_heads, self.head_dim = heads, dim // heads
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v)
        return self.proj(x.transpose(1, 2).reshape(B, N, C))


class _Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int) -> None:
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc2(F.gelu(self.fc1(x)))  # exact (erf) GELU: act_layer=None -> timm default nn.GELU


class _Block(nn.Module):
    def __init__(self, dim: int, heads: int, mlp_dim: int, layerscale: bool, eps: float) -> None:
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = _Attention(dim, heads)
        self.ls1 = _LayerScale(dim) if layerscale else nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = _Mlp(dim, mlp_dim)
        self.ls2 = _LayerScale(dim) if layerscale else nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x + self.ls1(self.attn(self.norm1(x)))
        return x + self.ls2(self.mlp(self.norm2(x)))


class _PatchEmbed(nn.Module):
    def __init__(self, dim: int, patch: int) -> None:
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch, bias=True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.proj(x).flatten(2).transpose(1, 2)


class OracleViT(nn.Module):
    def __init__(self, v) -> None:  # v: emmax_b200.configuration.ViTDims (duck-typed)
        super().__init__()
        self.v = v
        D = v.embed_dim
        self.patch_embed = _PatchEmbed(D, v.patch_size)
        self.pos_embed = nn.Parameter(torch.zeros(1, v.num_patches, D))
        if v.num_prefix_tokens > 0:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, D))
            self.reg_token = nn.Parameter(torch.zeros(1, v.num_prefix_tokens - 1, D))
        self.blocks = nn.ModuleList([_Block(D, v.num_heads, v.mlp_dim, v.layerscale, v.ln_eps) for _ in range(v.depth)])
        self.norm = nn.LayerNorm(D, eps=v.ln_eps)  # present in the checkpoint; unused on this path

    def forward(self, img: torch.Tensor, run_all_blocks: bool = False) -> torch.Tensor:
        v = self.v
        x = self.patch_embed(img) + self.pos_embed
        if v.num_prefix_tokens > 0:
            B = x.shape[0]
            x = torch.cat([self.cls_token.expand(B, -1, -1), self.reg_token.expand(B, -1, -1), x], dim=1)
        take = len(self.blocks) - 2
        out = None
        for i, blk in enumerate(self.blocks):
            # timm runs every block and keeps the tapped output; the last block cannot change it
            if i > take and not run_all_blocks:
                break
            x = blk(x)
            if i == take:
                out = x
        return out[:, v.num_prefix_tokens :]
