"""Independent pin of the ViT restatement (oracle/vit.py).

timm==0.9.10 — the third-party package whose `VisionTransformer` the reference drives through
`get_intermediate_layers(n={depth-2})` (/root/reference/prismatic/extern/hf/modeling_prismatic.py:78-101) — is neither vendored in
the reference nor installed here, and the reference holds no fixture for it ("parity unpinned" for a2/a3). What IS
available is an independent implementation of the same two architectures: `transformers.Dinov2WithRegistersModel` (DINOv2 with 4
register tokens, LayerScale) and `transformers.SiglipVisionModel` (no class token, no LayerScale). Mapping the oracle's timm-named
parameters into them and comparing the tapped hidden state checks, against code we did not write: patch-embed conv layout, the
position-embedding / cls / register token order (timm `no_embed_class=True`: pos_embed covers the patches only), the pre-norm block
with fused-qkv attention split q|k|v, exact GELU, LayerScale placement, the tap index depth-2, the prefix strip and the absence of a
final norm. CPU, fp32, toy widths; runs in about a second."""

import pytest
import torch

from emmax_b200.configuration import ViTDims
from oracle.vit import OracleViT

transformers = pytest.importorskip("transformers")


def _rand_init(m: torch.nn.Module, seed: int) -> None:
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if "scale_factor" in name:
                p.copy_(torch.rand(p.shape, generator=g) * 0.9 + 0.1)
            elif name.endswith("norm1.weight") or name.endswith("norm2.weight") or name == "norm.weight":
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))


def _copy_block(dst: dict, prefix: str, sd: dict, i: int, D: int, names: dict) -> None:
    b = f"blocks.{i}."
    qkv_w, qkv_b = sd[b + "attn.qkv.weight"], sd[b + "attn.qkv.bias"]
    for j, n in enumerate(("q", "k", "v")):  # fused qkv rows are q | k | v major (timm Attention.forward reshape [B,N,3,heads,hd])
        dst[prefix + names[n] + ".weight"] = qkv_w[j * D : (j + 1) * D]
        dst[prefix + names[n] + ".bias"] = qkv_b[j * D : (j + 1) * D]
    for ours, theirs in (("attn.proj", names["o"]), ("norm1", names["ln1"]), ("norm2", names["ln2"]), ("mlp.fc1", names["fc1"]), ("mlp.fc2", names["fc2"])):
        dst[prefix + theirs + ".weight"] = sd[b + ours + ".weight"]
        dst[prefix + theirs + ".bias"] = sd[b + ours + ".bias"]


def test_dinov2_reg4_restatement_matches_hf_implementation():
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel

    v = ViTDims("toy_dino", 64, 5, 2, 256, num_prefix_tokens=5, layerscale=True, image_size=56)
    ours = OracleViT(v).eval()
    _rand_init(ours, 0)
    sd = ours.state_dict()
    cfg = Dinov2WithRegistersConfig(hidden_size=64, num_hidden_layers=v.depth, num_attention_heads=2, mlp_ratio=4, image_size=56, patch_size=14,
                                    num_register_tokens=4, layer_norm_eps=v.ln_eps, hidden_act="gelu", qkv_bias=True, use_swiglu_ffn=False,
                                    attn_implementation="eager")  # fmt: skip
    hf = Dinov2WithRegistersModel(cfg).eval()
    new = {k: t.clone() for k, t in hf.state_dict().items()}
    new["embeddings.patch_embeddings.projection.weight"] = sd["patch_embed.proj.weight"]
    new["embeddings.patch_embeddings.projection.bias"] = sd["patch_embed.proj.bias"]
    new["embeddings.cls_token"] = sd["cls_token"]
    new["embeddings.register_tokens"] = sd["reg_token"]
    # HF adds position_embeddings[:, 0] to the cls token; timm (no_embed_class) adds nothing to it
    new["embeddings.position_embeddings"] = torch.cat([torch.zeros(1, 1, 64), sd["pos_embed"]], dim=1)
    names = dict(q="attention.attention.query", k="attention.attention.key", v="attention.attention.value", o="attention.output.dense",
                 ln1="norm1", ln2="norm2", fc1="mlp.fc1", fc2="mlp.fc2")  # fmt: skip
    for i in range(v.depth):
        _copy_block(new, f"encoder.layer.{i}.", sd, i, 64, names)
        new[f"encoder.layer.{i}.layer_scale1.lambda1"] = sd[f"blocks.{i}.ls1.scale_factor"]
        new[f"encoder.layer.{i}.layer_scale2.lambda1"] = sd[f"blocks.{i}.ls2.scale_factor"]
    hf.load_state_dict(new, strict=True)
    img = torch.randn(2, 3, 56, 56, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = hf(pixel_values=img, output_hidden_states=True).hidden_states[v.depth - 1][:, 5:]  # output of block index depth-2
        got = ours(img)
        got_all = ours(img, run_all_blocks=True)
    assert got.shape == (2, 16, 64)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5), (got - want).abs().max()
    assert torch.equal(got, got_all), "the last block must not influence get_intermediate_layers(n={depth-2})"


def test_siglip_restatement_matches_hf_implementation():
    from transformers import SiglipVisionConfig, SiglipVisionModel

    v = ViTDims("toy_siglip", 72, 4, 2, 272, num_prefix_tokens=0, layerscale=False, image_size=56)
    ours = OracleViT(v).eval()
    _rand_init(ours, 2)
    sd = ours.state_dict()
    cfg = SiglipVisionConfig(hidden_size=72, intermediate_size=272, num_hidden_layers=v.depth, num_attention_heads=2, image_size=56, patch_size=14,
                             layer_norm_eps=v.ln_eps, hidden_act="gelu", attn_implementation="eager")  # fmt: skip
    hf = SiglipVisionModel(cfg).eval()
    new = {k: t.clone() for k, t in hf.state_dict().items()}
    new["vision_model.embeddings.patch_embedding.weight"] = sd["patch_embed.proj.weight"]
    new["vision_model.embeddings.patch_embedding.bias"] = sd["patch_embed.proj.bias"]
    new["vision_model.embeddings.position_embedding.weight"] = sd["pos_embed"][0]
    names = dict(q="self_attn.q_proj", k="self_attn.k_proj", v="self_attn.v_proj", o="self_attn.out_proj", ln1="layer_norm1", ln2="layer_norm2",
                 fc1="mlp.fc1", fc2="mlp.fc2")  # fmt: skip
    for i in range(v.depth):
        _copy_block(new, f"vision_model.encoder.layers.{i}.", sd, i, 72, names)
    hf.load_state_dict(new, strict=True)
    img = torch.randn(2, 3, 56, 56, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = hf(pixel_values=img, output_hidden_states=True).hidden_states[v.depth - 1]
        got = ours(img)
    assert got.shape == (2, 16, 72)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5), (got - want).abs().max()
