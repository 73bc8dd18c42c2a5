"""Pins the oracle's WIRING against the reference's own HF model class (TEST INFRASTRUCTURE).

/root/reference/prismatic/extern/hf/{configuration,modeling}_prismatic.py are loaded by path and `OpenVLAForActionPrediction` is
instantiated on the CPU (fp32) with the toy widths of `emmax_b200.tiny_config()` and the seeded state dict of `emmax_b200.synthetic`.
What runs is the reference's code: state-dict naming (load_state_dict reports 0 missing / 0 unexpected keys), LayerScale patching
(:52-59), vision-backbone channel split and feature concat (:114-123), projector (:146-158), multimodal sequence assembly and the LLM
call (:362-415), the cached single-token branch (:325-341), and `predict_action`'s append-29871 / de-tokenise / un-normalise code (:506-537).
`GenerationMixin.generate` itself is NOT taken from transformers 5.x: its loop pre-creates the KV cache, which sends the reference's
`prepare_inputs_for_generation` (:466-468, written for 4.40.1) down the cached branch on the very first step. The greedy loop (argmax of
the last position, feed back one token with `past_key_values`) is therefore driven here over the reference's own `forward`, and
`predict_action` runs with `self.generate` bound to that loop. The state dict carries a planted continuation (emmax_b200.synthetic: a
low-rank script in lm_head) so that the 7 generated tokens ARE action tokens and the de-tokeniser sees meaningful ids.
Stand-ins, installed BEFORE the import:
  * `timm` (absent; 0.9.10 pinned by the reference): `create_model` returns oracle.vit.OracleViT wrapped with exactly the attributes
    the reference touches (`blocks`, `embed_dim`, `get_intermediate_layers(x, n={depth-2})` returning a tuple), and
    `timm.models.vision_transformer.LayerScale` with timm's `gamma` parameter, so the reference's own patch code renames it.
    The ViT INTERNALS therefore remain the restatement (cross-checked against transformers' DINOv2-reg / SigLIP in
    tests/test_oracle_vit_crosscheck.py);
  * transformers 5.x compatibility (the reference pins 4.40.1): `tie_weights` is called with keyword arguments (Llama-2 ties nothing:
    replaced by a no-op); `generate` as described above.
Outputs frozen in tests/golden/hf_model_golden.npz; tests/test_oracle_golden.py requires OracleVLA to reproduce them.
Usage (container with /root/reference):  python oracle/gen_golden_hf_model.py
"""
import hashlib
import importlib.util
import os
import sys
import types
from importlib.machinery import ModuleSpec

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_DIR = "/root/reference/prismatic/extern/hf"
OUT = os.path.join(ROOT, "tests", "golden", "hf_model_golden.npz")
N_IDS, N_NEW, SEED = 21, 12, 0


def case_inputs():
    ids = torch.tensor([[1] + np.random.default_rng(SEED).integers(3, 300, N_IDS - 1).tolist()])
    pv = torch.randn(1, 6, 224, 224, generator=torch.Generator().manual_seed(SEED))
    return ids, pv


def load_reference(cfg):
    import transformers.modeling_outputs  # noqa: F401  (resolve transformers' lazy imports before the timm stand-in exists)
    from transformers import AutoModelForCausalLM, PreTrainedModel  # noqa: F401

    from oracle.vit import OracleViT

    class LayerScale(torch.nn.Module):  # timm 0.9.10 vision_transformer.LayerScale
        def __init__(self, dim, init_values=1e-5, inplace=False):
            super().__init__()
            self.inplace = inplace
            self.gamma = torch.nn.Parameter(init_values * torch.ones(dim))

        def forward(self, x):
            return x.mul_(self.gamma) if self.inplace else x * self.gamma

    class StubViT(OracleViT):
        def __init__(self, v):
            super().__init__(v)
            self.embed_dim = v.embed_dim
            if v.layerscale:
                for b in self.blocks:
                    b.ls1, b.ls2 = LayerScale(v.embed_dim), LayerScale(v.embed_dim)

        def get_intermediate_layers(self, x, n):
            (k,) = tuple(n)
            assert k == len(self.blocks) - 2
            return (OracleViT.forward(self, x, run_all_blocks=True),)

    dims = {"vit_large_patch14_reg4_dinov2.lvd142m": cfg.vision_dims[0], "vit_so400m_patch14_siglip_224": cfg.vision_dims[1]}

    def create_model(name, pretrained=False, num_classes=0, img_size=224, act_layer=None):
        assert not pretrained and num_classes == 0 and img_size == 224 and act_layer is None
        return StubViT(dims[name])

    timm = types.ModuleType("timm")
    timm.__spec__, timm.__version__, timm.create_model = ModuleSpec("timm", None), "0.9.10", create_model
    tm, tv = types.ModuleType("timm.models"), types.ModuleType("timm.models.vision_transformer")
    tm.__spec__, tv.__spec__ = ModuleSpec("timm.models", None), ModuleSpec("timm.models.vision_transformer", None)
    tv.LayerScale, timm.models, tm.vision_transformer = LayerScale, tm, tv
    sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.vision_transformer": tv})
    pkg = types.ModuleType("ref_hf")
    pkg.__path__, pkg.__spec__ = [REF_DIR], ModuleSpec("ref_hf", None, is_package=True)
    sys.modules["ref_hf"] = pkg
    for name in ("configuration_prismatic", "modeling_prismatic"):
        spec = importlib.util.spec_from_file_location("ref_hf." + name, os.path.join(REF_DIR, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["ref_hf." + name] = m
        spec.loader.exec_module(m)
    return sys.modules["ref_hf.configuration_prismatic"], sys.modules["ref_hf.modeling_prismatic"]


def ref_greedy(model, input_ids, pixel_values, max_new_tokens, **_):
    """GenerationMixin's greedy search as of transformers 4.40.1, over the reference's own forward (multimodal prefill, then cached steps)."""
    with torch.no_grad():
        out = model(input_ids=input_ids, attention_mask=torch.ones_like(input_ids), pixel_values=pixel_values, use_cache=True)
        ids = input_ids
        for _t in range(max_new_tokens):
            nxt = out.logits[:, -1].argmax(-1, keepdim=True)
            ids = torch.cat([ids, nxt], dim=1)
            if _t + 1 < max_new_tokens:
                out = model(input_ids=nxt, past_key_values=out.past_key_values, use_cache=True)
    return ids


if __name__ == "__main__":
    from emmax_b200 import SyntheticLlamaTokenizer, tiny_config
    from emmax_b200.synthetic import make_state_dict

    cfg = tiny_config()
    rc, rm = load_reference(cfg)
    t = cfg.text_config
    text = dict(vocab_size=t.vocab_size, hidden_size=t.hidden_size, intermediate_size=t.intermediate_size, num_hidden_layers=t.num_hidden_layers,
                num_attention_heads=t.num_attention_heads, num_key_value_heads=t.num_key_value_heads, rms_norm_eps=t.rms_norm_eps,
                rope_theta=t.rope_theta, max_position_embeddings=t.max_position_embeddings, pad_token_id=t.pad_token_id, bos_token_id=1,
                eos_token_id=2, tie_word_embeddings=False)  # fmt: skip
    hf_cfg = rc.OpenVLAConfig(vision_backbone_id="dinosiglip-vit-so-224px", llm_backbone_id="llama2-7b-pure", arch_specifier="no-align+fused-gelu-mlp",
                              use_fused_vision_backbone=True, image_resize_strategy="resize-naive", text_config=text, norm_stats=cfg.norm_stats,
                              n_action_bins=256)  # fmt: skip
    hf_cfg._attn_implementation = "eager"
    rm.PrismaticForConditionalGeneration.tie_weights = lambda self, *a, **k: None
    model = rm.OpenVLAForActionPrediction(hf_cfg).eval()
    tok = SyntheticLlamaTokenizer()
    action_script = [int(a) for a in np.random.default_rng(5).permutation(np.arange(tok.action_id_lo + 1, tok.action_id_hi + 1))[:7]] + [2]
    sd = make_state_dict(cfg, seed=SEED, device="cpu", script=action_script, script_prev=29871)
    missing, unexpected = model.load_state_dict({k: v.float() for k, v in sd.items()}, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    ids, pv = case_inputs()
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=torch.ones_like(ids), pixel_values=pv, use_cache=True)
        nxt = out.logits[:, -1].argmax(-1, keepdim=True)
        step = model(input_ids=nxt, past_key_values=out.past_key_values, use_cache=True)
        feats = model.vision_backbone(pv)
        proj = model.projector(feats)
    gen = ref_greedy(model, ids, pv, N_NEW)
    model.generate = lambda input_ids, max_new_tokens, **kw: ref_greedy(model, input_ids, kw["pixel_values"], max_new_tokens)
    action = model.predict_action(input_ids=ids, pixel_values=pv, unnorm_key=None)  # (the reference appends 29871 itself, :512-515)
    ids29871 = torch.cat([ids, torch.tensor([[29871]])], dim=1)
    action_ids = ref_greedy(model, ids29871, pv, 7)[0, -7:]
    assert action_ids.tolist() == action_script[:7], "the planted script must surface through the reference's forward"
    logits = out.logits.float().contiguous()
    np.savez_compressed(
        OUT, input_ids=ids.numpy(), pixel_seed=np.array(SEED), state_dict_keys=np.array(sorted(sd.keys())), action_script=np.array(action_script),
        prefill_argmax=logits.argmax(-1).numpy(), prefill_last_logits=logits[0, -1].numpy(), prefill_sha256=np.array(hashlib.sha256(logits.numpy().tobytes()).hexdigest()),
        step_token=nxt.numpy(), step_logits=step.logits[0, -1].float().numpy(), features_probe=feats[0, ::37, ::53].numpy(), projected_probe=proj[0, ::37, ::29].numpy(),
        predict_action=np.asarray(action, dtype=np.float64), action_ids=action_ids.numpy(), generated_ids=gen.numpy(),
    )  # fmt: skip
    print("wrote", OUT, os.path.getsize(OUT), "bytes; logits", tuple(logits.shape), "action", action)
