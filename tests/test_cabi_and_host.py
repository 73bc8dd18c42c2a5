"""CPU-only checks: the C-ABI library loads and exports every symbol include/emmax.h declares, the ctypes mirror of
its structs matches the C layout, and the product path fails loudly without CUDA (no fallback)."""

import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from emmax_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return _lib


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "emmax.h")).read()
    declared = set(re.findall(r"\b(emx_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    handle = lib.load()
    missing = [s for s in sorted(declared) if not hasattr(handle, s)]
    assert not missing, missing
    assert set(lib.EXPORTS) == declared, set(lib.EXPORTS) ^ declared
    assert handle.emx_arch() == b"sm_100a" and handle.emx_abi_version() == 5 and handle.emx_decode_grid() == 148


def test_ctypes_structs_match_c_layout(lib):
    prog = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "emmax.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(emx_decode_state), offsetof(emx_decode_state, head_ticket),
             sizeof(emx_decode_params), offsetof(emx_decode_params, embed), offsetof(emx_decode_params, page_size),
             offsetof(emx_decode_params, out_tokens), offsetof(emx_decode_params, state), offsetof(emx_decode_params, debug_flags));
      printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(emx_decode_batch_state), offsetof(emx_decode_batch_state, limit),
             offsetof(emx_decode_batch_state, epoch), sizeof(emx_decode_batch_params), offsetof(emx_decode_batch_params, embed),
             offsetof(emx_decode_batch_params, page_size), offsetof(emx_decode_batch_params, x), offsetof(emx_decode_batch_params, out_tokens),
             offsetof(emx_decode_batch_params, state), offsetof(emx_decode_batch_params, l2_lookahead_stages));
      return 0;
    }"""
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    S, P = lib.DecodeState, lib.DecodeParams
    want = [C.sizeof(S), S.head_ticket.offset, C.sizeof(P), P.embed.offset, P.page_size.offset, P.out_tokens.offset, P.state.offset,
            P.debug_flags.offset]  # fmt: skip
    assert got[:8] == want, (got, want)
    BS, BP = lib.DecodeBatchState, lib.DecodeBatchParams
    want_b = [C.sizeof(BS), BS.limit.offset, BS.epoch.offset, C.sizeof(BP), BP.embed.offset, BP.page_size.offset, BP.x.offset, BP.out_tokens.offset,
              BP.state.offset, BP.l2_lookahead_stages.offset]  # fmt: skip
    assert got[8:] == want_b, (got[8:], want_b)


def test_argument_validation_without_gpu(lib):
    handle = lib.load()
    # error paths that return before any CUDA call
    assert handle.emx_gemm_bf16(None, 8, None, 8, None, 8, 0, 8, 8, None, None, None, 0, 0, 0, None) != 0
    assert b"empty problem" in handle.emx_last_error()
    assert handle.emx_layernorm(None, None, None, None, 4, 7, 1e-6, None) != 0
    assert b"multiple of 8" in handle.emx_last_error()
    assert handle.emx_attn_fwd(None, None, 0, 1, 1, 64, 0, 1.0, None) != 0


def test_no_cpu_fallback():
    from emmax_b200 import OpenVLAForActionPrediction, tiny_config
    from emmax_b200._lib import EmxError
    from emmax_b200.synthetic import make_state_dict

    cfg = tiny_config()
    model = OpenVLAForActionPrediction(cfg, make_state_dict(cfg, seed=0))
    with pytest.raises(EmxError):
        model.to("cpu")
    with pytest.raises(EmxError):
        model.generate(torch.ones((1, 4), dtype=torch.long), pixel_values=torch.zeros(1, 6, 224, 224), max_new_tokens=2)
    with pytest.raises(ValueError):
        model._check_unnorm_key(model.norm_stats, "missing")
    assert model.get_action_dim() == 7 and model.vocab_size == 32000 and model.bin_centers.shape == (255,)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "emmax_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_config_roundtrip(tmp_path):
    from emmax_b200 import OpenVLAConfig, emma_x_config

    cfg = emma_x_config()
    cfg.save_pretrained(str(tmp_path))
    back = OpenVLAConfig.from_pretrained(str(tmp_path))
    assert back.to_dict() == cfg.to_dict()
    assert back.text_config.vocab_size == 32064 and back.vision_embed_dim == 2176 and back.num_patches == 256
    with pytest.raises(ValueError):
        OpenVLAConfig(vision_backbone_id="clip-vit-l")


def test_roofline_denominator_matches_the_survey():
    """bench.py's algorithmic bytes per decode step are SURVEY.md §8(d) / BASELINE.md §3: 13,214,679,040 B of weights (32 x 202,375,168
    layer parameters + 131,334,144 lm_head, bf16) + 524,288 B per cached token (KV read) + 524,288 B (KV write)."""
    import bench
    from emmax_b200 import emma_x_config

    cfg = emma_x_config()
    assert bench.decode_bytes(cfg, 0) == 13_214_679_040 + 524_288
    assert bench.decode_bytes(cfg, 300) - bench.decode_bytes(cfg, 0) == 300 * 524_288
    assert round(bench.decode_bytes(cfg, 300) / 1e9, 2) == 13.37 and round(bench.decode_bytes(cfg, 812) / 1e9, 2) == 13.64


def test_decode_row_partition_covers_every_row_once():
    """The decode kernel's static schedule: each of the 148 CTAs owns a contiguous run of rows of every weight phase. Checked on the host
    through the same formula the kernel compiles (emx_decode_phase_rows): the runs tile [0, N) without gaps or overlap, never split an
    LL unit (row pair; gate/up quad), are balanced to one granule, and the residual-producing phases fit the kernel's staging buffer."""
    import ctypes as C

    from emmax_b200 import _lib, emma_x_config, tiny_config

    lib = _lib.load()
    grid = lib.emx_decode_grid()
    for cfg in (emma_x_config(), tiny_config()):
        t = cfg.text_config
        H, I, V = t.hidden_size, t.intermediate_size, t.vocab_size
        for n_rows, g in ((H, 2), (2 * I, 4), (V, 2)):
            prev_end, sizes = 0, []
            for cta in range(grid):
                b, e = C.c_int(), C.c_int()
                assert lib.emx_decode_phase_rows(n_rows, g, cta, grid, C.byref(b), C.byref(e)) == 0
                assert b.value == prev_end and e.value >= b.value and b.value % g == 0 and e.value % g == 0
                prev_end = e.value
                sizes.append(e.value - b.value)
            assert prev_end == n_rows, "all rows covered"
            assert max(sizes) - min(sizes) <= g, "balanced to one granule"
        assert max(1, H // 2 // grid + 2) <= 192, "residual pairs of one CTA fit DEC_MAX_RESID"
    b, e = C.c_int(), C.c_int()
    assert lib.emx_decode_phase_rows(10, 3, 0, grid, C.byref(b), C.byref(e)) != 0 and b"emx_decode_phase_rows" in lib.emx_last_error()


def test_sparse_text_config_takes_the_transformers_defaults(tmp_path):
    """A checkpoint's config.json may hold only the keys the HF exporter patches (vocab_size, pad_token_id; convert_openvla_weights_to_hf.py:175-177).
    The reference then runs `LlamaConfig(**text_config)` (configuration_prismatic.py:119-123), i.e. every other field is the transformers
    default - rms_norm_eps 1e-6, max_position_embeddings 2048 - and the loader here must resolve to the same numbers."""
    import json

    from transformers import LlamaConfig

    from emmax_b200 import OpenVLAConfig, emma_x_config

    d = emma_x_config().to_dict()
    d["text_config"] = {"vocab_size": 32064, "pad_token_id": 32000}
    with open(tmp_path / "config.json", "w") as f:
        json.dump(d, f)
    got = OpenVLAConfig.from_pretrained(str(tmp_path)).text_config
    want = LlamaConfig(vocab_size=32064, pad_token_id=32000)
    for k in ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "num_key_value_heads",
              "rms_norm_eps", "max_position_embeddings", "bos_token_id", "eos_token_id", "pad_token_id"):  # fmt: skip
        assert getattr(got, k) == getattr(want, k), (k, getattr(got, k), getattr(want, k))
    assert got.rope_theta == 10000.0
    # the synthetic / native configuration keeps Llama-2's own 1e-5, and a full text_config round-trips unchanged
    assert emma_x_config().text_config.rms_norm_eps == 1e-5
    emma_x_config().save_pretrained(str(tmp_path / "full"))
    assert OpenVLAConfig.from_pretrained(str(tmp_path / "full")).text_config.rms_norm_eps == 1e-5


def test_checkpoint_dir_without_tokenizer_warns(tmp_path):
    from emmax_b200.tokenization import SyntheticLlamaTokenizer, load_tokenizer

    with pytest.warns(UserWarning, match="SYNTHETIC"):
        tok = load_tokenizer(str(tmp_path))
    assert isinstance(tok, SyntheticLlamaTokenizer)
    assert isinstance(load_tokenizer(None), SyntheticLlamaTokenizer)


def test_native_run_dir_loader_name_map(tmp_path):
    """`load_vla(<run>/checkpoints/x.pt)` (prismatic/models/load.py:122-228): the native component dicts are renamed exactly as
    convert_openvla_weights_to_hf.py:74-116 does (projector.{0,2,4} -> fc{1,2,3}, llm. -> language_model., dino_/siglip_featurizer. ->
    vision_backbone.(fused_)featurizer., .gamma -> .scale_factor), statistics and proprio stats are attached, and the reference's own
    path validation applies."""
    import json

    from emmax_b200 import load_vla, tiny_config
    from emmax_b200.load import BASE_VLM_REGISTRY, remap_native_state_dict, to_native_state_dict
    from emmax_b200.synthetic import make_state_dict

    cfg = tiny_config()
    sd = make_state_dict(cfg, seed=5)
    native = to_native_state_dict(sd)
    assert set(native) == {"vision_backbone", "projector", "llm_backbone"}
    assert "projector.0.weight" in native["projector"] and "projector.4.bias" in native["projector"]
    assert "llm.model.embed_tokens.weight" in native["llm_backbone"] and "llm.lm_head.weight" in native["llm_backbone"]
    assert "dino_featurizer.blocks.0.ls1.gamma" in native["vision_backbone"] and "siglip_featurizer.pos_embed" in native["vision_backbone"]
    assert not any("scale_factor" in k for k in native["vision_backbone"])
    back = remap_native_state_dict(native)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)

    run = tmp_path / "run-1"
    (run / "checkpoints").mkdir(parents=True)
    torch.save({"model": native}, run / "checkpoints" / "latest-checkpoint.pt")
    stats = {"bridge_orig": {"action": {"q01": [-1.0] * 7, "q99": [1.0] * 7, "mask": [True] * 6 + [False]}}}
    with pytest.raises(AssertionError):  # config.json missing
        load_vla(run / "checkpoints" / "latest-checkpoint.pt")
    with open(run / "config.json", "w") as f:
        json.dump({"vla": {"base_vlm": "prism-dinosiglip-224px+7b", "vla_id": "emma-x"}}, f)
    with open(run / "dataset_statistics.json", "w") as f:
        json.dump(stats, f)
    with pytest.warns(UserWarning, match="SYNTHETIC"):  # no tokenizer files in the run dir
        vla = load_vla(run / "checkpoints" / "latest-checkpoint.pt", proprio_norm_stats={"Q1": [0.0] * 7, "Q99": [1.0] * 7}, config=cfg)
    assert set(vla._sd) == set(sd) and all(torch.equal(vla._sd[k], sd[k]) for k in sd)
    assert vla.norm_stats == stats and vla.get_action_dim() == 7 and vla.get_proprio_stats()["Q99"] == [1.0] * 7
    # registry lookup (no override): the Emma-X base VLM resolves to the accelerated backbone pair at full Llama-2-7B size
    with pytest.warns(UserWarning, match="SYNTHETIC"):
        full = load_vla(str(run / "checkpoints" / "latest-checkpoint.pt"))
    assert full.config.vision_backbone_id == "dinosiglip-vit-so-224px" and full.config.text_config.hidden_size == 4096
    assert full.config.text_config.rms_norm_eps == 1e-5 and BASE_VLM_REGISTRY["prism-dinosiglip-224px+7b"]["image_resize_strategy"] == "resize-naive"
    # reference validation: wrong location / hub ids
    bad = tmp_path / "x.pt"
    torch.save({"model": native}, bad)
    with pytest.raises(AssertionError, match="Invalid checkpoint"):
        load_vla(bad)
    with pytest.raises(ValueError, match="Couldn't find valid HF Hub Path"):
        load_vla("emma-x-7b")
    # the HF-style directory loader accepts the same native file too
    from emmax_b200.modeling import _load_checkpoint_tensors

    assert set(_load_checkpoint_tensors(str(run / "checkpoints"))) == set(sd)
    # experiments/robot/robot_utils.py:42 asks for float16; the engine stays bf16 and says so
    with pytest.warns(UserWarning, match="bf16"):
        vla._check_dtype(torch.float16)


def test_admission_order_of_the_continuous_batching_stream():
    """Engine.serve admits requests in `admission_order`: FIFO, or largest token limit first with ties in arrival order. A host-side model
    of the 8-slot stream (one launch = one token for every live slot) shows what the second buys on BASELINE.json configs[4]'s mix of
    128- and 512-token requests: the stream ends within one SHORT request of the ideal sum(limits) / 8 launches."""
    import pytest

    from emmax_b200.engine import admission_order

    limits = [512, 128] * 12
    assert admission_order(limits) == list(range(24))
    lpt = admission_order(limits, "longest_first")
    assert lpt == list(range(0, 24, 2)) + list(range(1, 24, 2))
    assert admission_order([], "longest_first") == []
    with pytest.raises(ValueError):
        admission_order(limits, "random")

    def launches(order):  # tokens after the first come from decode launches: limit - 1 per request
        queue, slots, n = list(order), [0] * 8, 0
        while queue or any(slots):
            for b in range(8):
                if slots[b] == 0 and queue:
                    slots[b] = limits[queue.pop(0)] - 1
            step = min(x for x in slots if x > 0)
            slots = [max(x - step, 0) for x in slots]
            n += step
        return n

    ideal = sum(x - 1 for x in limits) / 8
    assert launches(lpt) < launches(admission_order(limits)), (launches(lpt), launches(admission_order(limits)))
    assert launches(lpt) <= ideal + 127
