"""
Generate the committed golden fixtures under tests/golden/ (TEST INFRASTRUCTURE — see oracle/__init__.py).

Run in the BUILD container only (`python -m oracle.gen_golden`): it reads /root/reference, which does not exist on
the GPU box. Two kinds of fixture:

1. detok_golden.json — produced by the REFERENCE'S OWN CODE:
     * `ActionTokenizer` imported from /root/reference/prismatic/vla/action_tokenizer.py (file loaded by path; the
       `prismatic` package itself cannot be imported: `prismatic/__init__.py` pulls draccus/tensorflow),
     * `Solver` — the class body of /root/reference/prismatic/vla/solver.py executed without the module tail that
       downloads the gated Llama-2 tokenizer (solver.py:188-190),
   driven with `emmax_b200.tokenization.SyntheticLlamaTokenizer` (the real tokenizer is gated / offline).
   This pins the integer + fp64 de-tokeniser and the text parser.
2. tiny_vla_golden.npz — produced by the oracle restatement (oracle/model.py) on the tiny config with seeded weights
   and inputs: prompt ids, pixel values seed, greedy ids, per-step logits, patch features, action vectors. "Parity
   unpinned" for this part (no reference fixtures exist), it freezes the oracle against drift and gives the GPU
   tests a fixture that does not need the oracle at all.
"""

from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from emmax_b200.configuration import synthetic_norm_stats, tiny_config  # noqa: E402
from emmax_b200.synthetic import default_script, make_state_dict  # noqa: E402
from emmax_b200.tokenization import SyntheticLlamaTokenizer  # noqa: E402


def load_reference_detok():
    spec = importlib.util.spec_from_file_location("ref_action_tokenizer", f"{REF}/prismatic/vla/action_tokenizer.py")
    at = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(at)
    # make `from prismatic.vla.action_tokenizer import ActionTokenizer` resolve to the file above
    for name in ("prismatic", "prismatic.vla"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["prismatic.vla.action_tokenizer"] = at
    src = open(f"{REF}/prismatic/vla/solver.py").read()
    head = src.split("\ntokenizer = AutoTokenizer.from_pretrained", 1)[0]  # drop the network-bound module tail
    ns: dict = {}
    exec(compile(head, f"{REF}/prismatic/vla/solver.py", "exec"), ns)
    return at.ActionTokenizer, ns["Solver"]


def detok_cases(tok):
    a = [chr(0x4E00 + k) for k in range(256)]  # action code points, id = 31744 + k

    def pol(ks):
        return "".join(a[k] for k in ks)

    p1, p2 = pol([255, 128, 0, 1, 77, 200, 254]), pol([5, 6, 7, 8, 9, 10, 11])
    texts = {
        "two_policies": f"REASONING:\nlift the pot\nSUBTASK: x\n\nNEXT GRIPPER: [105, 74]\n\nMOVEMENT:\n{p2}\nPOLICIES:\n{p1};{p2}\n",
        "one_policy": f"MOVEMENT:\n{p1}\nPOLICIES:\n{p2}\n",
        "no_key": p1,
        "no_key_two": f"{p1};{p2}",
        "short_piece": f"POLICIES:\n{p1};{pol([1, 2, 3])}\n",
        "long_piece": f"POLICIES:\n{pol([1, 2, 3, 4, 5, 6, 7, 8, 9])}\n",
        "empty_after_key": "POLICIES:\n",
        "blank_lines_then_policy": f"POLICIES:\n\n   \n{p1}\n{p2}",
        "text_tokens_as_actions": "POLICIES:\nabcdefg\n",
        "trailing_semicolon": f"POLICIES:\n{p1};\n",
        "movement_text_form": "MOVEMENT:\nmove left 3;move upward 2;yaw upward 10;open gripper\nPOLICIES:\n" + p1,
        "movement_missing": "nothing here",
        "movement_bad_word": "MOVEMENT:\nfly sideways 3\n",
        "empty": "",
    }
    return texts


def gen_detok():
    RefActionTokenizer, RefSolver = load_reference_detok()
    tok = SyntheticLlamaTokenizer()
    rat = RefActionTokenizer(tok)
    solver = RefSolver(rat, verbose=False)
    ids = np.concatenate([np.arange(31700, 32064), np.array([0, 1, 2, 13, 29871, 15000])])
    out = {
        "source": "reference code: prismatic/vla/action_tokenizer.py + Solver class of prismatic/vla/solver.py @ /root/reference",
        "vocab_size": tok.vocab_size,
        "ids": ids.tolist(),
        "decoded": [float.hex(float(x)) for x in rat.decode_token_ids_to_actions(ids)],
        "action_token_begin_idx": rat.action_token_begin_idx,
        "bin_centers_hex": [float.hex(float(x)) for x in rat.bin_centers],
    }
    # encode direction: continuous -> text -> ids (action_tokenizer.py:38-47)
    rng = np.random.default_rng(7)
    acts = np.concatenate([rng.uniform(-1.2, 1.2, (6, 7)), np.array([[-1, -0.999, 0, 1e-9, 0.5, 0.999, 1.0]])])
    out["encode_actions"] = acts.tolist()
    out["encode_text"] = [rat(a) for a in acts]
    out["encode_batch_text"] = rat(acts)
    # solver parse
    cases = {}
    for name, text in detok_cases(tok).items():
        pol, remain = solver.extract_action_policies(text)
        req, mov = solver.extract_movement_plan(text)
        cases[name] = {
            "text": text,
            "policies_hex": [[float.hex(float(v)) for v in p] for p in pol],
            "remain": remain,
            "require_unorm": req,
            "movement_hex": [float.hex(float(v)) for v in np.asarray(mov, dtype=np.float64)],
        }
    out["solver_cases"] = cases
    # un-normalise (modeling_prismatic.py:528-535) evaluated with numpy exactly as written there
    stats = synthetic_norm_stats()["synthetic"]["action"]
    normalized = rat.decode_token_ids_to_actions(np.array([31999, 31872, 31745, 31744, 31900, 31800, 31750]))
    mask = stats.get("mask", np.ones_like(stats["q01"], dtype=bool))
    hi, lo = np.array(stats["q99"]), np.array(stats["q01"])
    un = np.where(mask, 0.5 * (normalized + 1) * (hi - lo) + lo, normalized)
    out["unnorm"] = {"stats": stats, "ids": [31999, 31872, 31745, 31744, 31900, 31800, 31750],
                     "normalized_hex": [float.hex(float(x)) for x in normalized], "actions_hex": [float.hex(float(x)) for x in un]}  # fmt: skip
    with open(os.path.join(OUT, "detok_golden.json"), "w") as f:
        json.dump(out, f, indent=1, ensure_ascii=False)
    print("wrote detok_golden.json:", len(cases), "solver cases")


def tiny_inputs(tok, n_new=40, seed=0):
    prompt = "In: What action should the robot take to achieve the instruction\nINSTRUCTION: \nput carrot in pot\nOut:"
    input_ids = tok(prompt, return_tensors="pt").input_ids
    rng = np.random.default_rng(seed)
    image = rng.integers(0, 256, (224, 224, 3), dtype=np.uint8)
    script = default_script(tok, n_new, seed=seed)
    return prompt, input_ids, image, script


def gen_tiny():
    from PIL import Image

    from emmax_b200.processing import PrismaticImageProcessor
    from oracle.model import OracleVLA

    tok = SyntheticLlamaTokenizer()
    cfg = tiny_config()
    prompt, input_ids, image, script = tiny_inputs(tok)
    sd = make_state_dict(cfg, seed=0, script=script, script_prev=int(input_ids[0, -1]))
    pv = PrismaticImageProcessor()(Image.fromarray(image), return_tensors="pt")["pixel_values"]
    m = OracleVLA.from_state_dict(cfg, sd, dtype=torch.bfloat16)
    pvb = pv.to(torch.bfloat16)
    ids, logits = m.generate(input_ids, pvb, len(script), return_logits=True)
    with torch.inference_mode():
        feats = m.vision_backbone(pvb)
        proj = m.projector(feats)
    np.savez_compressed(
        os.path.join(OUT, "tiny_vla_golden.npz"),
        input_ids=input_ids.numpy(), image=image, pixel_values=pv.numpy(), script=np.array(script),
        generated_ids=ids.numpy(), step_logits=logits.float().numpy().astype(np.float32),
        patch_features=feats.float().numpy(), projected=proj.float().numpy(),
        action_pred=m.predict_action(input_ids, pvb),
    )  # fmt: skip
    new = ids[0, input_ids.shape[1] :].tolist()
    print("wrote tiny_vla_golden.npz; greedy follows script:", new == script, "| text:", repr(tok.decode(new, skip_special_tokens=True))[:80])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    gen_detok()
    gen_tiny()
