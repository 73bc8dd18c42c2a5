"""Freezes the outputs of the reference's own `PurePromptBuilder` for tests/test_detok_host.py (TEST INFRASTRUCTURE).
The file /root/reference/prismatic/models/backbones/llm/prompting/base_prompter.py has no third-party imports, so it is loaded by
path (importing the `prismatic` package itself fails here: draccus / timm / tensorflow are absent) and driven with multi-turn cases.
Usage (container with /root/reference):  python oracle/gen_golden_prompts.py"""
import importlib.util
import json
import os

REF = "/root/reference/prismatic/models/backbones/llm/prompting/base_prompter.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "prompt_golden.json")

CASES = [
    [("human", "What action should the robot take to achieve the instruction\nINSTRUCTION: \nput carrot in pot\n")],
    [("human", "  <image>\nWhat is in the image?  "), ("gpt", "A carrot. "), ("human", "Where?")],
    [("human", "hello"), ("gpt", "")],
    [("human", "<s>In: nested"), ("gpt", "ok"), ("human", ""), ("gpt", "  spaced  ")],
]

if __name__ == "__main__":
    spec = importlib.util.spec_from_file_location("ref_base_prompter", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = []
    for turns in CASES:
        pb = ref.PurePromptBuilder("prismatic")
        wrapped = [pb.add_turn(role, msg) for role, msg in turns]
        out.append({"turns": turns, "wrapped": wrapped, "prompt": pb.get_prompt(), "potential": pb.get_potential_prompt("next question"),
                    "turn_count": pb.turn_count})
    with open(OUT, "w") as f:
        json.dump({"source": REF + ":28-73", "cases": out}, f, indent=1)
    print("wrote", OUT)
