"""
Parsing of generated grounded-CoT text into action vectors.

Behavioural mirror of /root/reference/prismatic/vla/solver.py:42-137 (`extract_movement_plan`,
`extract_action_policies`) — including the fall-backs the reference has: a missing "POLICIES:" key parses the whole
text as the policy line (:117-119), and any exception yields `[[0]*7]` with the full text as reasoning (:133-135) —
which, as written in the reference, includes a `;`-piece that does not give 7 values (:129-131: a list gets
`.tolist()`-ed, raising inside the `try`). The evaluation helpers of the reference Solver
(:15-40, :139-186) are training-time metrics and are out of scope.

Unlike the reference there is no module-level `AutoTokenizer.from_pretrained("meta-llama/Llama-2-7b-hf")` (:188):
the solver is built from the model's own tokenizer.
"""

from __future__ import annotations

from collections import defaultdict
from typing import Any, List, Optional, Tuple

import numpy as np

from .action_tokenizer import ActionTokenizer

_MOVE_TABLE = {
    # phrase -> (sign, axis); same table as solver.py:63-82
    "move_backward": (-1, "y"), "move_forward": (1, "y"), "move_right": (-1, "x"), "move_left": (1, "x"),
    "move_downward": (-1, "z"), "move_upward": (1, "z"), "roll_downward": (-1, "ox"), "roll_upward": (1, "ox"),
    "swing_downward": (-1, "ox"), "swing_upward": (1, "ox"), "pitch_downward": (-1, "oy"), "pitch_upward": (1, "oy"),
    "yaw_downward": (-1, "oz"), "yaw_upward": (1, "oz"), "rotate_clockwise": (-1, "oz"),
    "rotate_counterclockwise": (1, "oz"), "close_gripper": (-1, "grip"), "open_gripper": (1, "grip"),
}  # fmt: skip
_AXES = ["x", "y", "z", "ox", "oy", "oz", "grip"]


def _first_nonblank_line(s: str) -> str:
    lines = [ln for ln in s.split("\n") if len(ln.strip()) != 0]
    return lines[0].strip()


class Solver:
    coordinates_key = "NEXT GRIPPER:"
    movement_key = "MOVEMENT:"
    policy_key = "POLICIES:"

    def __init__(self, action_tokenizer: Optional[ActionTokenizer] = None, verbose: bool = True) -> None:
        self.action_tokenizer, self.verbose = action_tokenizer, verbose

    def _ids(self, text: str) -> np.ndarray:
        return np.array(self.action_tokenizer.tokenizer(text, add_special_tokens=False).input_ids)

    def extract_movement_plan(self, text: str) -> Tuple[Optional[bool], np.ndarray]:
        require_unorm = None
        try:
            line = _first_nonblank_line(text[text.index(self.movement_key) + len(self.movement_key) :])
            if "gripper" not in line:  # tokenised, normalised form
                require_unorm = True
                vals = self.action_tokenizer.decode_token_ids_to_actions(self._ids(line))[1:8]
                assert len(vals) == 7
                movement = vals
            else:  # plain-text form: "move_left 3;open gripper;..."
                require_unorm = False
                acc = defaultdict(int)
                for item in [o for o in line.split(";") if len(o) > 0][:7]:
                    words = item.split()
                    sign, axis = _MOVE_TABLE["_".join(words[:2])]
                    if "o" in axis:
                        scale = 1e-3
                    elif "grip" in axis:
                        scale = 1
                    else:
                        scale = 1 / 180 * np.pi
                    level = round("open" in item) if "grip" in axis else int(words[2])
                    acc[axis] += sign * scale * level
                movement = [acc[a] for a in _AXES]
        except Exception:
            movement = [-100] * 7
        return require_unorm, np.array(movement)

    def extract_action_policies(self, text: str) -> Tuple[List[List[float]], str]:
        try:
            if self.policy_key in text:
                cut = text.index(self.policy_key)
                line = _first_nonblank_line(text[cut + len(self.policy_key) :])
                remain = text[:cut]
            else:
                line, remain = text.strip(), ""
            out: List[List[float]] = []
            for piece in line.split(";"):
                vals = self.action_tokenizer.decode_token_ids_to_actions(self._ids(piece))
                vals = vals[1:][:7]  # first id is the SentencePiece dummy prefix
                if len(vals) != 7:
                    # The reference assigns a python list here and then calls `.tolist()` on it (:129-131), which
                    # raises and lands in the blanket `except` below: ONE wrong-length piece zeroes the whole result.
                    raise ValueError("policy piece does not hold 7 action tokens")
                out.append(vals.tolist())
        except Exception:
            out, remain = [[0] * 7], text
        return out, remain


def unnormalize(values: Any, stats: dict, low_key: str = "q01", high_key: str = "q99") -> np.ndarray:
    """`where(mask, 0.5*(a+1)*(hi-lo)+lo, a)` — prismatic.py:674-685 == openvla.py:95-102 == modeling_prismatic.py:528-535."""
    mask = stats.get("mask", np.ones_like(stats[low_key], dtype=bool))
    hi, lo = np.array(stats[high_key]), np.array(stats[low_key])
    a = np.array(values)
    return np.where(mask, 0.5 * (a + 1) * (hi - lo) + lo, a)
