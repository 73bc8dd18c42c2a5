"""
Prompt construction.

`PurePromptBuilder` mirrors /root/reference/prismatic/models/backbones/llm/prompting/base_prompter.py:28-73
("In: {msg}\\nOut: " turns, trailing whitespace removed by `get_prompt`). The task templates are the ones the robot
loops build (experiments/robot/bridge/run_bridgev2_eval.py:168, run_bridgev2_position_eval.py:157,
experiments/robot/openvla_utils.py:163,203-209).
"""

from __future__ import annotations

from typing import Optional


class PurePromptBuilder:
    def __init__(self, model_family: str = "prismatic", system_prompt: Optional[str] = None) -> None:
        self.model_family, self.system_prompt = model_family, system_prompt
        self.bos, self.eos = "<s>", "</s>"
        self.prompt, self.turn_count = "", 0

    def _human(self, msg: str) -> str:
        return f"In: {msg}\nOut: "

    def _gpt(self, msg: str) -> str:
        return f"{msg if msg != '' else ' '}{self.eos}"

    def add_turn(self, role: str, message: str) -> str:
        expected = "human" if self.turn_count % 2 == 0 else "gpt"
        assert role == expected
        message = message.replace("<image>", "").strip()
        wrapped = self._human(message) if expected == "human" else self._gpt(message)
        self.prompt += wrapped
        self.turn_count += 1
        return wrapped

    def get_potential_prompt(self, message: str) -> str:
        return (self.prompt + self._human(message)).removeprefix(self.bos).rstrip()

    def get_prompt(self) -> str:
        return self.prompt.removeprefix(self.bos).rstrip()


def emma_x_task_message(task_label: str) -> str:
    """The user message of the Emma-X robot loop (run_bridgev2_eval.py:168)."""
    return f"What action should the robot take to achieve the instruction\nINSTRUCTION: \n{task_label}\n"


def emma_x_prompt(task_label: str) -> str:
    pb = PurePromptBuilder("prismatic")
    pb.add_turn("human", emma_x_task_message(task_label))
    return pb.get_prompt()


def openvla_prompt(task_label: str) -> str:
    """Single-step OpenVLA prompt (openvla_utils.py:163)."""
    return f"In: What action should the robot take to {task_label.lower()}?\nOut:"
