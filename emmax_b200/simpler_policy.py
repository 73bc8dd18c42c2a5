"""
SimplerEnv policy wrapper on top of the accelerated model.

Behavioural mirror of /root/reference/experiments/SimplerEnv-OpenVLA/simpler_env/policies/openvla/openvla_model.py:12-147
(`OpenVLAInference`: constructor arguments, `reset`, `step` -> (raw_action, action) with the euler -> axis-angle conversion, the
google-robot sticky-gripper state machine and the widowx binarised gripper), with the model and processor coming from this package
instead of `transformers` (the reference hard-codes the hub id "openvla/openvla-7b" at :40; here `saved_model_path` is used for both).
`transforms3d` is not a dependency: `euler2axangle` restates transforms3d's `euler2quat(..., 'sxyz')` + `quat2axangle`.
The plotting helper `visualize_epoch` (:149-185) is evaluation tooling and is not mirrored.
"""

from __future__ import annotations

import math
import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
from PIL import Image

_FLOAT_EPS = float(np.finfo(np.float64).eps)


def euler2axangle(ai: float, aj: float, ak: float) -> Tuple[np.ndarray, float]:
    """transforms3d.euler.euler2axangle(ai, aj, ak, axes='sxyz'): static-frame x, y, z rotations -> (unit axis, angle)."""
    ai, aj, ak = ai / 2.0, aj / 2.0, ak / 2.0
    ci, si, cj, sj, ck, sk = math.cos(ai), math.sin(ai), math.cos(aj), math.sin(aj), math.cos(ak), math.sin(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    w, x, y, z = cj * cc + sj * ss, cj * sc - sj * cs, cj * ss + sj * cc, cj * cs - sj * sc
    nq = w * w + x * x + y * y + z * z
    if not np.isfinite(nq):
        return np.array([1.0, 0, 0]), float("nan")
    if nq < _FLOAT_EPS**2:
        return np.array([1.0, 0, 0]), 0.0
    if nq != 1:
        s = math.sqrt(nq)
        w, x, y, z = w / s, x / s, y / s, z / s
    len2 = x * x + y * y + z * z
    if len2 < (_FLOAT_EPS * 3) ** 2:  # identity rotation
        return np.array([1.0, 0, 0]), 0.0
    theta = 2 * math.acos(max(min(w, 1), -1))
    return np.array([x, y, z]) / math.sqrt(len2), theta


class OpenVLAInference:
    def __init__(self, saved_model_path: str = "openvla/openvla-7b", unnorm_key: Optional[str] = None, policy_setup: str = "widowx_bridge",
                 horizon: int = 1, pred_action_horizon: int = 1, exec_horizon: int = 1, image_size: List[int] = [224, 224],
                 action_scale: float = 1.0, device: str = "cuda:0", vla: Any = None, processor: Any = None) -> None:  # fmt: skip
        os.environ["TOKENIZERS_PARALLELISM"] = "false"
        if policy_setup == "widowx_bridge":
            unnorm_key = "bridge_orig" if unnorm_key is None else unnorm_key
            self.sticky_gripper_num_repeat = 1
        elif policy_setup == "google_robot":
            unnorm_key = "fractal20220817_data" if unnorm_key is None else unnorm_key
            self.sticky_gripper_num_repeat = 15
        else:
            raise NotImplementedError(
                f"Policy setup {policy_setup} not supported for octo models. The other datasets can be found in the huggingface config.json file."
            )
        self.policy_setup, self.unnorm_key, self.device = policy_setup, unnorm_key, device
        if processor is None or vla is None:  # (tests inject stand-ins)
            from . import AutoModelForVision2Seq, AutoProcessor

            processor = AutoProcessor.from_pretrained(saved_model_path, trust_remote_code=True)
            vla = AutoModelForVision2Seq.from_pretrained(saved_model_path, attn_implementation="flash_attention_2", torch_dtype=torch.bfloat16,
                                                         low_cpu_mem_usage=True, trust_remote_code=True).to(device)  # fmt: skip
        self.processor, self.vla = processor, vla
        self.image_size, self.action_scale = image_size, action_scale
        self.horizon, self.pred_action_horizon, self.exec_horizon = horizon, pred_action_horizon, exec_horizon
        self.task, self.task_description = None, None
        self.reset(None)

    def reset(self, task_description: Optional[str]) -> None:
        self.task_description = task_description
        self.num_image_history = 0
        self.sticky_action_is_on = False
        self.gripper_action_repeat = 0
        self.sticky_gripper_action = 0.0
        self.previous_gripper_action = None

    def step(self, image: np.ndarray, task_description: Optional[str] = None, *args: Any, **kwargs: Any) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
        """image: uint8 (H, W, 3). Returns (raw_action, action) with the keys of the reference (:78-88)."""
        if task_description is not None and task_description != self.task_description:
            self.reset(task_description)
        assert image.dtype == np.uint8
        image = self._resize_image(image)
        inputs = self.processor(task_description, Image.fromarray(image)).to(self.device, dtype=torch.bfloat16)
        raw_actions = self.vla.predict_action(**inputs, unnorm_key=self.unnorm_key, do_sample=False)[None]
        return self.postprocess(raw_actions)

    def postprocess(self, raw_actions: np.ndarray) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
        """Everything of `step` after the model call (:103-145): split the 7-DoF vector, euler -> axis-angle, gripper handling."""
        vec = np.asarray(raw_actions)[0]
        raw_action = {"world_vector": np.array(vec[:3]), "rotation_delta": np.array(vec[3:6]), "open_gripper": np.array(vec[6:7])}  # 1 = open
        axis, angle = euler2axangle(*np.asarray(raw_action["rotation_delta"], dtype=np.float64))
        action: Dict[str, np.ndarray] = {
            "world_vector": raw_action["world_vector"] * self.action_scale,
            "rot_axangle": axis * angle * self.action_scale,
        }
        if self.policy_setup == "widowx_bridge":  # binarised absolute command: +1 open, -1 close
            action["gripper"] = 2.0 * (raw_action["open_gripper"] > 0.5) - 1.0
        else:  # google_robot: relative command with a sticky repeat
            action["gripper"] = self._sticky_gripper(raw_action["open_gripper"])
        action["terminate_episode"] = np.array([0.0])
        return raw_action, action

    def _sticky_gripper(self, opening: np.ndarray) -> np.ndarray:
        """google_robot gripper (:121-140): the command is the CHANGE of the opening; a change larger than 0.5 latches and is repeated for
        `sticky_gripper_num_repeat` steps (the first latched step included), then the latch clears."""
        delta = np.array([0]) if self.previous_gripper_action is None else self.previous_gripper_action - opening
        self.previous_gripper_action = opening
        if not self.sticky_action_is_on and np.abs(delta) > 0.5:
            self.sticky_action_is_on, self.sticky_gripper_action = True, delta
        if self.sticky_action_is_on:
            self.gripper_action_repeat += 1
            delta = self.sticky_gripper_action
        if self.gripper_action_repeat == self.sticky_gripper_num_repeat:
            self.sticky_action_is_on, self.gripper_action_repeat, self.sticky_gripper_action = False, 0, 0.0
        return delta

    def _resize_image(self, image: np.ndarray) -> np.ndarray:
        import cv2 as cv  # the reference resizes with OpenCV's area interpolation (:147-149)

        return cv.resize(image, tuple(self.image_size), interpolation=cv.INTER_AREA)


class BatchedOpenVLAInference:
    """N SimplerEnv environments behind ONE model: `step(images, task_descriptions)` runs the N requests through `predict_action_batch`
    (8 sequences per pass over the weights, emx_decode_batch_step) and then each environment's own post-processing state machine
    (sticky gripper, task reset) exactly as N independent `OpenVLAInference` objects would. The reference evaluates its environments one
    by one because its cached generation asserts batch size 1 (modeling_prismatic.py:326, :460-463)."""

    def __init__(self, n_envs: int, **kwargs: Any) -> None:
        first = OpenVLAInference(**kwargs)
        shared = dict(kwargs, vla=first.vla, processor=first.processor)
        self.envs: List[OpenVLAInference] = [first] + [OpenVLAInference(**shared) for _ in range(n_envs - 1)]
        self.vla, self.processor = first.vla, first.processor

    def reset(self, task_descriptions: List[Optional[str]]) -> None:
        for env, t in zip(self.envs, task_descriptions):
            env.reset(t)

    def step(self, images: List[np.ndarray], task_descriptions: Optional[List[Optional[str]]] = None) -> List[Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]]:
        assert len(images) == len(self.envs)
        tasks = task_descriptions if task_descriptions is not None else [None] * len(self.envs)
        inputs = []
        for env, image, task in zip(self.envs, images, tasks):
            if task is not None and task != env.task_description:
                env.reset(task)
            assert image.dtype == np.uint8
            inputs.append(self.processor(task, Image.fromarray(env._resize_image(image))).to(env.device, dtype=torch.bfloat16))
        raw = self.vla.predict_action_batch(inputs, unnorm_key=self.envs[0].unnorm_key)
        return [env.postprocess(raw[i][None]) for i, env in enumerate(self.envs)]
