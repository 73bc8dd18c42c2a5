"""Freezes the outputs of the reference's OWN SimplerEnv policy class for tests/test_simpler_policy.py (TEST INFRASTRUCTURE).

Runs /root/reference/experiments/SimplerEnv-OpenVLA/simpler_env/policies/openvla/openvla_model.py in this container. The module imports
packages that are absent here (transforms3d) or irrelevant (matplotlib, the hub model); they are replaced by stand-ins BEFORE the import:
`transforms3d.euler.euler2axangle` by scipy (Rotation.from_euler('xyz').as_rotvec(), the same static-xyz convention), the model and
processor by stubs returning scripted 7-DoF actions. Only the class's own post-processing (:103-145: action split, axis-angle scaling,
sticky-gripper state machine, widowx binarisation, reset on a new task) is exercised.
Usage (container with /root/reference):  python oracle/gen_golden_simpler.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/experiments/SimplerEnv-OpenVLA/simpler_env/policies/openvla/openvla_model.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "simpler_policy_golden.json")


def _axangle(ai, aj, ak):
    from scipy.spatial.transform import Rotation

    v = Rotation.from_euler("xyz", [ai, aj, ak]).as_rotvec()
    n = float(np.linalg.norm(v))
    return (np.array([1.0, 0, 0]), 0.0) if n == 0 else (v / n, n)


t3d, t3e = types.ModuleType("transforms3d"), types.ModuleType("transforms3d.euler")
t3e.euler2axangle = _axangle
t3d.euler = t3e
sys.modules.update({"transforms3d": t3d, "transforms3d.euler": t3e})
if "matplotlib" not in sys.modules:
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mp, mpp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mp.pyplot = mpp
        sys.modules.update({"matplotlib": mp, "matplotlib.pyplot": mpp})
tr = types.ModuleType("transformers")
tr.AutoModelForVision2Seq = tr.AutoProcessor = object
sys.modules["transformers"] = tr

spec = importlib.util.spec_from_file_location("ref_openvla_model", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


class _Inputs(dict):
    def to(self, *a, **k):
        return self


class _VLA:
    def __init__(self, actions):
        self.actions, self.i = actions, 0

    def predict_action(self, unnorm_key=None, do_sample=False, **kw):
        a = np.asarray(self.actions[self.i], dtype=np.float64)
        self.i += 1
        return a


def run(setup: str, action_scale: float) -> dict:
    rng = np.random.default_rng(7 if setup == "widowx_bridge" else 11)
    n = 24
    acts = rng.uniform(-0.3, 0.3, (n, 7))
    grip = (np.sin(np.arange(n) * 0.9) > 0).astype(np.float64) * 0.9 + 0.05  # open/close toggles -> sticky gripper engages
    acts[:, 6] = grip
    tasks = ["put carrot on plate"] * 10 + ["stack the blocks"] * 14  # a new task resets the policy state
    pol = ref.OpenVLAInference.__new__(ref.OpenVLAInference)  # skip the hub download of the reference constructor
    pol.policy_setup = setup
    pol.unnorm_key = "bridge_orig" if setup == "widowx_bridge" else "fractal20220817_data"
    pol.sticky_gripper_num_repeat = 1 if setup == "widowx_bridge" else 15
    pol.processor, pol.vla = (lambda prompt, image: _Inputs()), _VLA(acts.tolist())
    pol.image_size, pol.action_scale = [224, 224], action_scale
    pol.task, pol.task_description, pol.num_image_history = None, None, 0
    pol.sticky_action_is_on, pol.gripper_action_repeat, pol.sticky_gripper_action, pol.previous_gripper_action = False, 0, 0.0, None
    steps = []
    img = np.zeros((256, 320, 3), dtype=np.uint8)
    for t in range(n):
        raw, act = pol.step(img, tasks[t])
        steps.append({"raw": {k: np.asarray(v, dtype=np.float64).tolist() for k, v in raw.items()},
                      "action": {k: np.asarray(v, dtype=np.float64).reshape(-1).tolist() for k, v in act.items()}})
    return {"action_scale": action_scale, "unnorm_key": pol.unnorm_key, "sticky_gripper_num_repeat": pol.sticky_gripper_num_repeat,
            "model_outputs": acts.tolist(), "tasks": tasks, "steps": steps}


if __name__ == "__main__":
    out = {"widowx_bridge": run("widowx_bridge", 1.0), "google_robot": run("google_robot", 0.75)}
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
