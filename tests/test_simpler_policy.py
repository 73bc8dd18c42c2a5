"""SimplerEnv policy wrapper pinned against the reference's own class: the source of
/root/reference/experiments/SimplerEnv-OpenVLA/simpler_env/policies/openvla/openvla_model.py is executed in this container with
stand-ins for the packages that are absent (transforms3d) or irrelevant (the hub model), and its `step` outputs are frozen in
tests/golden/simpler_policy_golden.json by oracle/gen_golden_simpler.py. Here the mirror must reproduce them exactly, and the
euler -> axis-angle restatement is cross-checked against scipy."""

import json
import os

import numpy as np
import pytest

from emmax_b200.simpler_policy import OpenVLAInference, euler2axangle


class _Inputs(dict):
    def to(self, *a, **k):
        return self


class _Proc:
    def __call__(self, prompt, image):
        return _Inputs()


class _VLA:
    def __init__(self, actions):
        self.actions, self.i = actions, 0

    def predict_action(self, unnorm_key=None, do_sample=False, **kw):
        a = np.asarray(self.actions[self.i], dtype=np.float64)
        self.i += 1
        return a


def test_euler2axangle_matches_scipy():
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(0)
    for r, p, y in rng.uniform(-1.2, 1.2, (200, 3)):
        ax, ang = euler2axangle(r, p, y)
        assert np.allclose(ax * ang, Rotation.from_euler("xyz", [r, p, y]).as_rotvec(), atol=1e-12)
    ax, ang = euler2axangle(0.0, 0.0, 0.0)
    assert ang == 0.0 and ax.tolist() == [1.0, 0.0, 0.0]


@pytest.mark.parametrize("setup", ["widowx_bridge", "google_robot"])
def test_step_matches_the_reference_class(golden_dir, setup):
    g = json.load(open(os.path.join(golden_dir, "simpler_policy_golden.json")))[setup]
    pol = OpenVLAInference(policy_setup=setup, action_scale=g["action_scale"], vla=_VLA(g["model_outputs"]), processor=_Proc(), device="cpu")
    assert pol.unnorm_key == g["unnorm_key"] and pol.sticky_gripper_num_repeat == g["sticky_gripper_num_repeat"]
    img = np.zeros((256, 320, 3), dtype=np.uint8)
    for t, want in enumerate(g["steps"]):
        raw, act = pol.step(img, g["tasks"][t])
        for k in ("world_vector", "rotation_delta", "open_gripper"):
            assert np.array_equal(np.asarray(raw[k], dtype=np.float64), np.asarray(want["raw"][k])), (t, k)
        for k in ("world_vector", "rot_axangle", "gripper", "terminate_episode"):
            assert np.allclose(np.asarray(act[k], dtype=np.float64).reshape(-1), np.asarray(want["action"][k]).reshape(-1), rtol=0, atol=1e-12), (t, k)
    with pytest.raises(NotImplementedError):
        OpenVLAInference(policy_setup="aloha", vla=_VLA([]), processor=_Proc())


def test_batched_policy_equals_independent_policies(golden_dir):
    """BatchedOpenVLAInference over N environments == N independent OpenVLAInference objects fed the same model outputs (per-environment
    sticky-gripper state, task resets), with the model called ONCE per step through predict_action_batch."""
    from emmax_b200 import BatchedOpenVLAInference

    g = json.load(open(os.path.join(golden_dir, "simpler_policy_golden.json")))["google_robot"]
    outs = np.asarray(g["model_outputs"], dtype=np.float64)
    n_env, T = 3, len(outs)
    # environment e sees the golden action stream shifted by e steps
    streams = [np.roll(outs, -e, axis=0) for e in range(n_env)]

    class _BatchVLA:
        calls, t = 0, 0

        def predict_action_batch(self, inputs, unnorm_key=None):
            assert len(inputs) == n_env and unnorm_key == g["unnorm_key"]
            self.calls += 1
            out = np.stack([streams[e][self.t] for e in range(n_env)])
            self.t += 1
            return out

    vla = _BatchVLA()
    pol = BatchedOpenVLAInference(n_env, policy_setup="google_robot", action_scale=g["action_scale"], vla=vla, processor=_Proc(), device="cpu")
    singles = [OpenVLAInference(policy_setup="google_robot", action_scale=g["action_scale"], vla=_VLA(streams[e]), processor=_Proc(), device="cpu") for e in range(n_env)]
    img = np.zeros((256, 320, 3), dtype=np.uint8)
    for t in range(T):
        tasks = [g["tasks"][t]] * n_env
        got = pol.step([img] * n_env, tasks)
        for e in range(n_env):
            raw, act = singles[e].step(img, tasks[e])
            for k in raw:
                assert np.array_equal(got[e][0][k], raw[k]), (t, e, k)
            for k in act:
                assert np.array_equal(np.asarray(got[e][1][k]), np.asarray(act[k])), (t, e, k)
    assert vla.calls == T
