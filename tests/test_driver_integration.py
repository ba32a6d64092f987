"""Drop-in integration: the reference's unmodified driver loop and DQN agents run for a few
episodes against (a) the reference's own environments and (b) this repository's environments with the same seed; the two
runs must make the same decisions and report the same numbers.  Needs /root/reference (build container); the GPU box
runs the GPU-backed replays of tests/test_env.py instead."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/agents"), reason="reference checkout not present")


def run(impl, *extra):
    out = subprocess.run([sys.executable, os.path.join(HERE, "driver_episode.py"), "--impl", impl, *extra],
                         capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


@pytest.mark.parametrize("extra", [
    ("--driver", "TensorRL_fixed_noiseless", "--experiment", "TensorRL_fixed/", "--config", "BEH26q_TNbond2",
     "--episodes", "3"),
    ("--driver", "TensorRL_training_and_structureRL_noiseless", "--experiment", "StructureRL/", "--config",
     "heisenberg_5q_TNbond2", "--episodes", "2", "--global-iters", "90"),
])
def test_unmodified_driver_makes_the_same_episodes(oracle, extra):
    ref = run("reference", *extra)
    new = run("b200", *extra)
    assert ref.pop("env_class").startswith("environments.")
    assert new.pop("env_class").startswith("tensorrl_qas_b200.environments.")
    assert ref.keys() == new.keys()
    for e in ref:
        assert new[e]["actions"] == ref[e]["actions"], f"episode {e}: action sequence"
        assert new[e]["nfev"] == ref[e]["nfev"]
        assert new[e]["errors"] == ref[e]["errors"] and new[e]["reward"] == ref[e]["reward"]
        assert new[e]["done_threshold"] == ref[e]["done_threshold"]
        assert len(ref[e]["actions"]) >= 1
