"""Free-running replay of the golden reference episodes on the GPU path (helper of tests/test_env.py and a small CLI):
no teacher forcing -- every step continues from the drop-in's own state.  Returns per-episode statistics of how far the
trajectory is from the reference's (same seeded actions; scipy's COBYLA on both sides).

    python tests/free_run_stats.py            # prints one JSON line per episode"""
import json
import sys
import warnings

import numpy as np


def free_run(key, tmp_path, monkeypatch):
    import torch
    import test_env as te
    from tensorrl_qas_b200.VQAs import _backend
    _backend.reset_backends()
    ep, env, table = te._make_env(key, tmp_path, monkeypatch, "gpu")
    d, n = ep.d, env.num_qubits
    env.reset()
    st = {"key": key, "steps": 0, "same_nfev": 0, "same_done": 0, "same_mask": 0, "same_gates": 0, "same_angles_f32": 0,
          "max_dE": 0.0, "max_dE_before_fork": 0.0, "max_dangle": 0.0, "final_dE": None, "first_fork_step": None,
          "cnot_count_equal": True, "rot_count_equal": True}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(ep.n_steps):
            st["same_mask"] += [int(a) for a in env.illegal_action_new()] == ep.illegal(i)
            obs, reward, done = env.step(list(table[int(d["action"][i])]))
            st["steps"] += 1
            same_gates = np.array_equal(env.state.numpy()[:, :n + 3], d["state"][i][:, :n + 3])
            st["same_gates"] += same_gates
            st["same_nfev"] += int(env.nfev) == int(d["nfev"][i])
            st["same_done"] += int(done) == int(d["done"][i])
            dE = abs(float(env.energy) - float(d["energy"][i]))
            st["max_dE"] = max(st["max_dE"], dE)
            st["final_dE"] = dE
            ang_same = np.array_equal(env.state.numpy(), d["state"][i])
            st["same_angles_f32"] += ang_same
            st["max_dangle"] = max(st["max_dangle"], float(np.abs(env.state.numpy()[:, n + 3:] - d["state"][i][:, n + 3:]).max()))
            if not ang_same and st["first_fork_step"] is None:
                st["first_fork_step"] = i
            if st["first_fork_step"] is None:   # same optimiser path so far: the energies are those of identical circuits
                st["max_dE_before_fork"] = max(st["max_dE_before_fork"], dE)
            t_mine, t_ref = env.state.numpy(), d["state"][i]
            st["cnot_count_equal"] &= int((t_mine[:, :n] != 0).sum()) == int((t_ref[:, :n] != 0).sum())
            st["rot_count_equal"] &= int((t_mine[:, n:n + 3] != 0).sum()) == int((t_ref[:, n:n + 3] != 0).sum())
            if done or int(d["done"][i]):
                break
    _backend.reset_backends()
    return st


if __name__ == "__main__":
    import os
    import tempfile
    import pytest
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import test_env as te
    mp = pytest.MonkeyPatch()
    for key in te.VARIANT:
        with tempfile.TemporaryDirectory() as tmp:
            print(json.dumps(free_run(key, tmp, mp)), flush=True)
    mp.undo()
