"""End-to-end environment steps per second: B `CircuitEnv` drop-ins (the reference's fixed BeH2-6q environment, inputs
rebuilt from tests/golden/env_golden.npz) taking the same number of random legal actions,
  serial    : one environment after the other, scipy's COBYLA            (= the reference's loop on the drop-in shims)
  lock-step : `LockstepEnvs` (one launch per COBYLA round), scipy's COBYLA / the native one (TQ_OPTIMIZER=native).
Prints one JSON line.  python profiles/bench_lockstep_envs.py [--envs 64] [--steps 6] [--serial-sample 8]"""
import argparse
import importlib
import json
import os
import sys
import tempfile
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=64)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--serial-sample", type=int, default=8)
    args = ap.parse_args()
    import torch
    import env_fixture as fx
    from tensorrl_qas_b200.VQAs import _backend
    from tensorrl_qas_b200.environments.utils import utils
    from tensorrl_qas_b200.lockstep import LockstepEnvs

    ep = fx.Episode("fixed_beh2")
    root = tempfile.mkdtemp()
    fx.materialize(root, ep)
    os.environ["TQ_DATA_ROOT"] = root
    mod = importlib.import_module("tensorrl_qas_b200.environments.environment_qulacs_TN_notin_agent")
    conf = ep.conf
    table = utils.dictionary_of_actions(conf["env"]["num_qubits"])
    rot_actions = [a for a, v in table.items() if v[2] < conf["env"]["num_qubits"]]   # rotations: something to optimise

    def fresh(n):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return [mod.CircuitEnv(conf, device=torch.device("cpu")) for _ in range(n)]

    def plan(b):
        rng = np.random.default_rng(1000 + b)
        return [int(rot_actions[int(rng.integers(len(rot_actions)))]) for _ in range(args.steps)]

    out = {"workload": f"fixed BeH2-6q CircuitEnv, {args.steps} rotation actions per environment, COBYLA maxiter "
                       f"{conf['non_local_opt']['global_iters']}"}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # serial, scipy
        os.environ["TQ_OPTIMIZER"] = "scipy"
        _backend.reset_backends()
        envs = fresh(args.serial_sample)
        t0 = time.perf_counter()
        nfev = 0
        for b, env in enumerate(envs):
            env.reset()
            for a in plan(b):
                env.step(list(table[a]))
                nfev += int(env.nfev)
        dt = time.perf_counter() - t0
        out["serial_scipy"] = {"envs": len(envs), "seconds": dt, "env_steps_per_s": len(envs) * args.steps / dt,
                               "evals": nfev, "ms_per_eval": 1e3 * dt / nfev}
        for opt in ("scipy", "native"):
            os.environ["TQ_OPTIMIZER"] = opt
            _backend.reset_backends()
            envs = fresh(args.envs)
            ls = LockstepEnvs(envs)
            t0 = time.perf_counter()
            ls.reset_all()
            nfev = 0
            for t in range(args.steps):
                ls.step_all([list(table[plan(b)[t]]) for b in range(args.envs)])
                nfev += sum(int(e.nfev) for e in envs)
            dt = time.perf_counter() - t0
            out[f"lockstep_{opt}"] = {"envs": args.envs, "seconds": dt, "env_steps_per_s": args.envs * args.steps / dt,
                                      "evals": nfev, "ms_per_eval": 1e3 * dt / nfev}
    out["speedup_lockstep_native_vs_serial_scipy"] = (out["lockstep_native"]["env_steps_per_s"] /
                                                      out["serial_scipy"]["env_steps_per_s"])
    out["speedup_lockstep_scipy_vs_serial_scipy"] = (out["lockstep_scipy"]["env_steps_per_s"] /
                                                     out["serial_scipy"]["env_steps_per_s"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
