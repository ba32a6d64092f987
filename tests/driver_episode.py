"""Runs a few training episodes of the reference's UNMODIFIED driver logic (TensorRL_*.py: one_episode + the DQN agents
from /root/reference/agents) against either the reference's own environments (--impl reference; qulacs / qiskit supplied
by the stand-ins of tests/golden/make_env_golden.py) or this repository's drop-in environments (--impl b200, energies
from the oracle-backed shim so that no GPU is needed), and prints the episode records as one JSON line.
Used by tests/test_driver_integration.py; needs /root/reference (build container only)."""
import argparse
import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"

DRIVERS = {   # driver module -> environment module it imports
    "TensorRL_fixed_noiseless": "environment_qulacs_TN_notin_agent",
    "TensorRL_training_and_structureRL_noiseless": "environment_qulacs",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", required=True, choices=["reference", "b200"])
    ap.add_argument("--driver", default="TensorRL_fixed_noiseless")
    ap.add_argument("--experiment", default="TensorRL_fixed/")
    ap.add_argument("--config", default="BEH26q_TNbond2")
    ap.add_argument("--episodes", type=int, default=3)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--global-iters", type=int, default=60)
    args = ap.parse_args()
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    if args.impl == "reference":
        sys.path.insert(0, os.path.join(HERE, "golden"))
        import make_env_golden  # noqa: F401  installs the qulacs / qiskit stand-ins, puts /root/reference on sys.path
    else:
        import tensorrl_qas_b200.environments as envs
        envs.install()
        import env_fixture as fx
        tn = DRIVERS[args.driver].startswith("environment_qulacs_TN")
        mod = __import__(f"tensorrl_qas_b200.environments.{DRIVERS[args.driver]}", fromlist=["CircuitEnv"])
        mod.CircuitEnv.vc = fx.oracle_vc(tn, False, False)
        mod.CircuitEnv._simulate_init_circuit = lambda self: fx.oracle_statevector(self.tenor_circ)
        sys.path.insert(0, REF)   # the driver script and the agents package (environments.* is already aliased in sys.modules)
    os.chdir(REF)              # data paths are relative to the working directory
    import warnings
    warnings.simplefilter("ignore")
    driver = __import__(args.driver)
    from environments.utils.utils import get_config
    import agents

    conf = get_config(args.experiment, f"{args.config}.cfg")
    conf["non_local_opt"]["global_iters"] = args.global_iters
    device = torch.device("cpu")
    driver.conf, driver.device = conf, device
    random.seed(args.seed)
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)

    env = driver.CircuitEnv(conf, device=device)
    agent = agents.__dict__[conf["agent"]["agent_type"]].__dict__[conf["agent"]["agent_class"]](
        conf, env.action_size, env.state_size, device)
    agent.saver = driver.Saver("/tmp", args.seed)
    if conf["agent"]["init_net"]:
        raise SystemExit("init_net cfgs need checkpoints")
    devnull = open(os.devnull, "w")
    real_stdout = sys.stdout
    sys.stdout = devnull       # the reference prints progress lines
    try:
        for e in range(args.episodes):
            driver.one_episode(e, env, agent, args.episodes)
    finally:
        sys.stdout = real_stdout
    rec = agent.saver.stats_file["train"]
    out = {str(e): {"actions": [int(a) for a in rec[e]["actions"]], "errors": [float(x) for x in rec[e]["errors"]],
                    "nfev": [int(x) for x in rec[e]["nfev"]], "reward": [float(x) for x in rec[e]["reward"]],
                    "done_threshold": float(rec[e]["done_threshold"])} for e in range(args.episodes)}
    out["env_class"] = f"{type(env).__module__}.{type(env).__name__}"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
