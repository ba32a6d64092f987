"""Generates tests/golden/env_golden.npz: episodes of the REFERENCE's own CircuitEnv classes
(/root/reference/environments/environment_qulacs*.py, imported unmodified) driven by seeded random legal actions.

The reference environments import qiskit and qulacs, neither of which exists in the build container.  They are
supplied here as thin stand-ins: `qulacs` by oracle/np_oracle.py (install_as_qulacs) and `qiskit` by the stub below
(QPY reading and ASAP layering from tensorrl_qas_b200.loaders, Statevector through the oracle, reverse_qargs as a
bit-reversal permutation; the Qubit repr mimics qiskit 2.0.0's `<Qubit register=(n, "q"), index=i>` because the
reference parses it as text, environments/environment_qulacs.py:293-297).  Everything else that runs is the
reference's code: tensor encoding of the MPS circuit, gate placement, the one-step optimisation lag, the COBYLA
call, float32 write-backs, reward, termination, curriculum and the stateful illegal-action mask.

What the file pins per step: the illegal-action list the driver would see, the action taken, reward, done flag,
energy, nfev, the optimised angles and the full state tensor.  Parity with qulacs itself stays unpinned (see
oracle/np_oracle.py).

Run in the build container only (needs /root/reference):   python tests/golden/make_env_golden.py
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import np_oracle  # noqa: E402
from tensorrl_qas_b200 import loaders  # noqa: E402

np_oracle.install_as_qulacs()


def _unused_qulacs_gate(name):
    def factory(*a, **k):
        raise NotImplementedError(f"qulacs.gate.{name} is imported by the reference but never called on the hot path")
    return factory


# environments/VQAs/VQE_qulacs_TN_notin_RL_noise.py:2 imports these three names without using them
for _name in ("BitFlipNoise", "DephasingNoise", "IndependentXZNoise"):
    setattr(sys.modules["qulacs.gate"], _name, _unused_qulacs_gate(_name))


# ------------------------------------------------------------------------------------------- qiskit stand-in ----
class _Qubit:
    def __init__(self, n, index):
        self._n, self._index = n, index

    def __repr__(self):
        return f'<Qubit register=({self._n}, "q"), index={self._index}>'


class _Op:
    def __init__(self, name, params):
        self.name, self.params = name, params


class _Node:
    def __init__(self, n, name, qubits, angle):
        self.op = _Op(name, [] if angle is None else [angle])
        self.qargs = tuple(_Qubit(n, q) for q in qubits)


class _LayerGraph:
    def __init__(self, nodes):
        self._nodes = nodes

    def op_nodes(self):
        return list(self._nodes)


class _Dag:
    def __init__(self, circ):
        self.circ = circ

    def layers(self):
        for layer in self.circ._ic.layers():
            yield {"graph": _LayerGraph([_Node(self.circ._ic.n_qubits, *op) for op in layer])}


class _Circuit:
    def __init__(self, ic):
        self._ic = ic

    def depth(self):
        return self._ic.depth()


class _Statevector:
    def __init__(self, circ):
        gl = loaders.init_circuit_gatelist(circ._ic, parametric=False)
        self.data = np_oracle.run_circuit(circ._ic.n_qubits, gl.tuples(), np.zeros(0))


class _Operator:
    def __init__(self, m):
        self._m = np.asarray(m, dtype=np.complex128)

    def reverse_qargs(self):
        return _Operator(loaders.reverse_qargs(self._m))

    def to_matrix(self):
        return self._m


def install_qiskit_stub():
    qk = types.ModuleType("qiskit")
    qk.__version__ = "2.0.0"
    qpy = types.ModuleType("qiskit.qpy")
    qpy.load = lambda f: [_Circuit(loaders.load_qpy_circuit(f.name))]
    conv = types.ModuleType("qiskit.converters")
    conv.circuit_to_dag = _Dag
    qi = types.ModuleType("qiskit.quantum_info")
    qi.Statevector, qi.Operator = _Statevector, _Operator
    qk.qpy, qk.converters, qk.quantum_info = qpy, conv, qi
    for name, mod in (("qiskit", qk), ("qiskit.qpy", qpy), ("qiskit.converters", conv), ("qiskit.quantum_info", qi)):
        sys.modules[name] = mod


install_qiskit_stub()

# the fixed environments call COBYLA with an EMPTY parameter vector on their first step(s); scipy's
# Fortran COBYLA (what the authors ran) tolerated that, the pyprima port in this container raises.  The stand-in
# below evaluates the cost once and reports nfev = 1 for an empty x0 -- the behaviour the drop-in implements -- and
# defers to scipy otherwise.
import scipy.optimize  # noqa: E402

_scipy_minimize = scipy.optimize.minimize


def _minimize_allowing_empty(fun, x0, *args, **kw):
    x0 = np.asarray(x0)
    if x0.size == 0:
        return scipy.optimize.OptimizeResult(x=np.zeros(0), fun=fun(x0), nfev=1, success=True)
    return _scipy_minimize(fun, x0, *args, **kw)


scipy.optimize.minimize = _minimize_allowing_empty
from environments.utils.utils import get_config, dictionary_of_actions  # noqa: E402
from environments.utils import utils_topology_restrict as utr  # noqa: E402
import environments.environment_qulacs as env_in  # noqa: E402
import environments.environment_qulacs_noise as env_in_noise  # noqa: E402
import environments.environment_qulacs_TN_notin_agent as env_fixed  # noqa: E402
import environments.environment_qulacs_TN_notin_agent_noise as env_fixed_noise  # noqa: E402
import environments.environment_qulacs_TN_notin_agent_noise_restricted as env_fixed_restricted  # noqa: E402

# episode key -> (env module, cfg dir, cfg name, overrides, max steps, seed)
EPISODES = {
    "fixed_h2o8": (env_fixed, "TensorRL_fixed/", "H2O8q_TNbond2", {}, 12, 11),
    "fixed_beh2": (env_fixed, "TensorRL_fixed/", "BEH26q_TNbond2", {}, 24, 12),
    "fixed_heis5": (env_fixed, "TensorRL_fixed/", "heisenberg_5q_TNbond2", {}, 10, 13),
    "trainable_beh2": (env_in, "TensorRL_trainable/", "BEH26q_TNbond2", {"global_iters": 150}, 6, 14),
    "structure_heis5": (env_in, "StructureRL/", "heisenberg_5q_TNbond2", {"global_iters": 120}, 6, 15),
    "trainable_h2o8_angles": (env_in, "TensorRL_trainable/", "H2O8q_TNbond2", {"global_iters": 25, "angles": 1}, 4, 16),
    "noise_trainable_h2o8": (env_in_noise, "TensorRL_trainable/", "H2O8q_TNbond2_noise", {"global_iters": 25}, 4, 17),
    "noise_fixed_h2o8": (env_fixed_noise, "TensorRL_fixed/", "H2O8q_TNbond2_noise", {"global_iters": 200}, 8, 18),
    "restricted_h2o8": (env_fixed_restricted, "TensorRL_fixed/", "H2O8q_TNbond2_noise_restricted",
                        {"global_iters": 200, "n_shots": 1000}, 8, 19),
}


def run_episode(mod, cfg_dir, cfg, overrides, max_steps, seed):
    conf = get_config(cfg_dir, f"{cfg}.cfg", path=os.path.join(REF, "configuration_files"))
    for k, v in overrides.items():
        for section in conf.values():
            if k in section:
                section[k] = v
    os.chdir(REF)  # the reference's data paths are CWD-relative
    np.random.seed(seed)
    torch.manual_seed(seed)
    np_oracle.seed(seed)
    rng = np.random.default_rng(seed)
    env = mod.CircuitEnv(conf, device=torch.device("cpu"))
    if mod is env_fixed_restricted:
        table = utr.dictionary_of_actions_hexagon_connectivity(env.num_qubits)  # the agent's (forward) dictionary
    else:
        table = dictionary_of_actions(env.num_qubits)
    rec = {k: [] for k in ("illegal", "action", "reward", "done", "energy", "nfev", "error", "opt_ang", "state",
                           "obs_sum", "done_threshold")}
    obs = env.reset()
    rec_first = float(env.prev_energy)
    obs0 = obs.numpy().copy()
    for _ in range(max_steps):
        ill = env.illegal_action_new()           # the driver asks before acting (TensorRL_fixed_noiseless.py:118)
        legal = [a for a in table if a not in ill]
        a = int(legal[int(rng.integers(len(legal)))])
        obs, reward, done = env.step(list(table[a]))
        rec["illegal"].append(json.dumps([int(i) for i in ill]))
        rec["action"].append(a)
        rec["reward"].append(float(reward))
        rec["done"].append(int(done))
        rec["energy"].append(float(env.energy))
        rec["nfev"].append(int(env.nfev))
        rec["error"].append(float(env.error))
        rec["opt_ang"].append(json.dumps([float(x) for x in np.asarray(env.opt_ang_save).reshape(-1)]))
        rec["state"].append(env.state.numpy().copy())
        rec["obs_sum"].append(float(obs.double().sum()))
        rec["done_threshold"].append(float(env.done_threshold))
        if done:
            break
    out = {
        "conf": json.dumps(conf), "first_energy": rec_first, "obs0": obs0, "obs_len": len(obs0),
        "action_size": env.action_size, "state_size": env.state_size,
        "num_layers_termination": env.num_layers_termination,
        "illegal": np.asarray(rec["illegal"]), "opt_ang": np.asarray(rec["opt_ang"]),
        "state": np.stack(rec["state"]).astype(np.float32),
    }
    for k in ("action", "reward", "done", "energy", "nfev", "error", "obs_sum", "done_threshold"):
        out[k] = np.asarray(rec[k])
    if hasattr(env, "TN_state"):
        out["tn_state"] = np.asarray(env.TN_state)
    return out, env


def artefacts(env):
    """What a test needs to rebuild the env's input files: the npz content and the init circuit's op list."""
    d = np.load(env_path(env, "mol_data", "", ".npz"))
    H = d["hamiltonian"]
    out = {"H_re": np.ascontiguousarray(H.real), "eigvals": d["eigvals"], "weights": d["weights"],
           "paulis": d["paulis"], "energy_shift": d["energy_shift"] if "energy_shift" in d.files else np.zeros(())}
    if np.abs(H.imag).max() != 0:
        out["H_im"] = np.ascontiguousarray(H.imag)
    ic = loaders.load_qpy_circuit(env_path(env, "init_state_circ", "init_", f"_TNbond{env.TN_bond}.qpy"))
    out["init_name"] = np.asarray([o[0] for o in ic.ops])
    out["init_q0"] = np.asarray([o[1][0] for o in ic.ops], dtype=np.int32)
    out["init_q1"] = np.asarray([o[1][1] if len(o[1]) > 1 else -1 for o in ic.ops], dtype=np.int32)
    out["init_theta"] = np.asarray([0.0 if o[2] is None else o[2] for o in ic.ops], dtype=np.float64)
    return out


def env_path(env, folder, prefix, suffix):
    if env.ham_type not in ("heisenberg", "tfim_j1_h0.05"):
        stem = f"{env.ham_type}_{env.num_qubits}q_geom_{env.geometry}_{env.ham_mapping}"
    else:
        stem = f"{env.ham_type}_{env.num_qubits}q"
    return os.path.join(REF, "dmrg-to-qc", folder, prefix + stem + suffix)


def main():
    out = {}
    problems = {}
    for key, spec in EPISODES.items():
        rec, env = run_episode(*spec)
        pkey = f"{env.ham_type}_{env.num_qubits}q"
        if pkey not in problems:
            problems[pkey] = artefacts(env)
        rec["problem"] = pkey
        for k, v in rec.items():
            out[f"{key}/{k}"] = v
        print(key, "steps", len(rec["action"]), "nfev", rec["nfev"].tolist(), "E", rec["energy"][-1], flush=True)
    for pkey, art in problems.items():
        for k, v in art.items():
            out[f"problem/{pkey}/{k}"] = v
    out["episodes"] = np.asarray(list(EPISODES))
    np.savez_compressed(os.path.join(HERE, "env_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "env_golden.npz"), os.path.getsize(os.path.join(HERE, "env_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
