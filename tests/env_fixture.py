"""Test fixtures for the CircuitEnv drop-ins: rebuilds the input files the environments read (dmrg-to-qc/mol_data/*.npz,
dmrg-to-qc/init_state_circ/*.qpy) from tests/golden/env_golden.npz into a temporary data root, and provides
oracle-backed stand-ins for the VQE_qulacs* shims so the host logic can be replayed without a GPU.
TEST INFRASTRUCTURE: the only place where the oracle meets the environments."""
import json
import os
import struct
import types

import numpy as np

from oracle import np_oracle
from tensorrl_qas_b200 import loaders
from tensorrl_qas_b200.circuit import decode_state_tensor

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "env_golden.npz")
_GATE_CLASS = {"rx": "RXGate", "ry": "RYGate", "rz": "RZGate", "cx": "CXGate"}
_cache = {}


def golden():
    if "g" not in _cache:
        _cache["g"] = dict(np.load(GOLDEN, allow_pickle=False))
    return _cache["g"]


def episodes():
    return [str(k) for k in golden()["episodes"]]


class Episode:
    def __init__(self, key):
        g = golden()
        self.key = key
        self.d = {k.split("/", 1)[1]: v for k, v in g.items() if k.startswith(key + "/")}
        self.conf = json.loads(str(self.d["conf"]))
        self.problem = str(self.d["problem"])
        self.art = {k.split("/", 2)[2]: v for k, v in g.items() if k.startswith(f"problem/{self.problem}/")}

    @property
    def n_steps(self):
        return len(self.d["action"])

    def illegal(self, i):
        return json.loads(str(self.d["illegal"][i]))

    def opt_ang(self, i):
        return np.asarray(json.loads(str(self.d["opt_ang"][i])), dtype=np.float64)


def write_qpy(path, n_qubits, ops, name="circuit"):
    """QPY format version 14 file (layout of qiskit 2.0.0) holding one circuit of rx/ry/rz/cx gates."""
    out = [struct.pack("!6sBBBBQ", b"QISKIT", 14, 2, 0, 0, 1), b"p", b"q"]
    meta = b"{}"
    phase = struct.pack("<d", 0.0)
    nm = name.encode()
    out.append(struct.pack("!H1cHIIQIQI", len(nm), b"f", len(phase), n_qubits, 0, len(meta), 0, len(ops), 0))
    out += [nm, phase, meta, struct.pack("!Q", 0)]
    for gate, qubits, angle in ops:
        cls = _GATE_CLASS[gate].encode()
        n_par = 0 if gate == "cx" else 1
        out.append(struct.pack("!HHHII?HqII", len(cls), 0, n_par, len(qubits), 0, False, 0, 0, 0, 0))
        out.append(cls)
        for q in qubits:
            out.append(struct.pack("!1cI", b"q", q))
        if n_par:
            out.append(struct.pack("!1cQ", b"f", 8) + struct.pack("<d", float(angle)))
    with open(path, "wb") as f:
        f.write(b"".join(out))


def materialize(root, ep):
    """Writes the artefacts of `ep`'s problem under root/dmrg-to-qc/... with the reference's file names."""
    env, prob = ep.conf["env"], ep.conf["problem"]
    n = env["num_qubits"]
    if prob["ham_type"] not in ("heisenberg", "tfim_j1_h0.05"):
        stem = f"{prob['ham_type']}_{n}q_geom_{prob['geometry'].replace(' ', '_')}_{prob['mapping']}"
    else:
        stem = f"{prob['ham_type']}_{n}q"
    a = ep.art
    os.makedirs(os.path.join(root, "dmrg-to-qc", "mol_data"), exist_ok=True)
    os.makedirs(os.path.join(root, "dmrg-to-qc", "init_state_circ"), exist_ok=True)
    H = a["H_re"].astype(np.complex128)
    if "H_im" in a:
        H = H + 1j * a["H_im"]
    np.savez(os.path.join(root, "dmrg-to-qc", "mol_data", stem + ".npz"), hamiltonian=H, eigvals=a["eigvals"],
             weights=a["weights"], paulis=a["paulis"], energy_shift=a["energy_shift"])
    ops = []
    for name, q0, q1, th in zip(a["init_name"], a["init_q0"], a["init_q1"], a["init_theta"]):
        name = str(name)
        ops.append((name, (int(q0), int(q1)) if name == "cx" else (int(q0),), None if name == "cx" else float(th)))
    write_qpy(os.path.join(root, "dmrg-to-qc", "init_state_circ", f"init_{stem}_TNbond{env['tn_bond']}.qpy"), n, ops)
    return root


# ------------------------------------------------------------------------------ oracle-backed VQA stand-ins ----
def oracle_vc(tn_state_arg, noise, shot_args, shot_noise=False):
    """A module-like object with the reference's Parametric_Circuit / get_energy_qulacs / get_exp_val signatures of
    the given variant, computing with oracle/np_oracle.py (same gate order and numpy expression as the reference's
    environments/VQAs/VQE_qulacs*.py on top of np_oracle, which is how the golden episodes were produced)."""
    m = types.SimpleNamespace()

    class Parametric_Circuit:
        def __init__(self, n_qubits, noise_models=[], noise_values=[]):
            self.n_qubits = n_qubits
            self.ansatz = np_oracle.ParametricQuantumCircuit(n_qubits)

        def construct_ansatz(self, state):
            gl = decode_state_tensor(state, self.n_qubits)
            for kind, q0, q1, _pidx, fixed in gl.tuples():
                if kind == 3:
                    self.ansatz.add_gate(np_oracle.CNOT(q0, q1))
                    if noise:
                        self.ansatz.add_gate(np_oracle.TwoQubitDepolarizingNoise(q0, q1, 0.05))
                else:
                    (self.ansatz.add_parametric_RX_gate, self.ansatz.add_parametric_RY_gate,
                     self.ansatz.add_parametric_RZ_gate)[kind](q0, fixed)
                    if noise:
                        self.ansatz.add_gate(np_oracle.DepolarizingNoise(q0, 0.01))
            return self.ansatz

    def get_exp_val(n_qubits, circuit, op, *rest):
        rest = list(rest)
        state = np_oracle.QuantumState(n_qubits)
        if tn_state_arg:
            state.load(rest.pop(0))
        circuit.update_quantum_state(state)
        psi = state.get_vector()
        expval = (np.conj(psi).T @ op @ psi).real
        if shot_noise:
            n_shots, weights = rest
            if n_shots != 0:
                expval = expval + np.real(weights.T @ np.random.normal(0, n_shots ** (-0.5), len(weights)))
        return expval

    def get_energy_qulacs(angles, observable, circuit, n_qubits, n_shots, phys_noise=False, which_angles=[],
                          TN_state=None, weights=None):
        idx = list(which_angles) or range(circuit.get_parameter_count())
        for i, j in enumerate(idx):
            circuit.set_parameter(j, angles[i])
        args = ([TN_state] if tn_state_arg else []) + ([n_shots, weights] if shot_args else [])
        return get_exp_val(n_qubits, circuit, observable, *args)

    m.Parametric_Circuit, m.get_exp_val, m.get_energy_qulacs = Parametric_Circuit, get_exp_val, get_energy_qulacs
    m.seed = np_oracle.seed
    return m


def oracle_statevector(init_circuit):
    gl = loaders.init_circuit_gatelist(init_circuit, parametric=False)
    return np_oracle.run_circuit(init_circuit.n_qubits, gl.tuples(), np.zeros(0))
