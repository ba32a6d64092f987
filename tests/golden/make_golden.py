"""Generates tests/golden/vqa_golden.npz by running the REFERENCE's own hot-path code
(/root/reference/environments/VQAs/VQE_qulacs*.py: construct_ansatz, get_energy_qulacs, get_exp_val) on the
reference's shipped artefacts (dmrg-to-qc/mol_data/*.npz, dmrg-to-qc/init_state_circ/*.qpy), with the qulacs
primitives supplied by oracle/np_oracle.py (qulacs itself cannot be imported here -- "parity unpinned" for the
gate arithmetic, pinned for everything the reference's Python does: tensor decoding, gate order, parameter mapping,
initial-state loading, dense expectation, noise-gate placement).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The GPU box never runs this; tests read the committed npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import np_oracle  # noqa: E402

np_oracle.install_as_qulacs()
from environments.VQAs import VQE_qulacs as vc_in  # noqa: E402  (TN in agent, from |0..0>)
from environments.VQAs import VQE_qulacs_TN_notin_RL as vc_notin  # noqa: E402  (state.load(TN_state))
from environments.VQAs import VQE_qulacs_noise as vc_noise  # noqa: E402  (probabilistic depolarising gates)

from tensorrl_qas_b200 import loaders  # noqa: E402

CASES = {
    "beh2_6q": ("BEH2_6q_geom_H_0.000_0.000_-1.330;_Be_0.000_0.000_0.000;_H_0.000_0.000_1.330_jordan_wigner", 6),
    "h2o_8q": ("H2O_8q_geom_H_-0.021_-0.002_0.000;_O_0.835_0.452_0.000;_H_1.477_-0.273_0.000_jordan_wigner", 8),
    "ch2_8q": ("CH2_8q_geom_C_0.000_0.000_0.000;_H_1.080_0.000_0.000;_H_-0.225_1.056_0.000_jordan_wigner", 8),
    "heis_5q": ("heisenberg_5q", 5),
}
AXIS = {"rx": 0, "ry": 1, "rz": 2}


def encode_init(circ, n, num_layers, zero_param_init=False):
    """The tensor the reference's reset() builds from the QPY circuit (environments/environment_qulacs.py:281-328):
    qiskit qubit p -> column n-1-p, angle -> -theta as float32, CNOT -> [n-1-targ][n-1-ctrl] = 1."""
    state = torch.zeros((num_layers, n + 6, n))
    for depth_no, layer in enumerate(circ.layers()):
        for name, qs, angle in layer:
            if name != "cx":
                col = n - 1 - qs[0]
                state[depth_no][n + AXIS[name]][col] = 1
                state[depth_no][n + 3 + AXIS[name]][col] = 0 if zero_param_init else -angle
            else:
                state[depth_no][n - 1 - qs[1]][n - 1 - qs[0]] = 1
    return state


def add_agent_gates(state, n, first_layer, n_gates, rng):
    """Random agent actions placed like step() does (environments/environment_qulacs.py:178-216): one gate per
    action at layer first_layer + moment; new rotations enter with a random angle so the parity test is not trivial."""
    moments = [0] * n
    for _ in range(n_gates):
        if rng.random() < 0.6:
            q, axis = int(rng.integers(n)), int(rng.integers(3))
            layer = first_layer + moments[q]
            if state[layer][n + axis][q] == 1:
                continue
            state[layer][n + axis][q] = 1
            state[layer][n + 3 + axis][q] = float(rng.uniform(-np.pi, np.pi))
            moments[q] += 1
        else:
            c = int(rng.integers(n))
            t = (c + int(rng.integers(1, n))) % n
            m = max(moments[c], moments[t])
            state[first_layer + m][t][c] = 1
            moments[c] = moments[t] = m + 1
    return state


def main():
    out = {}
    rng = np.random.default_rng(20261018)
    for key, (stem, n) in CASES.items():
        ham = np.load(f"{REF}/dmrg-to-qc/mol_data/{stem}.npz", allow_pickle=True)
        H = ham["hamiltonian"]
        circ = loaders.load_qpy_circuit(f"{REF}/dmrg-to-qc/init_state_circ/init_{stem}_TNbond2.qpy")
        depth = circ.depth()
        out[f"{key}/n"] = n
        out[f"{key}/paulis"] = np.asarray([str(s) for s in ham["paulis"]])
        out[f"{key}/weights"] = np.asarray(ham["weights"], dtype=np.float64)
        out[f"{key}/eig_min"] = float(min(ham["eigvals"]))
        out[f"{key}/h00"] = float(H[0, 0].real)
        out[f"{key}/init_name"] = np.asarray([o[0] for o in circ.ops])
        out[f"{key}/init_q0"] = np.asarray([o[1][0] for o in circ.ops], dtype=np.int32)
        out[f"{key}/init_q1"] = np.asarray([o[1][1] if len(o[1]) > 1 else -1 for o in circ.ops], dtype=np.int32)
        out[f"{key}/init_theta"] = np.asarray([o[2] if o[2] is not None else 0.0 for o in circ.ops], dtype=np.float64)
        out[f"{key}/init_depth"] = depth

        # ---- TN-in-agent ("trainable"): whole circuit from |0..0>, un-reversed H --------------------------------
        L = depth + 12
        tensor = add_agent_gates(encode_init(circ, n, L), n, depth, 14, rng)
        ansatz = vc_in.Parametric_Circuit(n_qubits=n).construct_ansatz(tensor)
        P = ansatz.get_parameter_count()
        e_tensor = vc_in.get_exp_val(n, ansatz, H)  # what CircuitEnv.get_energy returns for this tensor
        X = np.empty((6, P))
        E = np.empty(6)
        x0 = np.asarray([ansatz.get_parameter(j) for j in range(P)])
        for r in range(6):
            X[r] = x0 + rng.uniform(-0.3, 0.3, size=P)
            E[r] = vc_in.get_energy_qulacs(X[r], observable=H, circuit=ansatz, n_qubits=n, n_shots=0,
                                           phys_noise=False, which_angles=[])
        out[f"{key}/in_tensor"] = tensor.numpy()
        out[f"{key}/in_e_tensor"] = e_tensor
        out[f"{key}/in_X"] = X
        out[f"{key}/in_E"] = E
        e_first = vc_in.get_exp_val(n, vc_in.Parametric_Circuit(n).construct_ansatz(encode_init(circ, n, L)), H)
        out[f"{key}/in_e_first"] = e_first  # "Very first energy" of the trainable env (float32-rounded angles)
        e_zero = vc_in.get_exp_val(n, vc_in.Parametric_Circuit(n).construct_ansatz(encode_init(circ, n, L, True)), H)
        out[f"{key}/in_e_zero_param"] = e_zero  # StructureRL: zero_param_init=1

        # ---- TN-not-in-agent ("fixed"): agent gates only, state.load(TN_state), bit-reversed H ------------------
        tn_state = np_oracle.run_circuit(n, loaders.init_circuit_gatelist(circ).tuples(), np.zeros(1))
        Hrev = loaders.reverse_qargs(H)
        tensor2 = add_agent_gates(torch.zeros((20, n + 6, n)), n, 0, 16, rng)
        ansatz2 = vc_notin.Parametric_Circuit(n_qubits=n).construct_ansatz(tensor2)
        P2 = ansatz2.get_parameter_count()
        X2 = rng.uniform(-np.pi, np.pi, size=(6, P2))
        E2 = np.asarray([vc_notin.get_energy_qulacs(X2[r], observable=Hrev, circuit=ansatz2, n_qubits=n,
                                                    TN_state=tn_state, n_shots=0, phys_noise=False, which_angles=[])
                         for r in range(6)])
        out[f"{key}/notin_tn_state"] = tn_state
        out[f"{key}/notin_tensor"] = tensor2.numpy()
        out[f"{key}/notin_X"] = X2
        out[f"{key}/notin_E"] = E2
        out[f"{key}/notin_e_first"] = vc_notin.get_exp_val(
            n, vc_notin.Parametric_Circuit(n).construct_ansatz(torch.zeros((20, n + 6, n))), Hrev, tn_state)

        # ---- noise, TN-in-agent: one sampled Pauli trajectory per evaluation ------------------------------------
        if key in ("beh2_6q", "h2o_8q"):
            np_oracle.seed(7)
            ansatz3 = vc_noise.Parametric_Circuit(n_qubits=n).construct_ansatz(tensor)
            R = 8
            E3 = np.empty(R)
            codes = []
            for r in range(R):
                del np_oracle.noise_log[:]
                E3[r] = vc_noise.get_energy_qulacs(X[r % 6], observable=H, circuit=ansatz3, weights=ham["weights"],
                                                   n_qubits=n, n_shots=0, phys_noise=True, which_angles=[])
                codes.append(list(np_oracle.noise_log))
            out[f"{key}/noise_codes"] = np.asarray(codes, dtype=np.uint8)
            out[f"{key}/noise_E"] = E3
        print(key, "P =", P, "first E", e_first, "eig_min", out[f"{key}/eig_min"])
    np.savez_compressed(os.path.join(HERE, "vqa_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "vqa_golden.npz"), os.path.getsize(os.path.join(HERE, "vqa_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
