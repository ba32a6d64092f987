"""Drop-in for environments/VQAs/VQE_qulacs_TN_notin_RL_noise.py: the MPS state is loaded noise-free, the agent's
gates carry the hard-coded depolarising noise (reference :26-28,40-50,94-101)."""
import numpy as np

from . import _backend
from ._backend import CompiledAnsatz

P_ONE_QUBIT = 0.01   # VQE_qulacs_TN_notin_RL_noise.py:41
P_TWO_QUBIT = 0.05   # VQE_qulacs_TN_notin_RL_noise.py:27
NOISE_MODE = "trajectory"
rng = np.random.default_rng()


def seed(s):
    global rng
    rng = np.random.default_rng(s)


def shot_noise_np(weights, sigma):
    return np.real(weights.T @ np.random.normal(0, sigma, len(weights)))


class Parametric_Circuit:
    def __init__(self, n_qubits, noise_models=[], noise_values=[]):
        self.n_qubits = n_qubits
        self.ansatz = CompiledAnsatz(n_qubits)

    def construct_ansatz(self, state):
        return self.ansatz.load_tensor(state, noise=(P_ONE_QUBIT, P_TWO_QUBIT))


def get_energy_qulacs(angles, observable, circuit, weights, n_qubits, TN_state, n_shots, phys_noise=False,
                      which_angles=[]):
    _backend.apply_angles(circuit, angles, which_angles)
    return get_exp_val(n_qubits, circuit, observable, TN_state, n_shots, weights)


def get_exp_val(n_qubits, circuit, op, TN_state, n_shots, weights):
    sim = _backend.bind(n_qubits, circuit, op, TN_state, use_tn=True)
    p = circuit.params.reshape(1, -1)
    if NOISE_MODE == "density_matrix":
        return sim.energies_dm(p)[0]
    codes = _backend.sample_noise_codes(circuit.gates, _backend.noise_rng(rng), 1)
    return _backend.evaluate(sim, p, codes)


if __name__ == "__main__":
    pass
