"""Drop-in for environments/VQAs/VQE_qulacs_TN_notin_RL_noise_restricted.py.  Despite its name the reference module
applies NO gate noise; it only adds Gaussian shot noise Re(w^T N(0, n_shots^-1/2)) when n_shots != 0
(reference :47-48,84-96), drawn from numpy's global RNG -- kept on numpy's global RNG here so seeded runs agree."""
import numpy as np

from . import _backend
from ._backend import CompiledAnsatz


def shot_noise_np(weights, sigma):
    """reference: :47-48"""
    return np.real(weights.T @ np.random.normal(0, sigma, len(weights)))


class Parametric_Circuit:
    def __init__(self, n_qubits, noise_models=[], noise_values=[]):
        self.n_qubits = n_qubits
        self.ansatz = CompiledAnsatz(n_qubits)

    def construct_ansatz(self, state):
        return self.ansatz.load_tensor(state)


def get_energy_qulacs(angles, observable, circuit, weights, n_qubits, TN_state, n_shots, phys_noise=False,
                      which_angles=[]):
    _backend.apply_angles(circuit, angles, which_angles)
    return get_exp_val(n_qubits, circuit, observable, TN_state, n_shots, weights)


def get_exp_val(n_qubits, circuit, op, TN_state, n_shots, weights):
    sim = _backend.bind(n_qubits, circuit, op, TN_state, use_tn=True)
    expval = _backend.evaluate(sim, circuit.params.reshape(1, -1))
    if n_shots != 0:
        sigma = (n_shots) ** (-0.5)
        shot_noise = shot_noise_np(weights, sigma)
    else:
        shot_noise = 0
    return expval + shot_noise


if __name__ == "__main__":
    pass
