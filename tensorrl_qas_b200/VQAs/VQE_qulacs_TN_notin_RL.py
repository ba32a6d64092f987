"""Drop-in for environments/VQAs/VQE_qulacs_TN_notin_RL.py: agent gates only, applied to the loaded MPS state
(`state.load(TN_state)`, reference :82-83)."""
import numpy as np

from . import _backend
from ._backend import CompiledAnsatz


class Parametric_Circuit:
    """reference: VQE_qulacs_TN_notin_RL.py:6-45"""

    def __init__(self, n_qubits, noise_models=[], noise_values=[]):
        self.n_qubits = n_qubits
        self.ansatz = CompiledAnsatz(n_qubits)

    def construct_ansatz(self, state):
        return self.ansatz.load_tensor(state)


def get_energy_qulacs(angles, observable, circuit, n_qubits, TN_state, n_shots, phys_noise=False, which_angles=[]):
    """reference: VQE_qulacs_TN_notin_RL.py:48-78"""
    _backend.apply_angles(circuit, angles, which_angles)
    return get_exp_val(n_qubits, circuit, observable, TN_state)


def get_exp_val(n_qubits, circuit, op, TN_state):
    """reference: VQE_qulacs_TN_notin_RL.py:80-87"""
    sim = _backend.bind(n_qubits, circuit, op, TN_state, use_tn=True)
    return _backend.evaluate(sim, circuit.params.reshape(1, -1))


def get_energy_qulacs_batch(angles, observable, circuit, n_qubits, TN_state):
    sim = _backend.bind(n_qubits, circuit, observable, TN_state, use_tn=True)
    return sim.energies(np.asarray(angles, dtype=np.float64))


if __name__ == "__main__":
    pass
