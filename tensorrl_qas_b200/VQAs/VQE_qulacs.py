"""Drop-in for the reference's environments/VQAs/VQE_qulacs.py (noiseless, TN circuit encoded in the agent state).

Same names, argument order and return types as the reference module; the qulacs circuit object, the
`update_quantum_state` sweep and the dense numpy expectation (VQE_qulacs.py:79-86) are replaced by one fused
libtqsim launch.  `get_energy_qulacs_batch` is the batched extension used by the lock-step drivers and bench.py.
"""
import numpy as np

from . import _backend
from ._backend import CompiledAnsatz


class Parametric_Circuit:
    """reference: VQE_qulacs.py:6-44 -- noise_models / noise_values are accepted and ignored there as well."""

    def __init__(self, n_qubits, noise_models=[], noise_values=[]):
        self.n_qubits = n_qubits
        self.ansatz = CompiledAnsatz(n_qubits)

    def construct_ansatz(self, state):
        return self.ansatz.load_tensor(state)


def get_energy_qulacs(angles, observable, circuit, n_qubits, n_shots, phys_noise=False, which_angles=[]):
    """reference: VQE_qulacs.py:47-77"""
    _backend.apply_angles(circuit, angles, which_angles)
    return get_exp_val(n_qubits, circuit, observable)


def get_exp_val(n_qubits, circuit, op):
    """reference: VQE_qulacs.py:79-86 -- returns numpy float64 like `(np.conj(psi).T @ op @ psi).real`."""
    sim = _backend.bind(n_qubits, circuit, op)
    return _backend.evaluate(sim, circuit.params.reshape(1, -1))


def get_energy_qulacs_batch(angles, observable, circuit, n_qubits):
    """[B][P] angle sets -> [B] energies in one call (all columns are parameters of `circuit`, in order)."""
    sim = _backend.bind(n_qubits, circuit, observable)
    return sim.energies(np.asarray(angles, dtype=np.float64))


if __name__ == "__main__":
    pass
