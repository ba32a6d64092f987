"""Drop-in for environments/VQAs/VQE_qulacs_noise.py: DepolarizingNoise(q, 0.01) after every rotation and
TwoQubitDepolarizingNoise(ctrl, targ, 0.05) after every CNOT (hard-coded there, :31-33,44-54), applied to a PURE
state -- i.e. one sampled Pauli trajectory per evaluation.

NOISE_MODE = "trajectory" reproduces that (codes sampled on the host from `rng`, seed with `seed()`; qulacs' own
RNG is unseeded in the reference); NOISE_MODE = "density_matrix" returns the exact channel average instead."""
import numpy as np

from . import _backend
from ._backend import CompiledAnsatz

P_ONE_QUBIT = 0.01   # VQE_qulacs_noise.py:45
P_TWO_QUBIT = 0.05   # VQE_qulacs_noise.py:32
NOISE_MODE = "trajectory"
rng = np.random.default_rng()


def seed(s):
    global rng
    rng = np.random.default_rng(s)


def shot_noise_np(weights, sigma):
    """reference: VQE_qulacs_noise.py:61-62 (only reachable from commented-out code there)"""
    return np.real(weights.T @ np.random.normal(0, sigma, len(weights)))


class Parametric_Circuit:
    """reference: VQE_qulacs_noise.py:11-58"""

    def __init__(self, n_qubits, noise_models=[], noise_values=[]):
        self.n_qubits = n_qubits
        self.ansatz = CompiledAnsatz(n_qubits)

    def construct_ansatz(self, state):
        return self.ansatz.load_tensor(state, noise=(P_ONE_QUBIT, P_TWO_QUBIT))


def get_energy_qulacs(angles, observable, circuit, weights, n_qubits, n_shots, phys_noise=False, which_angles=[]):
    """reference: VQE_qulacs_noise.py:64-95"""
    _backend.apply_angles(circuit, angles, which_angles)
    return get_exp_val(n_qubits, circuit, observable, n_shots, weights)


def get_exp_val(n_qubits, circuit, op, n_shots, weights):
    """reference: VQE_qulacs_noise.py:97-108 (shot noise is commented out there)"""
    sim = _backend.bind(n_qubits, circuit, op)
    p = circuit.params.reshape(1, -1)
    if NOISE_MODE == "density_matrix":
        return sim.energies_dm(p)[0]
    codes = _backend.sample_noise_codes(circuit.gates, _backend.noise_rng(rng), 1)
    return _backend.evaluate(sim, p, codes)


def get_energy_qulacs_batch(angles, observable, circuit, n_qubits, codes=None):
    """[B][P] angle sets -> [B] energies; one independent trajectory per element (or the given codes)."""
    sim = _backend.bind(n_qubits, circuit, observable)
    a = np.asarray(angles, dtype=np.float64)
    if NOISE_MODE == "density_matrix":
        return sim.energies_dm(a)
    if codes is None:
        codes = _backend.sample_noise_codes(circuit.gates, rng, a.shape[0])
    return sim.energies_traj(a, codes)


if __name__ == "__main__":
    pass
