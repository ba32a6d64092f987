"""Largest single-handle register (-m gpu): 30 qubits = 16 GiB of state on one B200.  Size-independent known answers of the
Heisenberg chain: |0...0> gives (n-1) + n, the Neel state -(n-1); a shallow layer of rotations and
a CNOT staircase gives the product-state value computed in closed form on the host."""
import numpy as np
import pytest

from tensorrl_qas_b200 import Simulator, loaders
from tensorrl_qas_b200.circuit import GateList

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [27, 30])
def test_large_register_known_answers(built_lib, n):
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    sim = Simulator(n, 0)
    sim.set_pauli_hamiltonian(x, z, w)
    sim.set_circuit(GateList(n))
    assert abs(sim.energies(np.zeros((1, 1)))[0] - (2 * n - 1)) < 1e-10
    neel = GateList(n)
    for q in range(1, n, 2):
        neel.add_pauli("X", q)
    sim.set_circuit(neel)
    assert abs(sim.energies(np.zeros((1, 1)))[0] + (n - 1) - (n % 2)) < 1e-10     # sum Z = +1 for odd n
    # product state RY(t_q)|0> on every qubit: <Z_q> = cos t_q, <X_q> = -sin t_q (RY = exp(+i t/2 Y)), <Y_q> = 0
    rng = np.random.default_rng(n)
    t = rng.uniform(-np.pi, np.pi, n)
    prod = GateList(n)
    for q in range(n):
        prod.add_rotation(1, q, float(t[q]))
    sim.set_circuit(prod)
    zq, xq = np.cos(t), -np.sin(t)
    want = float(np.sum(zq) + np.sum(zq[:-1] * zq[1:]) + np.sum(xq[:-1] * xq[1:]))
    got = sim.energies(t[None, :])[0]
    sim.close()
    assert abs(got - want) < 1e-10
