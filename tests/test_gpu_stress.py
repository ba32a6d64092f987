"""Seeded random sweep over the planner / kernel configuration space of the tensor-core path (-m gpu): tile sizes, forced
low bits, circuit shapes (staircases, stars, random), Pauli / CNOT-heavy circuits, loaded initial states, random Pauli
Hamiltonians with wide flip masks and long Z strings.  Every case is checked against the oracle (1e-10 Ha, states 1e-12)."""
import numpy as np
import pytest

from tensorrl_qas_b200 import Simulator
from tensorrl_qas_b200.circuit import GateList, append_random_gates, parameter_batch

pytestmark = pytest.mark.gpu
TOL = 1e-10


def random_problem(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(9, 16))
    gl = GateList(n)
    shape = int(rng.integers(4))
    if shape == 0:      # staircase of bricks + scattered gates
        for q in range(n - 1):
            append_random_gates(gl, int(rng.integers(3, 12)), rng, (q, q + 1))
        append_random_gates(gl, int(rng.integers(0, 40)), rng)
    elif shape == 1:    # star around one qubit
        hub = int(rng.integers(n))
        for q in rng.permutation(n):
            if q != hub:
                append_random_gates(gl, int(rng.integers(2, 8)), rng, (hub, int(q)))
    elif shape == 2:    # fully random, CNOT-heavy, with fixed Paulis
        for _ in range(int(rng.integers(20, 160))):
            u = rng.random()
            if u < 0.45:
                c, t = rng.choice(n, size=2, replace=False)
                gl.add_cnot(int(c), int(t))
            elif u < 0.55:
                gl.add_pauli("XYZ"[int(rng.integers(3))], int(rng.integers(n)))
            else:
                gl.add_rotation(int(rng.integers(3)), int(rng.integers(n)), float(rng.uniform(-np.pi, np.pi)),
                                parametric=bool(rng.random() < 0.8))
    else:               # gates on a few high qubits only (most of the register stays |0>), then a little everywhere
        hi = [int(q) for q in rng.choice(np.arange(n // 2, n), size=3, replace=False)]
        append_random_gates(gl, int(rng.integers(10, 40)), rng, hi)
        append_random_gates(gl, int(rng.integers(0, 12)), rng)
    T = int(rng.integers(5, 40))
    x = np.zeros(T, dtype=np.uint64)
    for t in range(T):
        for q in rng.choice(n, size=int(rng.integers(0, 5)), replace=False):
            x[t] |= np.uint64(1 << int(q))
    z = rng.integers(0, 1 << n, size=T).astype(np.uint64)
    w = rng.normal(size=T)
    init = None
    if rng.random() < 0.3:
        v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        init = v / np.linalg.norm(v)
    env = {"TQ_TILE_BITS": str(int(rng.choice([9, 10, 11, 12]))), "TQ_LOW_BITS": str(int(rng.choice([2, 3, 4])))}
    return n, gl, (x, z, w), init, env


@pytest.mark.parametrize("seed", range(40))
def test_random_problem_matches_oracle(built_lib, oracle, monkeypatch, seed):
    n, gl, ham, init, env = random_problem(1000 + seed)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    sim = Simulator(n)
    sim.set_circuit(gl)
    sim.set_pauli_hamiltonian(*ham)
    if init is not None:
        sim.set_init_state(init)
    p = parameter_batch(gl, 3)
    e = sim.energies(p)
    assert np.abs(e - oracle.energies(gl, p, pauli=ham, init=init)).max() < TOL, (n, env, len(gl))
    st = sim.states(p[:1])[0]
    assert np.abs(st - oracle.state(gl, p[0], init=init)).max() < 1e-12
    assert np.array_equal(sim.energies(p), e)   # deterministic
    sim.close()


@pytest.mark.parametrize("seed", range(16))
def test_sparse_circuits_with_lone_diagonal_gates(built_lib, oracle, monkeypatch, seed):
    """Few gates on many qubits: lone RZ / Z gates whose qubit nothing else in the window touches, CNOT triples (SWAPs),
    loaded states.  Regression for [CNOT(11,9), RX(11), RZ(7)] on 12 qubits (a phase applied per thread inside a DMMA whose
    B fragment all rows share)."""
    rng = np.random.default_rng(7000 + seed)
    n = 12 if seed == 0 else int(rng.integers(9, 16))
    gl = GateList(n)
    if seed == 0:
        gl.add_cnot(11, 9)
        gl.add_rotation(0, 11, 0.3)
        gl.add_rotation(2, 7, 0.7)
    else:
        for _ in range(int(rng.integers(3, 40))):
            u = rng.random()
            if u < 0.45:
                gl.add_rotation(2, int(rng.integers(n)), float(rng.uniform(-3, 3)))
            elif u < 0.55:
                gl.add_pauli("Z", int(rng.integers(n)))
            elif u < 0.75:
                gl.add_rotation(int(rng.integers(2)), int(rng.integers(n)), float(rng.uniform(-3, 3)))
            elif u < 0.9:
                c, t = rng.choice(n, size=2, replace=False)
                gl.add_cnot(int(c), int(t))
            else:
                a, b = (int(v) for v in rng.choice(n, size=2, replace=False))
                gl.add_cnot(a, b), gl.add_cnot(b, a), gl.add_cnot(a, b)
    monkeypatch.setenv("TQ_TILE_BITS", str(int(rng.choice([9, 10, 12]))))
    init = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    init /= np.linalg.norm(init)
    p = parameter_batch(gl, 2)
    sim = Simulator(n)
    sim.set_circuit(gl)
    for ini in (None, init):
        sim.set_init_state(ini)
        st = sim.states(p[:1])[0]
        assert np.abs(st - oracle.state(gl, p[0], init=ini)).max() < 1e-12
    sim.close()
