"""Seeded fuzz of the GPU paths against the oracle (-m gpu): SHORT circuits (1-50 gates of every kind on 9-16 qubits: the
regime where single gates stay un-fused and land on every kind of tile position), all tile sizes / forced low bits, |0...0>
and loaded states, random Pauli sums; and noisy circuits on the trajectory path (tensor-core and FP64-pipe tiles) and the
exact density-matrix path.  Energies 1e-10 Ha, states 1e-12.  (A 1500 + 600 case run of the same generators was clean in
round 1; the lone-diagonal-gate bug of the tensor-core planner was found by circuits of this shape.)"""
import numpy as np
import pytest

from tensorrl_qas_b200 import Simulator
from tensorrl_qas_b200.circuit import GateList

pytestmark = pytest.mark.gpu


def _random_pauli_sum(rng, n, max_terms, max_flips):
    T = int(rng.integers(1, max_terms))
    x = np.zeros(T, dtype=np.uint64)
    for t in range(T):
        for q in rng.choice(n, size=int(rng.integers(0, min(max_flips, n) + 1)), replace=False):
            x[t] |= np.uint64(1 << int(q))
    z = rng.integers(0, 1 << n, size=T).astype(np.uint64)
    return x, z, rng.normal(size=T)


def _random_state(rng, n):
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


@pytest.mark.parametrize("chunk", range(8))
def test_short_circuits(built_lib, oracle, monkeypatch, chunk):
    for it in range(40):
        seed = 40 * chunk + it
        rng = np.random.default_rng(seed)
        n = int(rng.integers(9, 17))
        env = {"TQ_TILE_BITS": str(int(rng.choice([9, 10, 11, 12]))), "TQ_LOW_BITS": str(int(rng.choice([2, 3, 4])))}
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        gl = GateList(n)
        style = int(rng.integers(4))
        for _ in range(int(rng.choice([1, 2, 3, 4, 6, 10, 20, 50]))):
            u = rng.random()
            q = int(rng.integers(n))
            if style == 0 and u < 0.5:        # diagonal-heavy
                gl.add_rotation(2, q, float(rng.uniform(-3, 3)), parametric=bool(rng.random() < 0.7))
            elif style == 1 and u < 0.5:      # CNOT-heavy
                c, t = (int(v) for v in rng.choice(n, size=2, replace=False))
                gl.add_cnot(c, t)
            elif u < 0.6:
                gl.add_rotation(int(rng.integers(3)), q, float(rng.uniform(-3, 3)), parametric=bool(rng.random() < 0.7))
            elif u < 0.7:
                gl.add_pauli("XYZ"[int(rng.integers(3))], q)
            else:
                c, t = (int(v) for v in rng.choice(n, size=2, replace=False))
                gl.add_cnot(c, t)
        ham = _random_pauli_sum(rng, n, 25, 3)
        init = _random_state(rng, n) if rng.random() < 0.5 else None
        p = np.asarray(gl.initial_angles, dtype=np.float64)[None, :] if gl.n_params else np.zeros((1, 1))
        p = np.repeat(p, 2, axis=0)
        p[1] += 0.05
        sim = Simulator(n)
        sim.set_circuit(gl)
        sim.set_pauli_hamiltonian(*ham)
        if init is not None:
            sim.set_init_state(init)
        e = sim.energies(p)
        st = sim.states(p[:1])[0]
        sim.close()
        what = (seed, n, env, len(gl), init is not None)
        assert np.abs(e - oracle.energies(gl, p, pauli=ham, init=init)).max() < 1e-10, what
        assert np.abs(st - oracle.state(gl, p[0], init=init)).max() < 1e-12, what


@pytest.mark.parametrize("chunk", range(4))
def test_noisy_circuits(built_lib, oracle, monkeypatch, chunk):
    from tensorrl_qas_b200.VQAs._backend import sample_noise_codes
    for it in range(50):
        seed = 50000 + 50 * chunk + it
        rng = np.random.default_rng(seed)
        dm = rng.random() < 0.35
        n = int(rng.integers(2, 7)) if dm else int(rng.integers(4, 15))
        monkeypatch.setenv("TQ_TILE_BITS", str(int(rng.choice([9, 10, 12]))))
        gl = GateList(n)
        for _ in range(int(rng.choice([1, 3, 8, 25, 60]))):
            q = int(rng.integers(n))
            if rng.random() < 0.55:
                gl.add_rotation(int(rng.integers(3)), q, float(rng.uniform(-3, 3)))
                if rng.random() < 0.8:
                    gl.add_depol1(q, float(rng.choice([0.01, 0.1, 0.3])))
            else:
                c, t = (int(v) for v in rng.choice(n, size=2, replace=False))
                gl.add_cnot(c, t)
                if rng.random() < 0.8:
                    gl.add_depol2(c, t, float(rng.choice([0.05, 0.2])))
        ham = _random_pauli_sum(rng, n, 15, 3)
        init = _random_state(rng, n) if rng.random() < 0.4 else None
        p = np.asarray(gl.initial_angles, dtype=np.float64)[None, :] if gl.n_params else np.zeros((1, 1))
        p = np.repeat(p, 3, axis=0)
        p[1] += 0.05
        p[2] -= 0.07
        sim = Simulator(n)
        sim.set_circuit(gl)
        sim.set_pauli_hamiltonian(*ham)
        if init is not None:
            sim.set_init_state(init)
        if dm:
            e, ref = sim.energies_dm(p), oracle.dm_energies(gl, p, pauli=ham, init=init)
        else:
            codes = sample_noise_codes(gl, rng, 3)
            e, ref = sim.energies_traj(p, codes), oracle.energies(gl, p, pauli=ham, init=init, codes=codes)
        sim.close()
        assert np.abs(e - ref).max() < 1e-10, (seed, dm, n, len(gl))
