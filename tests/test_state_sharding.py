"""Single-state sharding: schedule invariants, and the whole driver on CPU with the numpy
stand-in engine -- R virtual ranks in one process, and two real gloo ranks -- against the oracle on the full state."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import np_oracle
from sharded_np_engine import NumpyEngine
from tensorrl_qas_b200 import loaders
from tensorrl_qas_b200.circuit import KIND, GateList, brickwork_circuit, synthetic_circuit
from tensorrl_qas_b200.sharded import LocalComm, ShardedSimulator, TorchComm, plan_state_sharding, rank_terms


def _random_pauli_sum(n, n_terms, seed, max_weight=3):
    rng = np.random.default_rng(seed)
    xs, zs, cs = [], [], []
    for _ in range(n_terms):
        qs = rng.choice(n, size=int(rng.integers(1, max_weight + 1)), replace=False)
        x = z = 0
        for q in qs:
            p = int(rng.integers(3))   # X, Y, Z
            if p in (0, 1):
                x |= 1 << int(q)
            if p in (1, 2):
                z |= 1 << int(q)
        xs.append(x)
        zs.append(z)
        cs.append(float(rng.normal()))
    xs.append(0), zs.append(0), cs.append(0.75)   # an identity term: the shards are not normalised one by one
    return xs, zs, cs


def _oracle_energy(gl, params, pauli):
    psi = np_oracle.run_circuit(gl.n_qubits, gl.tuples(), params)
    return np_oracle.expect_pauli(psi, *pauli)


@pytest.mark.parametrize("n,g,seed", [(7, 1, 0), (8, 2, 1), (9, 3, 2), (10, 3, 3), (8, 0, 4)])
def test_schedule_invariants(n, g, seed):
    gl = synthetic_circuit(n, 60, seed)
    xs, _, _ = _random_pauli_sum(n, 12, seed)
    masks = sorted(set(xs))
    steps = plan_state_sharding(gl.tuples(), n, g, masks)
    nl = n - g
    seen, n_orig = set(), 0
    for st in steps:
        if st[0] == "evolve":
            for kind, p0, p1, pidx, fixed in st[1]:
                if kind == KIND["CNOT"]:
                    assert 0 <= p1 < nl and 0 <= p0 < n and p0 != p1   # the target is always local
                else:
                    assert 0 <= p0 < nl
                n_orig += pidx >= 0
        elif st[0] == "expect":
            pos = st[2]
            assert sorted(pos) == list(range(n))
            for gi in st[1]:
                assert gi not in seen
                seen.add(gi)
                assert all(pos[q] < nl for q in range(n) if (masks[gi] >> q) & 1)
    assert seen == set(range(len(masks)))
    assert n_orig == gl.n_params
    if g == 0:
        assert not any(st[0] == "exchange" for st in steps)


def test_rank_terms_fold_rank_bit_signs():
    # Z on a rank bit: a sign per rank; X/Y and local Z bits are relabelled
    pos = [2, 0, 3, 1]          # logical -> physical, n_local = 3: logical qubit 2 is the rank bit
    x, z, c = rank_terms([(0b0001, 0b0101, 2.0)], pos, 3, rank=1)
    assert (int(x[0]), int(z[0]), float(c[0])) == (0b100, 0b100, -2.0)
    x, z, c = rank_terms([(0b0001, 0b0101, 2.0)], pos, 3, rank=0)
    assert float(c[0]) == 2.0


@pytest.mark.parametrize("n,ranks,seed", [(7, 2, 10), (8, 4, 11), (9, 8, 12), (10, 4, 13)])
def test_virtual_ranks_match_the_oracle(n, ranks, seed):
    gl = synthetic_circuit(n, 70, seed)
    pauli = _random_pauli_sum(n, 15, seed)
    rng = np.random.default_rng(seed)
    params = np.asarray(gl.initial_angles) + rng.uniform(-0.1, 0.1, gl.n_params)
    sim = ShardedSimulator(n, LocalComm(ranks), engine=NumpyEngine(n - ranks.bit_length() + 1))
    sim.set_circuit(gl)
    sim.set_pauli_hamiltonian(*pauli)
    e = sim.energy(params)
    assert abs(e - _oracle_energy(gl, params, pauli)) < 1e-12
    assert sim.n_exchanges >= 1
    # a second evaluation reuses the compiled schedule
    params2 = params + 0.05
    assert abs(sim.energy(params2) - _oracle_energy(gl, params2, pauli)) < 1e-12


def test_heisenberg_brickwork_known_answers():
    n = 10
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    sim = ShardedSimulator(n, LocalComm(4), engine=NumpyEngine(n - 2))
    sim.set_pauli_hamiltonian(x, z, w)
    sim.set_circuit(GateList(n))                      # no gates: |0...0>, E = (n-1) + n
    assert abs(sim.energy() - (2 * n - 1)) < 1e-12
    neel = GateList(n)
    for q in range(1, n, 2):
        neel.add_pauli("X", q)
    sim.set_circuit(neel)                             # Neel state: ZZ = -(n-1), sum Z = 0
    assert abs(sim.energy() + (n - 1)) < 1e-12
    gl = brickwork_circuit(n, 21, 20, 5)
    sim.set_circuit(gl)
    p = np.asarray(gl.initial_angles)
    assert abs(sim.energy(p) - _oracle_energy(gl, p, (x, z, w))) < 1e-12


def test_errors():
    with pytest.raises(ValueError):
        ShardedSimulator(8, LocalComm(3), engine=NumpyEngine(6))
    with pytest.raises(ValueError):
        plan_state_sharding([], 5, 3, [])                                     # too few local qubits
    with pytest.raises(ValueError):
        plan_state_sharding([(KIND["DEPOL1"], 0, 0, 0, 0.1)], 8, 1, [])       # pure-state kinds only
    with pytest.raises(ValueError):
        plan_state_sharding([], 6, 2, [0b111111])                             # flips more qubits than a shard holds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 9
    gl = synthetic_circuit(n, 80, 21)
    pauli = _random_pauli_sum(n, 15, 21)
    params = np.asarray(gl.initial_angles)
    sim = ShardedSimulator(n, TorchComm(), engine=NumpyEngine(n - 1))
    sim.set_circuit(gl)
    sim.set_pauli_hamiltonian(*pauli)
    e = sim.energy(params)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.asarray([e, sim.n_exchanges, _oracle_energy(gl, params, pauli)]))
    dist.destroy_process_group()


def test_two_gloo_ranks_match_the_oracle(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        e, n_ex, ref = np.load(tmp_path / f"rank{r}.npy")
        assert n_ex >= 1
        assert abs(e - ref) < 1e-12


def test_schedule_quality_on_the_benchmark_shape():
    """The C5-shaped brickwork circuit + Heisenberg chain at the sizes measured in round 1 (profiles/state_sharding_r01.json):
    two qubit exchanges per evaluation, whatever the rank count (pure planning, no state is touched)."""
    for n, g in ((24, 2), (28, 1), (28, 3), (31, 3)):
        gl = brickwork_circuit(n, 21, 41, 5)
        paulis, _ = loaders.heisenberg_terms(n)
        x, _ = loaders.pauli_masks(paulis, n)
        steps = plan_state_sharding(gl.tuples(), n, g, sorted(set(int(v) for v in x)))
        assert sum(1 for s in steps if s[0] == "exchange") <= 2, (n, g)
        # every original gate appears exactly once; the only additions are the SWAP triples in front of an exchange
        n_ops = sum(len(s[1]) for s in steps if s[0] == "evolve")
        assert n_ops >= len(gl) and (n_ops - len(gl)) % 3 == 0 and n_ops - len(gl) <= 3 * g * 2
