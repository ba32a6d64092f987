"""The five BASELINE.json configurations as GPU parity cases at their stated sizes:
oracle comparison on a sample of the batch, size-independent properties on the whole batch (batch independence,
re-run determinism, known answers).  Tolerance 1e-10 Ha (BASELINE.json north_star)."""
import numpy as np
import pytest

from tensorrl_qas_b200 import Simulator, loaders
from tensorrl_qas_b200.circuit import GateList, append_random_gates, brickwork_circuit, parameter_batch, synthetic_circuit

from golden_util import Case

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _with_agent_gates(gl, n_gates, seed):
    append_random_gates(gl, n_gates, np.random.default_rng(seed))
    return gl


def _noisy(gl, p1=0.01, p2=0.05):
    out = GateList(gl.n_qubits)
    for kind, q0, q1, pidx, fixed in gl.tuples():
        if kind == 3:
            out.add_cnot(q0, q1)
            out.add_depol2(q0, q1, p2)
        else:
            out.add_rotation(kind, q0, fixed)
            out.add_depol1(q0, p1)
    return out


def _check_batch(sim_energies, oracle_energies, params, sample):
    e = sim_energies(params)
    assert e.shape == (params.shape[0],) and np.all(np.isfinite(e))
    idx = np.linspace(0, params.shape[0] - 1, sample).astype(int)
    assert np.abs(e[idx] - oracle_energies(params[idx])).max() < TOL
    # batch independence + determinism: a permuted batch gives the permuted energies, bit for bit
    perm = np.random.default_rng(0).permutation(params.shape[0])
    assert np.array_equal(sim_energies(params[perm]), e[perm])
    return e


def test_c1_four_qubit_dense_hamiltonian_latency_case(built_lib, oracle):
    """C1 (bench_workloads.build("C1")): 4 qubits, the reference's shipped LiH-4q dense parity Hamiltonian
    (tests/golden/lih_4q_parity.npz; it has no Pauli list: the trace projection gives 100 terms in 16 flip groups), loaded
    initial state from a 27-gate brickwork, 20 agent gates; B = 1 and B = 4096."""
    import bench_workloads
    w = bench_workloads.build("C1")
    assert w.dense.shape == (16, 16) and np.abs(w.dense - w.dense.conj().T).max() == 0 and w.groups == 16
    # the shipped eigvals are those of the shipped matrix (complex64 precision)
    assert abs(np.linalg.eigvalsh(w.dense).min() - w.eig_min) < 1e-6
    # the loaded state (made with plain numpy in bench_workloads) is what the oracle gets for the same circuit
    init_gl = synthetic_circuit(4, 27, 0)
    assert np.abs(oracle.state(init_gl, parameter_batch(init_gl, 1)[0]) - w.init).max() < 1e-14
    sim = w.bind(Simulator(w.n))
    one = w.params(1)
    assert abs(sim.energies(one)[0] - w.oracle_energies(one)[0]) < TOL
    e = _check_batch(sim.energies, w.oracle_energies, w.params(4096), 32)
    assert e.min() >= w.eig_min - 1e-6
    # the dense form through its Pauli decomposition (trace projection) agrees
    x, z, c = loaders.dense_to_pauli(w.dense)
    assert len(c) == 100
    sim.set_pauli_hamiltonian(x, z, c)
    assert abs(sim.energies(one)[0] - w.oracle_energies(one)[0]) < TOL
    sim.close()


def test_c2_beh2_trainable_256_parameter_sets(built_lib, oracle):
    """C2: BeH2-6q (the shipped substitute for the non-existent LiH-6q), trainable-env circuit from the shipped QPY
    (mirrored, negated, float32 angles) + agent gates, P ~ 105, B = 256."""
    import bench_workloads
    w = bench_workloads.build("C2")
    sim = w.bind(Simulator(w.n))
    e = _check_batch(sim.energies, w.oracle_energies, w.params(256), 24)
    assert e.min() >= w.eig_min - 1e-9   # variational bound against the shipped eigvals
    sim.close()


def test_c3_h2o_fixed_4096_cost_evaluations(built_lib, oracle):
    """C3: H2O-8q fixed env: TN state from the shipped QPY loaded, bit-reversed Hamiltonian, 20 agent gates, B = 4096."""
    import bench_workloads
    w = bench_workloads.build("C3")
    sim = w.bind(Simulator(w.n))
    e = _check_batch(sim.energies, w.oracle_energies, w.params(4096), 32)
    assert e.min() >= w.eig_min - 1e-9
    sim.close()


def test_c4_eight_qubit_depolarising_density_matrix(built_lib, oracle):
    """C4: environment_qulacs_noise semantics on H2O-8q: shipped QPY circuit (150 gates) + 40 agent gates, p1 = 0.01
    after every rotation, p2 = 0.05 after every CNOT, exact channels on rho (2^16 entries); B = 64."""
    import bench_workloads
    w = bench_workloads.build("C4")
    gl = w.gl
    assert gl.n_unitary >= 150 and w.mode == "dm"
    sim = w.bind(Simulator(w.n))
    params = w.params(64)
    e = sim.energies_dm(params)
    idx = [0, 21, 42, 63]
    assert np.abs(e[idx] - w.oracle_energies(params[idx])).max() < TOL
    assert np.array_equal(sim.energies_dm(params[::-1].copy()), e[::-1])
    # a depolarised state sits above the noise-free energy of the same circuit and inside the spectrum
    clean = sim.energies(params)
    assert np.all(e >= w.eig_min - 1e-9) and np.abs(e - clean).max() > 1e-3
    sim.close()


@pytest.mark.parametrize("name", ["C5L", "C5G"])
def test_c5_companions(built_lib, oracle, name):
    """C5's companions of the bench: a loaded dense initial state (known-zero skipping off) and 440 gates from the generic
    generator (no brick structure); two elements against the oracle, batch independence on eight."""
    import bench_workloads
    w = bench_workloads.build(name)
    sim = w.bind(Simulator(w.n))
    params = w.params(8)
    e = sim.energies(params)
    assert np.abs(e[:2] - w.oracle_energies(params[:2])).max() < TOL
    assert np.array_equal(sim.energies(params[::-1].copy()), e[::-1])
    sim.close()


def test_c5_twenty_qubit_heisenberg_bench_batch(built_lib, oracle):
    """C5 (the bench workload): 440-gate brickwork circuit, 77-term Heisenberg chain, B = 64; oracle on 3 elements,
    properties on all."""
    n = 20
    gl = brickwork_circuit(n, 21, 41, 5)
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    sim = Simulator(n)
    sim.set_circuit(gl)
    sim.set_pauli_hamiltonian(x, z, w)
    params = parameter_batch(gl, 64)
    e = sim.energies(params)
    idx = [0, 31, 63]
    assert np.abs(e[idx] - oracle.energies(gl, params[idx], pauli=(x, z, w))).max() < TOL
    assert np.array_equal(sim.energies(params[::-1].copy()), e[::-1])   # batch independence, bit for bit
    counts = sim.plan_counts(0)
    assert counts["tensor_core_blocks"] > 0 and counts["fp64_pipe_windows"] == 0   # the DMMA path is the one that ran
    # known answers: no gates -> E(|0..0>) = 2n - 1; Neel state (X on odd sites) -> -(n - 1)
    sim.set_circuit(GateList(n))
    assert abs(sim.energies(np.zeros((1, 1)))[0] - (2 * n - 1)) < TOL
    neel = GateList(n)
    for q in range(1, n, 2):
        neel.add_pauli("X", q)
    sim.set_circuit(neel)
    assert abs(sim.energies(np.zeros((1, 1)))[0] + (n - 1)) < TOL
    sim.close()
