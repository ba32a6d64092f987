"""Streaming pass kernel (tq_stream.cu: persistent CTAs, TMA tile I/O, -m gpu): multi-tile problems against the oracle, with
the kernel on and off.  Tolerances: 1e-10 Ha on energies (BASELINE.json north_star), 1e-12 on amplitudes."""
import numpy as np
import pytest

from tensorrl_qas_b200 import Simulator, loaders
from tensorrl_qas_b200.circuit import GateList, brickwork_circuit, parameter_batch, synthetic_circuit

pytestmark = pytest.mark.gpu


def heisenberg(n):
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    return x, z, w


def make_sim(n, gl, ham, monkeypatch, stream, init=None):
    monkeypatch.setenv("TQ_STREAM", "2" if stream else "0")   # 2: gate passes AND expectation-only passes
    monkeypatch.setenv("TQ_VALIDATE_PLAN", "1")
    sim = Simulator(n, 0)
    sim.set_pauli_hamiltonian(*ham)
    if init is not None:
        sim.set_init_state(init)
    sim.set_circuit(gl)
    return sim


@pytest.mark.parametrize("n,gates,seed,brick,batch", [(13, 150, 1, False, 5), (14, 300, 2, False, 3), (14, 120, 3, True, 4),
                                                        (16, 400, 4, False, 2), (17, 420, 6, True, 2), (15, 40, 7, False, 3),
                                                        (13, 12, 8, False, 2)])
def test_stream_energies_match_oracle(built_lib, oracle, monkeypatch, n, gates, seed, brick, batch):
    gl = brickwork_circuit(n, 21, max(0, gates - 21 * (n - 1)), seed) if brick else synthetic_circuit(n, gates, seed)
    ham = heisenberg(n)
    p = parameter_batch(gl, batch)
    want = oracle.energies(gl, p, pauli=ham)
    sim = make_sim(n, gl, ham, monkeypatch, True)
    got = sim.energies(p)
    again = sim.energies(p)
    launches = sim.plan_counts()["stream_launches"]
    sim.close()
    ref = make_sim(n, gl, ham, monkeypatch, False)
    off = ref.energies(p)
    assert ref.plan_counts()["stream_launches"] == 0
    ref.close()
    assert launches > 0, "the streaming kernel did not run"
    assert np.abs(got - want).max() < 1e-10
    assert np.abs(off - want).max() < 1e-10
    assert np.array_equal(got, again)   # deterministic


@pytest.mark.parametrize("n,gates,seed", [(13, 200, 11), (14, 260, 12), (15, 90, 13)])
def test_stream_states_match_oracle(built_lib, oracle, monkeypatch, n, gates, seed):
    gl = synthetic_circuit(n, gates, seed)
    p = parameter_batch(gl, 3)
    sim = make_sim(n, gl, heisenberg(n), monkeypatch, True)
    got = sim.states(p)
    launches = sim.plan_counts()["stream_launches"]
    sim.close()
    assert launches > 0
    for b in range(3):
        assert np.abs(got[b] - oracle.state(gl, p[b])).max() < 1e-12


def test_stream_loaded_initial_state(built_lib, oracle, monkeypatch):
    """A loaded initial state (the fixed environments' TN state) switches the known-zero skipping off: dense loads."""
    n = 14
    rng = np.random.default_rng(5)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    gl = synthetic_circuit(n, 220, 21)
    ham = heisenberg(n)
    p = parameter_batch(gl, 3)
    sim = make_sim(n, gl, ham, monkeypatch, True, init=psi)
    got = sim.energies(p)
    st = sim.states(p[:1])[0]
    launches = sim.plan_counts()["stream_launches"]
    sim.close()
    assert launches > 0
    assert np.abs(got - oracle.energies(gl, p, pauli=ham, init=psi)).max() < 1e-10
    assert np.abs(st - oracle.state(gl, p[0], init=psi)).max() < 1e-12


def test_stream_untouched_qubits(built_lib, oracle, monkeypatch):
    """Gates on a few qubits only: most tiles are known zeros (skipped, or zero-filled in the last gate pass)."""
    n = 15
    gl = GateList(n)
    rng = np.random.default_rng(3)
    for rep in range(3):
        for q in (0, 1, 5, 13, 14):
            gl.add_rotation(int(rng.integers(0, 3)), q, float(rng.uniform(-1, 1)))
        gl.add_cnot(0, 13)
        gl.add_cnot(14, 5)
        gl.add_cnot(1, 0)
    ham = heisenberg(n)
    p = parameter_batch(gl, 2)
    sim = make_sim(n, gl, ham, monkeypatch, True)
    got = sim.energies(p)
    st = sim.states(p[:1])[0]
    sim.close()
    assert np.abs(got - oracle.energies(gl, p, pauli=ham)).max() < 1e-10
    assert np.abs(st - oracle.state(gl, p[0])).max() < 1e-12


def test_stream_bench_shape_matches_oracle_sample(built_lib, oracle, monkeypatch):
    """The bench workload (C5, 20 qubits): two elements against the oracle, the whole batch against the kernel switched off."""
    import bench_workloads
    w = bench_workloads.build("C5")
    gl, ham = w.gl, w.pauli
    p = w.params(8)
    sim = make_sim(20, gl, ham, monkeypatch, True)
    got = sim.energies(p)
    launches = sim.plan_counts()["stream_launches"]
    sim.close()
    ref = make_sim(20, gl, ham, monkeypatch, False)
    off = ref.energies(p)
    ref.close()
    assert launches >= 3
    assert np.abs(got[:2] - oracle.energies(gl, p[:2], pauli=ham)).max() < 1e-10
    assert np.abs(got - off).max() < 1e-11


def test_stream_random_pauli_sum(built_lib, oracle, monkeypatch):
    """Random Pauli strings (odd numbers of Y, long Z strings, masks wider than a window): the rare expectation paths of
    the streaming kernel (imaginary class coefficients, index-dependent signs) and its fall-backs against the oracle."""
    n = 14
    rng = np.random.default_rng(11)
    x = np.zeros(40, dtype=np.uint64)
    z = np.zeros(40, dtype=np.uint64)
    for t in range(40):
        weight = int(rng.integers(1, 5)) if t < 30 else int(rng.integers(6, 10))
        for q in rng.choice(n, size=weight, replace=False):
            kind = int(rng.integers(0, 3)) if t % 3 else 2      # every third string is a pure Z string
            if kind in (0, 1):
                x[t] |= np.uint64(1 << int(q))
            if kind in (1, 2):
                z[t] |= np.uint64(1 << int(q))
    w = rng.normal(size=40)
    ham = (x, z, w)
    gl = synthetic_circuit(n, 160, 31)
    p = parameter_batch(gl, 3)
    want = oracle.energies(gl, p, pauli=ham)
    sim = make_sim(n, gl, ham, monkeypatch, True)
    got = sim.energies(p)
    sim.close()
    ref = make_sim(n, gl, ham, monkeypatch, False)
    off = ref.energies(p)
    ref.close()
    assert np.abs(off - want).max() < 1e-10
    assert np.abs(got - want).max() < 1e-10


@pytest.mark.parametrize("n,gates,seed,brick", [(17, 0, 4, True), (17, 60, 7, False), (18, 5, 0, True), (18, 41, 1, True),
                                                 (18, 30, 3, False), (18, 200, 4, False), (18, 100, 0, False), (18, 41, 3, True),
                                                 (18, 200, 5, False)])
def test_early_expectation_matches_oracle(built_lib, oracle, monkeypatch, n, gates, seed, brick):
    """Light-cone assignment (tq_plan.cpp attach_expectation): Hamiltonian groups that no later gate touches are evaluated
    in an earlier gate pass when that saves an expectation-only pass.  Every case here has such groups (checked on the
    plan); the energies must agree with the oracle and with the plan that evaluates everything on the final state."""
    from tensorrl_qas_b200.simulator import plan_dump
    gl = brickwork_circuit(n, 21, gates, seed) if brick else synthetic_circuit(n, gates, seed)
    masks = [(1 << q) | (1 << (q + 1)) for q in range(n - 1)]
    plan = plan_dump(gl, 16, 12, 4, cover_masks=masks)
    last_gate_pass = max(i for i, p in enumerate(plan) if p["ops"])
    assert any(p.get("exp_groups") for p in plan[:last_gate_pass]), "case does not exercise early evaluation"
    ham = heisenberg(n)
    p = parameter_batch(gl, 3)
    want = oracle.energies(gl, p, pauli=ham)
    for stream in (True, False):
        sim = make_sim(n, gl, ham, monkeypatch, stream)
        got = sim.energies(p)
        again = sim.energies(p)
        sim.close()
        assert np.abs(got - want).max() < 1e-10, (stream, np.abs(got - want).max())
        assert np.array_equal(got, again)
    monkeypatch.setenv("TQ_EARLY_EXPECT", "0")
    sim = make_sim(n, gl, ham, monkeypatch, True)
    late = sim.energies(p)
    sim.close()
    assert np.abs(late - want).max() < 1e-10


@pytest.mark.parametrize("n,gates,seed,brick", [(16, 60, 5, False), (16, 30, 6, False), (17, 41, 0, True), (18, 41, 3, True),
                                                 (18, 200, 5, False), (17, 200, 6, False), (18, 30, 1, False)])
def test_skip_last_store_matches_oracle(built_lib, oracle, monkeypatch, n, gates, seed, brick):
    """When no gate of the last gate pass touches a group evaluated in the expectation-only passes, that pass writes nothing
    back and they read its input (ExpPlan::last_store_needed, tq_plan.h).  Every case here has such a plan; energies against
    the oracle, against the plan that stores, and through tq_evolve_states (which does store: the states are its output)."""
    import torch
    from tensorrl_qas_b200.simulator import plan_dump
    gl = brickwork_circuit(n, 21, gates, seed) if brick else synthetic_circuit(n, gates, seed)
    masks = [(1 << q) | (1 << (q + 1)) for q in range(n - 1)]
    plan = plan_dump(gl, 16, 12, 4, cover_masks=masks)
    assert not plan[0]["last_store_needed"], "case does not exercise the skipped write-back"
    ham = heisenberg(n)
    p = parameter_batch(gl, 3)
    want = oracle.energies(gl, p, pauli=ham)
    for stream in (True, False):
        sim = make_sim(n, gl, ham, monkeypatch, stream)
        got = sim.energies(p)
        again = sim.energies(p)
        # the same circuit applied to |0...0> states in place: the final states come back, the energies agree
        st = torch.zeros((3, 1 << n), dtype=torch.complex128, device="cuda")
        st[:, 0] = 1.0
        e_ev = sim.evolve_states(st, torch.as_tensor(p, device="cuda").contiguous(), energies=True)
        torch.cuda.synchronize()
        psi = sim.states(p)
        sim.close()
        assert np.abs(got - want).max() < 1e-10, (stream, np.abs(got - want).max())
        assert np.array_equal(got, again)
        assert np.abs(e_ev.cpu().numpy() - want).max() < 1e-10
        assert np.abs(st.cpu().numpy() - psi).max() < 1e-12
    monkeypatch.setenv("TQ_SKIP_LAST_STORE", "0")
    sim = make_sim(n, gl, ham, monkeypatch, True)
    stored = sim.energies(p)
    sim.close()
    assert np.abs(stored - want).max() < 1e-10


def chain_hamiltonian(n, kind, seed=0):
    """Nearest-neighbour chains: 'xxz' = J (XX + YY) + D ZZ with one J per bond (exchange classes with one coefficient),
    'xyz' = independent weights on XX and YY (not an exchange class: the generic ops), 'gaps' = every third bond missing,
    'field' = Heisenberg bonds + a Z field (diagonal terms on single qubits)."""
    rng = np.random.default_rng(seed)
    paulis, w = [], []
    for q in range(n - 1):
        if kind == "gaps" and q % 3 == 2:
            continue
        j, d = rng.normal(), rng.normal()
        jy = rng.normal() if kind == "xyz" else j
        for s, c in (("XX", j), ("YY", jy), ("ZZ", d)):
            paulis.append("I" * q + s + "I" * (n - q - 2))
            w.append(c)
    if kind == "field":
        for q in range(n):
            paulis.append("I" * q + "Z" + "I" * (n - q - 1))
            w.append(rng.normal())
    x, z = loaders.pauli_masks(paulis, n)
    return x, z, np.asarray(w)


@pytest.mark.parametrize("n,kind,brick,seed", [(14, "xxz", False, 1), (16, "xxz", True, 2), (15, "gaps", False, 3),
                                                (14, "xyz", False, 4), (15, "field", True, 5)])
def test_chain_windows_match_generic_ops_and_oracle(built_lib, oracle, monkeypatch, n, kind, brick, seed):
    """Expectation windows of nearest-neighbour chains run as one fused routine (tq_stream.cu, chain_window); it must agree
    with the dispatched ops (TQ_STREAM_CHAIN=0) and with the oracle, for per-bond couplings, missing bonds, XX != YY (which
    must stay on the generic ops) and extra single-qubit diagonal terms."""
    gl = brickwork_circuit(n, 21, 40, seed) if brick else synthetic_circuit(n, 150, seed)
    ham = chain_hamiltonian(n, kind, seed)
    p = parameter_batch(gl, 3)
    want = oracle.energies(gl, p, pauli=ham)
    sim = make_sim(n, gl, ham, monkeypatch, True)
    got = sim.energies(p)
    again = sim.energies(p)
    assert sim.plan_counts()["stream_launches"] > 0
    sim.close()
    monkeypatch.setenv("TQ_STREAM_CHAIN", "0")
    ref = make_sim(n, gl, ham, monkeypatch, True)
    generic = ref.energies(p)
    ref.close()
    assert np.abs(generic - want).max() < 1e-10
    assert np.abs(got - want).max() < 1e-10
    assert np.abs(got - generic).max() < 1e-11
    assert np.array_equal(got, again)
