"""GPU parity suite (-m gpu): libtqsim through its C ABI (ctypes) against the oracle and the committed golden
vectors.  Tolerance: 1e-10 Ha absolute energy (BASELINE.json north_star); states compared at 1e-12."""
import numpy as np
import pytest

from tensorrl_qas_b200 import Simulator, TqError, loaders
from tensorrl_qas_b200.circuit import GateList, brickwork_circuit, parameter_batch, synthetic_circuit

from golden_util import CASES, Case

pytestmark = pytest.mark.gpu
TOL = 1e-10


def noisy_copy(gl, p1=0.01, p2=0.05):
    out = GateList(gl.n_qubits)
    for kind, q0, q1, pidx, fixed in gl.tuples():
        if kind == 3:
            out.add_cnot(q0, q1)
            out.add_depol2(q0, q1, p2)
        else:
            out.add_rotation(kind, q0, fixed)
            out.add_depol1(q0, p1)
    return out


def random_pauli_sum(n, T, seed, max_flips=None):
    """random Pauli strings; max_flips bounds the number of X/Y factors (a flip mask must fit one tile next to the
    forced low bits when n exceeds the tile size)"""
    rng = np.random.default_rng(seed)
    if max_flips is None:
        x = rng.integers(0, 1 << n, size=T).astype(np.uint64)
    else:
        x = np.zeros(T, dtype=np.uint64)
        for t in range(T):
            for q in rng.choice(n, size=int(rng.integers(0, max_flips + 1)), replace=False):
                x[t] |= np.uint64(1 << int(q))
    z = rng.integers(0, 1 << n, size=T).astype(np.uint64)
    return x, z, rng.normal(size=T)


@pytest.mark.parametrize("key", CASES)
def test_golden_energies_both_env_conventions(built_lib, key):
    """Reference-code golden energies (tests/golden/make_golden.py) through the C ABI, dense and Pauli Hamiltonians."""
    c = Case(key)
    g = c.g
    sim = Simulator(c.n)
    gl = c.gatelist("in")
    sim.set_circuit(gl)
    sim.set_dense_hamiltonian(c.dense(False))
    assert np.abs(sim.energies(g["in_X"]) - g["in_E"]).max() < TOL
    assert abs(sim.energies([gl.initial_angles])[0] - float(g["in_e_tensor"])) < TOL
    x, z = c.masks(False)
    sim.set_pauli_hamiltonian(x, z, c.weights)
    assert np.abs(sim.energies(g["in_X"]) - g["in_E"]).max() < TOL
    # fixed environments: loaded TN state + bit-reversed Hamiltonian
    gl2 = c.gatelist("notin")
    sim.set_circuit(gl2)
    sim.set_init_state(g["notin_tn_state"])
    sim.set_dense_hamiltonian(c.dense(True))
    assert np.abs(sim.energies(g["notin_X"]) - g["notin_E"]).max() < TOL
    x, z = c.masks(True)
    sim.set_pauli_hamiltonian(x, z, c.weights)
    assert np.abs(sim.energies(g["notin_X"]) - g["notin_E"]).max() < TOL
    sim.set_circuit(GateList(c.n))
    assert abs(sim.energies(np.zeros((1, 1)))[0] - float(g["notin_e_first"])) < TOL
    sim.set_init_state(None)
    assert abs(sim.energies(np.zeros((1, 1)))[0] - c.h00) < TOL  # E(|0..0>) = H[0,0] (un-reversed == reversed at index 0)
    sim.close()


@pytest.mark.parametrize("key", ["beh2_6q", "h2o_8q"])
def test_golden_noise_trajectories(built_lib, key):
    c = Case(key)
    g = c.g
    sim = Simulator(c.n)
    gl = c.gatelist("in", noise=(0.01, 0.05))
    sim.set_circuit(gl)
    sim.set_dense_hamiltonian(c.dense(False))
    X = np.stack([g["in_X"][r % 6] for r in range(len(g["noise_E"]))])
    assert np.abs(sim.energies_traj(X, g["noise_codes"]) - g["noise_E"]).max() < TOL
    # noise-free entry point ignores the noise gates
    assert np.abs(sim.energies(g["in_X"]) - g["in_E"]).max() < TOL
    sim.close()


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 9, 10, 11, 12])
def test_random_circuits_single_tile_vs_oracle(built_lib, oracle, n):
    gl = synthetic_circuit(n, 90, 100 + n)
    gl.add_pauli("X", 0)
    gl.add_pauli("Y", n - 1)
    gl.add_pauli("Z", n // 2)
    params = parameter_batch(gl, 5)
    sim = Simulator(n)
    sim.set_circuit(gl)
    states = sim.states(params)
    for b in range(5):
        assert np.abs(states[b] - oracle.state(gl, params[b])).max() < 1e-12
    x, z, w = random_pauli_sum(n, 9, n)
    sim.set_pauli_hamiltonian(x, z, w)
    assert np.abs(sim.energies(params) - oracle.energies(gl, params, pauli=(x, z, w))).max() < TOL
    rng = np.random.default_rng(n)
    init = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    init /= np.linalg.norm(init)
    sim.set_init_state(init)
    assert np.abs(sim.energies(params) - oracle.energies(gl, params, pauli=(x, z, w), init=init)).max() < TOL
    sim.close()


@pytest.mark.parametrize("n,gates", [(13, 60), (14, 80), (16, 120), (18, 150)])
def test_random_circuits_multi_tile_vs_oracle(built_lib, oracle, n, gates):
    """n above the tile size: several fused passes + Pauli-group expectation (+ expectation-only passes)."""
    gl = synthetic_circuit(n, gates, 200 + n)
    params = parameter_batch(gl, 3)
    sim = Simulator(n)
    sim.set_circuit(gl)
    info = sim.plan_info(0)
    assert info["tile_bits"] == 12
    states = sim.states(params[:2])
    for b in range(2):
        assert np.abs(states[b] - oracle.state(gl, params[b])).max() < 1e-12
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    xr, zr, wr = random_pauli_sum(n, 6, n, max_flips=6)
    x, z, w = np.concatenate([x, xr]), np.concatenate([z, zr]), np.concatenate([w, wr])
    sim.set_pauli_hamiltonian(x, z, w)
    got = sim.energies(params)
    want = oracle.energies(gl, params, pauli=(x, z, w))
    assert np.abs(got - want).max() < TOL
    assert np.array_equal(got, sim.energies(params))  # deterministic reduction: bit-identical re-run
    sim.close()


def test_heisenberg_20_known_answers_and_properties(built_lib, oracle):
    """BASELINE config 5 at full size through size-independent properties: E(|0..0>) = 39, E(Neel) = -19, norm
    preserved, batch elements independent; plus a direct oracle comparison on two elements."""
    n = 20
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    sim = Simulator(n)
    sim.set_pauli_hamiltonian(x, z, w)
    sim.set_circuit(GateList(n))
    assert abs(sim.energies(np.zeros((1, 1)))[0] - 39.0) < 1e-12
    neel = GateList(n)
    for q in range(1, n, 2):
        neel.add_pauli("X", q)
    sim.set_circuit(neel)
    assert abs(sim.energies(np.zeros((1, 1)))[0] + 19.0) < 1e-12
    gl = brickwork_circuit(n, 21, 41, 5)
    assert len(gl) == 440
    sim.set_circuit(gl)
    info = sim.plan_info(0)
    assert info["gate_passes"] <= 6 and info["groups"] == 20
    params = parameter_batch(gl, 4)
    e = sim.energies(params)
    want = oracle.energies(gl, params[:2], pauli=(x, z, w))
    assert np.abs(e[:2] - want).max() < TOL
    # batch independence: element 2 alone gives the same bits
    assert sim.energies(params[2:3])[0] == e[2]
    st = sim.states(params[:1])[0]
    assert abs(np.vdot(st, st).real - 1.0) < 1e-12
    # identity: -19 <= E/.. bounded by the spectrum of the open chain: |E| <= 3(n-1) + n
    assert np.all(np.abs(e) <= 3 * (n - 1) + n)
    sim.close()


@pytest.mark.parametrize("n", [2, 3, 5, 6, 8])
def test_density_matrix_path_vs_oracle(built_lib, oracle, n):
    gl = noisy_copy(synthetic_circuit(n, 24 if n < 8 else 40, 300 + n))
    params = parameter_batch(gl, 3)
    paulis, w = loaders.heisenberg_terms(max(n, 2))
    x, z = loaders.pauli_masks(paulis, n)
    sim = Simulator(n)
    sim.set_circuit(gl)
    sim.set_pauli_hamiltonian(x, z, w)
    got = sim.energies_dm(params)
    want = oracle.dm_energies(gl, params, pauli=(x, z, w))
    assert np.abs(got - want).max() < TOL
    if n <= 6:
        rho = sim.density_matrices(params[:1])[0]
        ref = oracle.density_matrix(gl, params[0])
        assert np.abs(rho - ref).max() < 1e-12
        assert abs(np.trace(rho) - 1) < 1e-12 and np.abs(rho - rho.conj().T).max() < 1e-12
    # p = 0 channel == pure-state path
    clean = noisy_copy(synthetic_circuit(n, 24, 300 + n), 0.0, 0.0)
    sim.set_circuit(clean)
    p2 = parameter_batch(clean, 2)
    assert np.abs(sim.energies_dm(p2) - sim.energies(p2)).max() < 1e-12
    sim.close()


def test_density_matrix_with_loaded_state_and_dense_h(built_lib, oracle):
    c = Case("beh2_6q")
    gl = c.gatelist("notin", noise=(0.01, 0.05))
    sim = Simulator(c.n)
    sim.set_circuit(gl)
    sim.set_init_state(c.g["notin_tn_state"])
    H = c.dense(True)
    sim.set_dense_hamiltonian(H)
    got = sim.energies_dm(c.g["notin_X"])
    want = oracle.dm_energies(gl, c.g["notin_X"], dense=H, init=c.g["notin_tn_state"])
    assert np.abs(got - want).max() < TOL
    sim.close()


def test_trajectory_mean_converges_to_density_matrix(built_lib):
    from tensorrl_qas_b200.VQAs._backend import sample_noise_codes
    n = 5
    gl = noisy_copy(synthetic_circuit(n, 30, 9), 0.02, 0.08)
    p = parameter_batch(gl, 1)
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    sim = Simulator(n)
    sim.set_circuit(gl)
    sim.set_pauli_hamiltonian(x, z, w)
    exact = sim.energies_dm(p)[0]
    S = 20000
    codes = sample_noise_codes(gl, np.random.default_rng(3), S)
    e = sim.energies_traj(np.repeat(p, S, axis=0), codes)
    assert abs(e.mean() - exact) < 5 * e.std() / np.sqrt(S)
    sim.close()


def test_device_buffer_entry_point_and_stream(built_lib, oracle):
    import torch
    n = 10
    gl = synthetic_circuit(n, 60, 4)
    params = parameter_batch(gl, 64)
    x, z, w = random_pauli_sum(n, 8, 1)
    sim = Simulator(n)
    sim.set_circuit(gl)
    sim.set_pauli_hamiltonian(x, z, w)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        p_dev = torch.from_numpy(params).cuda()
        out = sim.energies_dev(p_dev)
    stream.synchronize()
    assert np.abs(out.cpu().numpy() - oracle.energies(gl, params, pauli=(x, z, w))).max() < TOL
    assert sim.launch_count >= 1
    sim.close()


def test_errors_are_reported_not_fatal(built_lib):
    sim = Simulator(4)
    with pytest.raises(TqError):
        sim.energies(np.zeros((1, 1)))  # no circuit
    gl = GateList(4)
    gl.add_rotation(0, 1, 0.3)
    sim.set_circuit(gl)
    with pytest.raises(TqError):
        sim.energies(np.zeros((1, 1)))  # no Hamiltonian
    bad = GateList(4)
    bad.add_cnot(2, 2)
    with pytest.raises(TqError):
        sim.set_circuit(bad)
    sim.set_pauli_hamiltonian([1], [0], [1.0])
    assert abs(sim.energies([[0.3]])[0]) < 1e-15  # <X_0> on RX(q1)|0000> = 0
    assert sim.energies(np.zeros((0, 1))).shape == (0,)  # empty batch
    sim.close()
    with pytest.raises(TqError):
        Simulator(40)


@pytest.mark.parametrize("key", ["h2o_8q", "heis_5q"])
def test_vqa_shims_mirror_reference_calls(built_lib, key):
    """The drop-in modules called exactly the way environment_qulacs*.py calls the reference's VQE_qulacs*.py."""
    import torch
    from tensorrl_qas_b200.VQAs import VQE_qulacs as vc
    from tensorrl_qas_b200.VQAs import VQE_qulacs_TN_notin_RL as vc2
    from tensorrl_qas_b200.VQAs import VQE_qulacs_noise as vcn
    from tensorrl_qas_b200.VQAs import _backend
    c = Case(key)
    g = c.g
    H = c.dense(False)
    state = torch.from_numpy(g["in_tensor"])
    circ = vc.Parametric_Circuit(n_qubits=c.n, noise_models=[], noise_values=[]).construct_ansatz(state)
    assert circ.get_parameter_count() == g["in_X"].shape[1]
    e = vc.get_exp_val(c.n, circ, H)
    assert isinstance(e, np.float64) and abs(e - float(g["in_e_tensor"])) < TOL
    for r in range(3):
        x32 = g["in_X"][r].astype(np.float32)  # pyprima COBYLA hands float32 trial points
        e = vc.get_energy_qulacs(x32, observable=H, circuit=circ, n_qubits=c.n, n_shots=0, phys_noise=False, which_angles=[])
        sim_ref = vc.get_energy_qulacs_batch(x32.astype(np.float64).reshape(1, -1), H, circ, c.n)[0]
        assert e == sim_ref
        e64 = vc.get_energy_qulacs(g["in_X"][r], observable=H, circuit=circ, n_qubits=c.n, n_shots=0)
        assert abs(e64 - g["in_E"][r]) < TOL
    Hrev = c.dense(True)
    st2 = torch.from_numpy(g["notin_tensor"])
    circ2 = vc2.Parametric_Circuit(n_qubits=c.n).construct_ansatz(st2)
    for r in range(3):
        e = vc2.get_energy_qulacs(g["notin_X"][r], observable=Hrev, circuit=circ2, n_qubits=c.n,
                                  TN_state=g["notin_tn_state"], n_shots=0, phys_noise=False, which_angles=[])
        assert abs(e - g["notin_E"][r]) < TOL
    # back to the |0..0> start after a TN-state evaluation on the same handle
    assert abs(vc.get_exp_val(c.n, circ, H) - vc.get_energy_qulacs_batch(circ.params.reshape(1, -1), H, circ, c.n)[0]) == 0
    if key == "h2o_8q":
        circ3 = vcn.Parametric_Circuit(n_qubits=c.n).construct_ansatz(state)
        vcn.NOISE_MODE = "density_matrix"
        e_dm = vcn.get_energy_qulacs(g["in_X"][0], observable=H, circuit=circ3, weights=c.weights, n_qubits=c.n, n_shots=0)
        vcn.NOISE_MODE = "trajectory"
        vcn.seed(5)
        es = vcn.get_energy_qulacs_batch(np.repeat(g["in_X"][:1], 4000, axis=0), H, circ3, c.n)
        assert abs(es.mean() - e_dm) < 5 * es.std() / np.sqrt(len(es))
        assert c.eig_min - 1e-9 < e_dm < 0
    _backend.reset_backends()


@pytest.mark.parametrize("n,gates", [(9, 1500), (12, 2500), (14, 1200)])
def test_long_circuits_many_windows_per_pass(built_lib, oracle, n, gates):
    """Passes with more windows than the kernel keeps headers for in shared memory (32) and more ops than it stages at
    once (48): the header and op restaging paths of the tensor-core kernel."""
    from tensorrl_qas_b200.simulator import plan_dump
    gl = synthetic_circuit(n, gates, 77 + n)
    plan = plan_dump(gl, 0, 12, 3)
    if n <= 12:
        assert max(len(p["windows"]) for p in plan) > 32
    else:
        assert len(plan) > 10   # many passes instead
    params = parameter_batch(gl, 2)
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    sim = Simulator(n)
    sim.set_circuit(gl)
    sim.set_pauli_hamiltonian(x, z, w)
    st = sim.states(params[:1])[0]
    assert np.abs(st - oracle.state(gl, params[0])).max() < 1e-11
    assert np.abs(sim.energies(params) - oracle.energies(gl, params, pauli=(x, z, w))).max() < TOL
    sim.close()


def test_tensor_core_path_with_loaded_state_trajectories_and_imaginary_terms(built_lib, oracle):
    """13 qubits (two tiles per state): loaded initial state, sampled-Pauli trajectories, and a Hamiltonian whose terms
    have odd numbers of Y factors (imaginary class coefficients in the expectation windows)."""
    from tensorrl_qas_b200.VQAs._backend import sample_noise_codes
    n = 13
    base = synthetic_circuit(n, 70, 5)
    init = oracle.state(synthetic_circuit(n, 40, 6), parameter_batch(synthetic_circuit(n, 40, 6), 1)[0])
    x, z, w = random_pauli_sum(n, 40, 9, max_flips=4)
    sim = Simulator(n)
    sim.set_init_state(init)
    sim.set_pauli_hamiltonian(x, z, w)
    sim.set_circuit(base)
    p = parameter_batch(base, 3)
    assert np.abs(sim.energies(p) - oracle.energies(base, p, pauli=(x, z, w), init=init)).max() < TOL
    noisy = noisy_copy(base, 0.05, 0.1)
    sim.set_circuit(noisy)
    pn = parameter_batch(noisy, 4)
    codes = sample_noise_codes(noisy, np.random.default_rng(4), 4)
    assert codes.any()
    got = sim.energies_traj(pn, codes)
    want = oracle.energies(noisy, pn, pauli=(x, z, w), init=init, codes=codes)
    assert np.abs(got - want).max() < TOL
    sim.close()


def test_zero_skipping_from_the_all_zero_state(built_lib, oracle, monkeypatch):
    """A run from |0...0> skips the tiles / amplitudes no gate has populated yet (TQ_SPARSE_INIT): same energies and
    states, bit for bit, as with the skipping switched off, and equal to the oracle -- including circuits whose first
    passes touch only a few (high) qubits, CNOTs controlled by untouched qubits, and back-to-back calls that leave stale
    data in the skipped regions of the scratch buffer."""
    n = 16
    rng = np.random.default_rng(123)
    circuits = []
    gl = GateList(n)                      # (a) only high qubits at first, then a CNOT fan-out, then everything
    from tensorrl_qas_b200.circuit import append_random_gates
    append_random_gates(gl, 30, rng, (13, 14, 15))
    gl.add_cnot(2, 9)                     # control never populated: must stay a no-op
    gl.add_cnot(15, 3)
    append_random_gates(gl, 60, rng, (3, 4, 5, 15))
    append_random_gates(gl, 80, rng)
    circuits.append(gl)
    circuits.append(synthetic_circuit(n, 25, 5, qubits=(0, 1)))            # (b) most of the register untouched at the end
    circuits.append(brickwork_circuit(n, 21, 30, 8))                        # (c) the bench shape
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    results = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("TQ_SPARSE_INIT", flag)
        sim = Simulator(n)
        sim.set_pauli_hamiltonian(x, z, w)
        out = []
        for gl in circuits + circuits[:1]:          # the first circuit again at the end: stale scratch contents
            sim.set_circuit(gl)
            p = parameter_batch(gl, 3)
            out.append((sim.energies(p), sim.states(p[:1])[0]))
        results[flag] = out
        sim.close()
    for (e1, s1), (e0, s0), gl in zip(results["1"], results["0"], circuits + circuits[:1]):
        assert np.array_equal(e1, e0) and np.array_equal(s1, s0)
        p = parameter_batch(gl, 3)
        assert np.abs(e1 - oracle.energies(gl, p, pauli=(x, z, w))).max() < TOL
        assert np.abs(s1 - oracle.state(gl, p[0])).max() < 1e-12


def test_out_of_memory_is_an_error_code(built_lib, monkeypatch):
    """TQ_ENOMEM with a message instead of an abort: a state larger than the scratch limit, and an impossible device
    allocation (include/tqsim.h error model)."""
    import ctypes
    from tensorrl_qas_b200 import _lib
    from tensorrl_qas_b200.simulator import TqError
    monkeypatch.setenv("TQ_MAX_SCRATCH_MB", "8")
    n = 20
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    gl = brickwork_circuit(n, 21, 41, 5)
    sim = Simulator(n, 0)
    sim.set_pauli_hamiltonian(x, z, w)
    sim.set_circuit(gl)
    with pytest.raises(TqError) as e:
        sim.energies(parameter_batch(gl, 2))
    assert e.value.code == -4 and "TQ_MAX_SCRATCH_MB" in str(e.value)
    sim.close()
    out = ctypes.c_void_p()
    assert _lib.lib().tq_device_alloc(0, 1 << 46, ctypes.byref(out)) == -4 and not out.value


def test_plan_cache_restores_plans_bit_for_bit(built_lib, oracle, monkeypatch):
    """tq_set_circuit keeps compiled plans per gate list: re-binding the bound circuit is a no-op,
    a circuit seen before comes back from the cache, and the energies are the ones of a fresh compilation, bit for bit."""
    n = 8
    c = Case("h2o_8q")
    H = c.dense(False)
    circuits = [synthetic_circuit(n, 30 + 3 * i, 70 + i) for i in range(5)]
    params = [parameter_batch(g, 3) for g in circuits]
    sim = Simulator(n, 0)
    sim.set_dense_hamiltonian(H)
    first = []
    for g, p in zip(circuits, params):
        sim.set_circuit(g)
        sim.set_circuit(g)                      # what an environment step does: the same structure twice
        first.append(sim.energies(p))
    st = sim.plan_cache_stats()
    assert st["same"] == 5 and st["misses"] == 5 and st["hits"] == 0
    for i in (2, 0, 4, 1, 3):                   # revisit in another order: every plan comes from the cache
        sim.set_circuit(circuits[i])
        assert np.array_equal(sim.energies(params[i]), first[i])
    st = sim.plan_cache_stats()
    assert st["hits"] == 5 and st["misses"] == 5
    for g, p, e in zip(circuits, params, first):
        assert np.abs(e - oracle.energies(g, p, dense=H)).max() < 1e-10
    sim.set_dense_hamiltonian(2.0 * H)          # a new Hamiltonian drops the cache
    assert sim.plan_cache_stats()["entries"] == 0
    sim.set_circuit(circuits[0])
    assert np.abs(sim.energies(params[0]) - 2.0 * first[0]).max() < 1e-9
    sim.close()
    monkeypatch.setenv("TQ_PLAN_CACHE", "0")    # switched off: every new circuit is compiled
    sim = Simulator(n, 0)
    sim.set_dense_hamiltonian(H)
    for i in (0, 1, 0):
        sim.set_circuit(circuits[i])
        assert np.array_equal(sim.energies(params[i]), first[i])
    assert sim.plan_cache_stats()["hits"] == 0
    sim.close()


@pytest.mark.parametrize("n,dense", [(4, True), (6, True), (8, True), (7, False)])
def test_launch_shape_does_not_change_the_bits(built_lib, oracle, monkeypatch, n, dense):
    """One evaluation per call (the latency path: block matrices evaluated inside the pass, four warps staging), a batch that
    fills the device (one warp per CTA) and the separate prep kernel (TQ_FUSE_PREP=0) must give the SAME bits for the same
    parameter set: COBYLA trajectories depend on it (serial calls, batched calls and lock-step launches are interchangeable)."""
    gl = synthetic_circuit(n, 60, 40 + n)
    rng = np.random.default_rng(n)
    p = parameter_batch(gl, 700)          # > 2 CTAs per SM: the batched launch shape
    if dense:
        h = rng.normal(size=(1 << n, 1 << n)) + 1j * rng.normal(size=(1 << n, 1 << n))
        h = h + h.conj().T
        kw = dict(dense=h)
    else:
        kw = dict(pauli=random_pauli_sum(n, 12, 3))

    def run(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        sim = Simulator(n, 0)
        sim.set_circuit(gl)
        if dense:
            sim.set_dense_hamiltonian(h)
        else:
            sim.set_pauli_hamiltonian(*kw["pauli"])
        batched = sim.energies(p)
        serial = np.array([sim.energies(p[i:i + 1])[0] for i in (0, 1, 350, 699)])
        sim.close()
        for k in env:
            monkeypatch.delenv(k)
        return batched, serial

    batched, serial = run({})
    unfused, serial_unfused = run({"TQ_FUSE_PREP": "0"})
    assert np.abs(batched - oracle.energies(gl, p, **kw)).max() < TOL
    assert np.array_equal(serial, batched[[0, 1, 350, 699]])
    assert np.array_equal(unfused, batched)
    assert np.array_equal(serial_unfused, serial)
