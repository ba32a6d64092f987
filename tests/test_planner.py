"""CPU tests of the circuit compiler (tq_plan.cpp: fusion, pass packing, register windows) through the no-GPU
tq_plan_dump entry point."""
import os

import numpy as np
import pytest

from tensorrl_qas_b200.circuit import GateList, brickwork_circuit, parameter_batch, synthetic_circuit
from tensorrl_qas_b200.simulator import plan_dump

from plan_emulator import check_invariants, run_plan, run_plan_windows


def noisy_circuit(n, g, seed):
    base = synthetic_circuit(n, g, seed)
    gl = GateList(n)
    for kind, q0, q1, pidx, fixed in base.tuples():
        if kind == 3:
            gl.add_cnot(q0, q1)
            gl.add_depol2(q0, q1, 0.05)
        else:
            gl.add_rotation(kind, q0, fixed)
            gl.add_depol1(q0, 0.01)
    return gl


def n_blocks(plan):
    return sum(len(p["ops"]) for p in plan["passes"])


@pytest.mark.parametrize("n,tile_bits,low_bits,seed", [(1, 12, 4, 9), (2, 12, 4, 8), (3, 12, 4, 7), (5, 12, 4, 0),
                                                        (9, 8, 4, 1), (10, 8, 2, 2), (11, 8, 4, 3), (12, 9, 3, 4),
                                                        (12, 12, 4, 5)])
def test_pure_plan_equals_gate_list(built_lib, oracle, n, tile_bits, low_bits, seed):
    gl = synthetic_circuit(n, 120 if n < 12 else 600, seed)
    gl.add_pauli("X", 0)
    gl.add_pauli("Y", n - 1)
    gl.add_pauli("Z", n // 2)
    params = parameter_batch(gl, 1)[0]
    plan = plan_dump(gl, 0, tile_bits, low_bits, with_mats=True)
    check_invariants(plan, n, tile_bits, low_bits)
    # every gate lands in exactly one block program (a lone CNOT owns a one-gate program of its own)
    assert sum(len(m["gates"]) for m in plan["mats"]) == len(gl)
    want = oracle.state(gl, params)
    assert np.abs(run_plan(plan, n, params) - want).max() < 1e-12
    assert np.abs(run_plan_windows(plan, n, params) - want).max() < 1e-12
    if n > tile_bits:
        assert len(plan["passes"]) > 1
    if n >= 2:
        assert n_blocks(plan) < len(gl)  # fusion merged something


@pytest.mark.parametrize("n,tile_bits,low_bits,seed,gates", [(9, 12, 4, 21, 200), (10, 9, 3, 22, 300), (13, 10, 3, 23, 400),
                                                              (14, 12, 3, 24, 500), (13, 11, 2, 25, 900)])
def test_tensor_core_windows_equal_gate_list(built_lib, oracle, n, tile_bits, low_bits, seed, gates):
    """Tiles of >= 2^9 amplitudes are scheduled as DMMA windows (QL + five register qubits, M_SWAPQL relabelling)."""
    gl = synthetic_circuit(n, gates, seed)
    gl.add_pauli("Z", 0)
    params = parameter_batch(gl, 1)[0]
    plan = plan_dump(gl, 0, tile_bits, low_bits, with_mats=True)
    check_invariants(plan, n, tile_bits, low_bits)
    wins = [w for p in plan["passes"] for w in p["windows"]]
    assert wins and all(w.get("mma") for w in wins)
    codes = {o[0] for w in wins for o in w["ops"]}
    assert 16 in codes  # dense blocks on the tensor cores
    assert np.abs(run_plan_windows(plan, n, params) - oracle.state(gl, params)).max() < 1e-12


def test_register_windows_still_available(built_lib, oracle, monkeypatch):
    """TQ_MMA=0 keeps the FP64-pipe register-window schedule for large tiles too (A/B comparisons)."""
    monkeypatch.setenv("TQ_MMA", "0")
    gl = synthetic_circuit(12, 300, 31)
    params = parameter_batch(gl, 1)[0]
    plan = plan_dump(gl, 0, 12, 4, with_mats=True)
    assert not any(w.get("mma") for p in plan["passes"] for w in p["windows"])
    assert np.abs(run_plan_windows(plan, 12, params) - oracle.state(gl, params)).max() < 1e-12


def test_unfused_plan_matches_too(built_lib, oracle, monkeypatch):
    monkeypatch.setenv("TQ_FUSE", "0")
    gl = synthetic_circuit(9, 80, 3)
    params = parameter_batch(gl, 1)[0]
    plan = plan_dump(gl, 0, 8, 4, with_mats=True)
    assert n_blocks(plan) == len(gl)
    assert np.abs(run_plan_windows(plan, 9, params) - oracle.state(gl, params)).max() < 1e-12


def test_brickwork_c5_plan_is_few_passes(built_lib, oracle):
    gl = brickwork_circuit(20, 21, 41, 5)
    assert len(gl) == 440
    plan = plan_dump(gl, 0, 12, 4, with_mats=True)
    check_invariants(plan, 20, 12, 4)
    assert len(plan["passes"]) <= 6  # 440 gates fused into a handful of HBM passes
    n_windows = sum(len(p["windows"]) for p in plan["passes"])
    assert n_windows <= 40, n_windows  # ... a few dozen shared-memory redistributions
    assert n_blocks(plan) <= 80, n_blocks(plan)  # ... and a few dozen dense blocks (19 bricks + agent gates)
    # a 12-qubit circuit of the same shape is cheap enough to emulate end to end
    g12 = brickwork_circuit(12, 21, 30, 6)
    params = parameter_batch(g12, 1)[0]
    p12 = plan_dump(g12, 0, 8, 4, with_mats=True)
    assert np.abs(run_plan_windows(p12, 12, params) - oracle.state(g12, params)).max() < 1e-12


def test_cover_masks_pulled_into_last_pass(built_lib):
    gl = synthetic_circuit(14, 10, 7, qubits=(0, 1, 2))
    mask = (1 << 12) | (1 << 13)
    passes = plan_dump(gl, 0, 10, 4, cover_masks=[mask])
    assert 12 in passes[-1]["local"] and 13 in passes[-1]["local"]


@pytest.mark.parametrize("n,tile_bits,seed", [(3, 12, 0), (5, 8, 1), (6, 8, 2)])
def test_density_plan_equals_oracle(built_lib, oracle, n, tile_bits, seed):
    gl = noisy_circuit(n, 14, seed)
    params = parameter_batch(gl, 1)[0]
    plan = plan_dump(gl, 1, tile_bits, 4, with_mats=True)
    check_invariants(plan, 2 * n, tile_bits, 4)
    want = oracle.density_matrix(gl, params).reshape(-1)
    got = run_plan(plan, 2 * n, params)
    assert np.abs(got - want).max() < 1e-12
    assert np.abs(run_plan_windows(plan, 2 * n, params) - want).max() < 1e-12
    rho = got.reshape(1 << n, 1 << n)
    assert abs(np.trace(rho) - 1) < 1e-12 and np.abs(rho - rho.conj().T).max() < 1e-12


def test_density_plan_without_noise_fuses(built_lib, oracle):
    n = 5
    gl = synthetic_circuit(n, 40, 11)
    params = parameter_batch(gl, 1)[0]
    plan = plan_dump(gl, 1, 8, 4, with_mats=True)
    want = oracle.density_matrix(gl, params).reshape(-1)
    assert np.abs(run_plan_windows(plan, 2 * n, params) - want).max() < 1e-12


@pytest.mark.parametrize("n,tile_bits,seed", [(6, 12, 0), (10, 8, 1)])
def test_trajectory_plan_equals_oracle(built_lib, oracle, n, tile_bits, seed):
    gl = noisy_circuit(n, 40, seed)
    params = parameter_batch(gl, 1)[0]
    rng = np.random.default_rng(seed)
    codes = np.zeros(gl.n_slots, dtype=np.uint8)
    for g, (kind, *_r) in enumerate(gl.tuples()):
        if kind == 7:
            codes[gl.pidx[g]] = rng.integers(4)
        elif kind == 8:
            codes[gl.pidx[g]] = rng.integers(16)
    plan = plan_dump(gl, 2, tile_bits, 4, with_mats=True)
    check_invariants(plan, n, tile_bits, 4)
    want = oracle.state(gl, params, codes=codes)
    assert np.abs(run_plan(plan, n, params, codes=codes) - want).max() < 1e-12
    assert np.abs(run_plan_windows(plan, n, params, codes=codes) - want).max() < 1e-12
    # the noise-free plan skips the noise gates
    clean = run_plan_windows(plan_dump(gl, 0, tile_bits, 4, with_mats=True), n, params)
    assert np.abs(clean - oracle.state(gl, params)).max() < 1e-12


def test_bad_gate_is_reported(built_lib):
    gl = GateList(4)
    gl.add_cnot(1, 1)
    with pytest.raises(ValueError):
        plan_dump(gl, 0, 12, 4)
    gl = GateList(4)
    gl.add_rotation(0, 7, 0.1)
    with pytest.raises(ValueError):
        plan_dump(gl, 0, 12, 4)


@pytest.mark.parametrize("seed", range(12))
def test_lone_diagonal_gates_never_sit_on_lane_bits(built_lib, oracle, seed):
    """A diagonal one-qubit block whose qubit no other gate of the window touches (a lone RZ / Z) must not be applied as
    a per-thread phase inside a DMMA on a lane-group position: the B fragment is shared by all rows of the product.
    (Regression: [CNOT(11,9), RX(11), RZ(7)] on 12 qubits came out wrong on the GPU; the emulator now asserts it.)"""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(9, 15))
    gl = GateList(n)
    if seed == 0:
        n = 12
        gl = GateList(n)
        gl.add_cnot(11, 9)
        gl.add_rotation(0, 11, 0.3)
        gl.add_rotation(2, 7, 0.7)
    else:
        for _ in range(int(rng.integers(3, 30))):
            u = rng.random()
            if u < 0.5:
                gl.add_rotation(2, int(rng.integers(n)), float(rng.uniform(-3, 3)))      # lone RZ gates, many qubits
            elif u < 0.6:
                gl.add_pauli("Z", int(rng.integers(n)))
            elif u < 0.8:
                gl.add_rotation(int(rng.integers(2)), int(rng.integers(n)), float(rng.uniform(-3, 3)))
            else:
                c, t = rng.choice(n, size=2, replace=False)
                gl.add_cnot(int(c), int(t))
    params = parameter_batch(gl, 1)[0]
    tile_bits = int(rng.choice([9, 10, 12]))
    plan = plan_dump(gl, 0, tile_bits, 3, with_mats=True)
    init = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    want = oracle.state(gl, params, init=init)
    assert np.abs(run_plan_windows(plan, n, params, init=init) - want).max() < 1e-12


@pytest.mark.parametrize("n,gates,seed,brick", [(13, 200, 41, False), (14, 400, 42, False), (16, 300, 43, False),
                                                 (20, 440, 5, True), (18, 600, 44, False), (15, 60, 45, False),
                                                 (22, 500, 46, False), (17, 300, 47, True)])
def test_streaming_layouts_validate(built_lib, n, gates, seed, brick):
    """Streaming kernel (tq_stream.cu): every multi-tile tensor-core pass gets TMA box orders whose dims reproduce the tile's
    physical offsets, and every window's resolved slot offsets address distinct slots inside the box (validate_stream)."""
    gl = brickwork_circuit(n, 21, gates - 21 * (n - 1), seed) if brick else synthetic_circuit(n, gates, seed)
    masks = [(1 << q) | (1 << (q + 1)) for q in range(n - 1)]
    passes = plan_dump(gl, 16, 12, 3, cover_masks=masks)
    assert len(passes) > 1
    streamed = [p for p in passes if p.get("stream")]
    assert streamed, "no pass was given a streaming layout"
    for p in streamed:
        assert p["streamcheck"] == "ok", p["streamcheck"]
        for lay in p["layouts"].values():
            assert lay["box_of"][:3] == [0, 1, 2] and 1 <= lay["ops"] <= 32
            assert lay["ops"] * lay["box_bytes"] == 16 << lay["live"]
    # the expectation windows that follow gate windows read the store layout: the plan must still cover every mask
    codes = [o[0] for p in passes for w in p["windows"] for o in w["ops"]]
    assert codes.count(20) >= len(masks)   # one M_EXPC class (or more) per flip mask


def _touched(p):
    """physical qubits the gate ops of a dumped pass act on (OP_* of tq_plan.h: a / b are tile positions unless *_NL)"""
    loc, m = p["local"], 0
    for op, a, b, *_r in p["ops"]:
        if op in (0, 4):
            m |= (1 << loc[a]) | (1 << loc[b])
        elif op in (1, 2):
            m |= 1 << loc[a]
        elif op == 3:
            m |= 1 << a
        elif op == 5:
            m |= (1 << a) | (1 << loc[b])
    return m


@pytest.mark.parametrize("n,extra,seed,brick", [(20, 41, 5, True), (18, 5, 0, True), (18, 41, 1, True), (17, 60, 7, False),
                                                 (18, 200, 5, False), (16, 300, 43, False), (22, 500, 46, False)])
def test_early_expectation_light_cone(built_lib, monkeypatch, n, extra, seed, brick):
    """attach_expectation: every Hamiltonian group is evaluated exactly once; a group evaluated before the last gate pass has
    its flips local there and no later gate touches its qubits; the headline shape (20 qubits, brickwork) needs ONE
    expectation-only pass instead of two."""
    gl = brickwork_circuit(n, 21, extra, seed) if brick else synthetic_circuit(n, extra, seed)
    masks = [(1 << q) | (1 << (q + 1)) for q in range(n - 1)]
    plan = plan_dump(gl, 16, 12, 4, cover_masks=masks)
    seen = sorted(g for p in plan for g in p.get("exp_groups", []))
    assert seen == list(range(len(masks) + 1))   # (+1: the diagonal group)
    gate_passes = [i for i, p in enumerate(plan) if p["ops"]]
    last = gate_passes[-1]
    for i in gate_passes[:-1]:
        after = 0
        for j in gate_passes:
            if j > i:
                after |= _touched(plan[j])
        local = sum(1 << q for q in plan[i]["local"])
        for g in plan[i].get("exp_groups", []):
            assert g < len(masks), "the diagonal group belongs to the final state"
            assert masks[g] & ~local == 0 and masks[g] & after == 0
    n_exp_only = len(plan) - (last + 1)
    # the last gate pass leaves the state unwritten only if none of its gates touches a group evaluated after it
    if not plan[0]["last_store_needed"]:
        assert n_exp_only > 0 and len(gate_passes) >= 2
        for p in plan[last + 1:]:
            for g in p["exp_groups"]:
                assert g < len(masks) and masks[g] & _touched(plan[last]) == 0
    if (n, extra, seed, brick) == (20, 41, 5, True):
        assert not plan[0]["last_store_needed"]
    monkeypatch.setenv("TQ_EARLY_EXPECT", "0")
    late = plan_dump(gl, 16, 12, 4, cover_masks=masks)
    assert all(not p.get("exp_groups") for p in late[:last])
    assert n_exp_only <= len(late) - (last + 1)
    if (n, extra, seed, brick) == (20, 41, 5, True):
        assert (n_exp_only, len(late) - (last + 1)) == (1, 2)


def test_exchange_classes_of_a_bond_carry_the_one_coefficient_flag(built_lib):
    """A bare XX + YY coupling on a bond is an 'exchange class' (rb2 bit 1: only the 01 <-> 10 pairs carry a coefficient)
    whose eight coefficients are equal (rb2 bit 3) -- what the streaming kernel's fused chain windows key on (tq_stream.cu).
    The flip masks of a nearest-neighbour chain sit on adjacent register-bit pairs of their window."""
    n = 16
    gl = brickwork_circuit(n, 21, 20, 3)
    masks = [(1 << q) | (1 << (q + 1)) for q in range(n - 1)]
    passes = plan_dump(gl, 16, 12, 4, cover_masks=masks)
    seen = 0
    for p in passes:
        for w in p["windows"]:
            for code, rb, rb2, qsel, flags, t, fixed in w["ops"]:
                if code != 20:          # M_EXPC
                    continue
                seen += 1
                assert rb2 & 2 and rb2 & 8 and not rb2 & 1, (rb2, flags)
                assert flags in (3, 6, 12, 24), flags
    assert seen == n - 1


@pytest.mark.parametrize("n,tile_bits,low_bits,seed,gates", [(13, 10, 3, 31, 400), (14, 10, 3, 32, 500), (15, 11, 4, 33, 450),
                                                              (16, 12, 4, 34, 300)])
def test_pack_search_plans_equal_the_gate_list_and_need_no_more_passes(built_lib, oracle, monkeypatch, n, tile_bits, low_bits,
                                                                       seed, gates):
    """TQ_PACK_SEARCH=1 (opt-in): a pass's local qubits are also grown qubit (pair) by qubit (pair) for the number of blocks
    they let it execute; the plan must still be the circuit (numpy execution of the emitted windows against the oracle) and
    must not need more passes than the first-fit plan."""
    gl = synthetic_circuit(n, gates, seed)
    params = parameter_batch(gl, 1)[0]
    base = plan_dump(gl, 0, tile_bits, low_bits, with_mats=True)
    monkeypatch.setenv("TQ_PACK_SEARCH", "1")
    plan = plan_dump(gl, 0, tile_bits, low_bits, with_mats=True)
    check_invariants(plan, n, tile_bits, low_bits)
    assert sum(len(m["gates"]) for m in plan["mats"]) == len(gl)
    want = oracle.state(gl, params)
    assert np.abs(run_plan(plan, n, params) - want).max() < 1e-12
    assert np.abs(run_plan_windows(plan, n, params) - want).max() < 1e-12
    assert len(plan["passes"]) <= len(base["passes"])
