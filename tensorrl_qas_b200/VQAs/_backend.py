"""Shared machinery of the VQE_qulacs* shims: one libtqsim handle per (n_qubits, device), and identity-keyed caches
so that the COBYLA loop (up to 1000 `get_energy_qulacs` calls on the same circuit / Hamiltonian objects,
environments/environment_qulacs.py:429-445) uploads the circuit and the Hamiltonian once."""
import os
import threading

import numpy as np

from ..circuit import decode_state_tensor
from ..simulator import Simulator

_sims = {}
_ctx = threading.local()   # per-thread evaluation context, set by tensorrl_qas_b200.lockstep for its worker threads


def default_device():
    for var in ("TQSIM_DEVICE", "LOCAL_RANK"):
        if os.environ.get(var, "") != "":
            return int(os.environ[var])
    return 0


class _Slot:
    """A Simulator plus the Python objects currently bound to it (strong refs, so `is` checks are safe)."""

    def __init__(self, n, device):
        self.sim = Simulator(n, device)
        self.circuit = None
        self.circuit_version = -1
        self.op = None
        self.tn_state = None
        self.tn_bound = False


def _slot(n, device=None):
    """One handle per (n_qubits, device) -- and per lock-step worker, whose environments hold different circuits."""
    key = (int(n), default_device() if device is None else int(device), getattr(_ctx, "slot_key", None))
    if key not in _sims:
        _sims[key] = _Slot(key[0], key[1])
    return _sims[key]


def evaluate(sim, params, codes=None):
    """One energy of the problem bound to `sim`.  Directly (tq_energy_batch_host / tq_energy_traj_batch_host), or, inside
    a lock-step worker thread, through the group's coordinator: the evaluations of all workers of a round become ONE
    tq_energy_multi_host launch."""
    capture = getattr(_ctx, "capture", None)
    if capture is not None:
        # the lock-step coordinator is probing a cost closure for its (handle, angles, noise codes): nothing is evaluated
        # here, the whole round goes to the GPU as one launch afterwards
        capture.append((sim, np.array(params, dtype=np.float64).reshape(-1), codes))
        return np.float64(0.0)
    group = getattr(_ctx, "group", None)
    if group is not None:
        return group.submit(_ctx.worker, sim, params, codes)
    if codes is not None:
        return sim.energies_traj(params, codes)[0]
    return sim.energies(params)[0]


def noise_rng(module_rng):
    """The generator the noise shims sample their Pauli codes from: the module-level one, or the worker's own inside a
    lock-step thread (so that every environment keeps a reproducible stream whatever the thread interleaving)."""
    return getattr(_ctx, "rng", None) or module_rng


def release_slots(tag):
    """Close and forget the handles of one lock-step driver (slot keys (tag, worker)): its streams, pinned staging and device
    scratch go back, and so do the references to its circuits and Hamiltonians."""
    for key in [k for k in _sims if isinstance(k[2], tuple) and k[2][0] == tag]:
        _sims.pop(key).sim.close()


def reset_backends():
    """Drop all cached handles (tests)."""
    for s in _sims.values():
        s.sim.close()
    _sims.clear()


class CompiledAnsatz:
    """What `Parametric_Circuit.construct_ansatz` returns in place of a qulacs ParametricQuantumCircuit: the gate
    arrays libtqsim consumes plus the current parameter vector.  Keeps the three methods the reference calls on
    the qulacs object (VQE_qulacs.py:66,74): get_parameter_count / set_parameter / get_parameter."""

    def __init__(self, n_qubits):
        self.n_qubits = int(n_qubits)
        self.gates = None
        self.params = np.zeros(0, dtype=np.float64)
        self.version = 0

    def load_tensor(self, state, noise=None):
        self.gates = decode_state_tensor(state, self.n_qubits, noise=noise)
        self.params = np.asarray(self.gates.initial_angles, dtype=np.float64).copy()
        self.version += 1
        return self

    def get_parameter_count(self):
        return int(self.params.shape[0])

    def get_parameter(self, j):
        return float(self.params[int(j)])

    def set_parameter(self, j, value):
        self.params[int(j)] = float(value)  # exact promotion of float32 trial angles

    def get_gate_count(self):
        return 0 if self.gates is None else len(self.gates)


def bind(n_qubits, circuit, op, tn_state=None, use_tn=False):
    """Make the handle for n_qubits evaluate `circuit` against `op` (dense matrix, as the reference passes it) from
    `tn_state` (or |0..0>); re-uploads only what changed since the last call."""
    if circuit.gates is None:
        raise ValueError("construct_ansatz was not called on this circuit")
    s = _slot(n_qubits)
    if s.circuit is not circuit or s.circuit_version != circuit.version:
        s.sim.set_circuit(circuit.gates)
        s.circuit, s.circuit_version = circuit, circuit.version
    if s.op is not op:
        s.sim.set_dense_hamiltonian(np.asarray(op))
        s.op = op
    if use_tn:
        if s.tn_state is not tn_state or not s.tn_bound:
            s.sim.set_init_state(np.asarray(tn_state))
            s.tn_state, s.tn_bound = tn_state, True
    elif s.tn_bound:
        s.sim.set_init_state(None)
        s.tn_state, s.tn_bound = None, False
    return s.sim


def apply_angles(circuit, angles, which_angles):
    """The set_parameter loop of get_energy_qulacs (VQE_qulacs.py:66-74): parameter j <- angles[i] for the i-th entry j
    of which_angles (all parameters in order when which_angles is empty).  Same assignments, done as one array write
    (any float dtype is promoted exactly); like the reference it raises when angles is too short."""
    count = circuit.get_parameter_count()
    a = np.asarray(angles)
    if not list(which_angles):
        if a.shape[0] < count:
            raise IndexError(f"index {a.shape[0]} is out of bounds for angles of size {a.shape[0]}")
        circuit.params[:count] = a[:count]
        return
    for i, j in enumerate(which_angles):
        circuit.set_parameter(j, angles[i])


def sample_noise_codes(gates, rng, batch=1):
    """One Pauli code per noise gate and evaluation, with qulacs' probabilities: DepolarizingNoise -> X, Y, Z with
    p/3 each; TwoQubitDepolarizingNoise -> each of the 15 non-identity pairs with p/15 (code = pa + 4 pb)."""
    codes = np.zeros((batch, max(gates.n_slots, 1)), dtype=np.uint8)
    for g, kind in enumerate(gates.kind):
        if kind == 7 or kind == 8:
            p = gates.fixed[g]
            u = rng.random(batch)
            m = 3 if kind == 7 else 15
            hit = u < p
            if p > 0:
                codes[:, gates.pidx[g]] = np.where(hit, 1 + np.minimum((u / (p / m)).astype(np.int64), m - 1), 0)
    return codes
