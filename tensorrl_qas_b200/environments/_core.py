"""CircuitEnv of the TensorRL-QAS environments, backed by libtqsim.

One implementation serves the reference's five near-identical environment modules (SURVEY.md section 8a, rows
a6-a11); each public module (`environment_qulacs*.py`) derives `CircuitEnv` from `CircuitEnvBase` and only sets
the class attributes that distinguish it:

    vc               the VQE_qulacs* shim it evaluates energies with
    tn_in_agent      True : the MPS init circuit is written INTO the state tensor, so the agent sees it and COBYLA
                            optimises its angles (environments/environment_qulacs.py:285-328, agent gates are placed
                            after it, :205-207);
                     False: the MPS circuit is simulated once into `TN_state`, the Hamiltonian is bit-reversed
                            (environments/environment_qulacs_TN_notin_agent.py:158-163,377) and the tensor holds
                            agent gates only (:268-270)
    shot_args        the noise variants pass (n_shots, weights) on to the shim
    restricted       hexagon action dictionary (environment_qulacs_TN_notin_agent_noise_restricted.py:139,545)

Behaviour kept from the reference, because it shapes trajectories: the COBYLA call optimises
the circuit BEFORE the new gate and the new rotation enters at angle 0 (Q1); angles live in a float32 tensor (Q2);
`get_energy` returns the same value twice (Q5); `illegal_action_new` is stateful and is also called by the driver
(Q9); data paths are relative to the working directory (Q11; `TQ_DATA_ROOT` may point elsewhere).
Deviations, all result-neutral for the shipped cfgs: qubit indices come from the QPY record instead of
`str(qargs)` (Q8), the npz file is read once and cached instead of on every reset (Q7), and an empty parameter
vector is not handed to scipy (Q19: under scipy >= 1.16 COBYLA raises on it) -- the energy is evaluated once.
"""
import copy
import os
from sys import stdout

import numpy as np
import scipy.optimize
import torch

from .. import loaders
from ..simulator import Simulator
from .utils import curricula, utils, utils_topology_restrict

_LATTICE_MODELS = ("heisenberg", "tfim_j1_h0.05")   # file names without geometry / mapping (environment_qulacs.py:76,101)
_AXIS_ROW = {"rx": 0, "ry": 1, "rz": 2}
_VERBOSE = bool(int(os.environ.get("TQ_ENV_VERBOSE", "0")))


def _data_root():
    return os.environ.get("TQ_DATA_ROOT", "")


def _say(*args):
    if _VERBOSE:
        print(*args)


class CircuitEnvBase:
    vc = None
    tn_in_agent = True
    shot_args = False
    restricted = False

    # ------------------------------------------------------------------------------------------ construction ----
    def __init__(self, conf, device):
        env, problem = conf["env"], conf["problem"]
        self.device = device
        n = self.num_qubits = env["num_qubits"]
        self.num_layers = env["num_layers"]
        self.random_halt = int(env["rand_halt"])
        self.TN_init = env["tn_init"]
        self.n_shots = int(env["n_shots"]) if self.shot_args else env["n_shots"]
        self.ham_type = problem["ham_type"]
        self.ham_mapping = self.ham_model = problem["mapping"]
        self.geometry = problem["geometry"].replace(" ", "_")
        self.zero_param_init = int(env["zero_param_init"])

        # cfg noise keys are parsed like the reference does but never reach the simulator: the VQA modules hard-code
        # their strengths
        raw = env["noise_values"]
        if raw != 0:
            comma = raw.index(",")
            self.noise_values = [float(raw[1:comma]), float(raw[comma + 1:-1])]
        else:
            self.noise_values = []
        self.noise_models = ["depolarizing", "two_depolarizing", "amplitude_damping"][:len(self.noise_values)]
        self.phys_noise = len(self.noise_models) > 0
        self.err_mitig = env["err_mitig"]
        self.fake_min_energy = env.get("fake_min_energy")
        self.fn_type = env["fn_type"]
        self.cnot_rwd_weight = env.get("cnot_rwd_weight", 1.)

        # ---- MPS init circuit (a13): QPY -> ASAP layers; the episode length shrinks by its depth ----
        self.TN_bond = int(env["tn_bond"])
        if self.TN_bond:
            self.tenor_circ = loaders.load_qpy_circuit(self._artefact("init_state_circ", "init_", f"_TNbond{self.TN_bond}.qpy"))
            self.depth_wise_gates = self.tenor_circ.layers()
            self._tn_depth = len(self.depth_wise_gates)
            self.num_layers_termination = self.num_layers - self._tn_depth
            _say("THE DEPTH OF TENSOE CIRCUIT:", self._tn_depth)
        else:
            self._tn_depth = 0
            self.num_layers_termination = self.num_layers

        self.noise_flag = True
        self.state_with_angles = conf["agent"]["angles"]
        self.current_number_of_cnots = 0

        # ---- Hamiltonian (a12) ----
        self._ham_cache = None
        eigvals = self._load_hamiltonian()
        if not self.tn_in_agent:
            # a9: |TN> = U_mps |0..0> in qiskit's (= qulacs') little-endian order, simulated once on the device
            self.TN_state = self._simulate_init_circuit()
            if _VERBOSE:
                psi = np.asmatrix(self.TN_state)
                print("Initial energy:", (psi @ self.hamiltonian) @ psi.getH())

        min_eig = env["fake_min_energy"] if "fake_min_energy" in env else min(eigvals)
        self.min_eig = self.fake_min_energy if self.fake_min_energy is not None else min(eigvals)
        self.max_eig = max(eigvals)
        self.curriculum_dict = {
            self.ham_type: curricula.__dict__[env["curriculum_type"]](env, target_energy=min_eig)}

        self.done_threshold = env["accept_err"]
        stdout.flush()
        self.state_size = self.num_layers * n * (n + 3 + 3)
        self.step_counter = -1
        self.prev_energy = None
        self.moments = [0] * n
        self.illegal_actions = [[]] * n
        self.energy = 0
        self.opt_ang_save = 0
        self.action_size = len(self._action_table()) if self.restricted else n * (n + 2)
        self.previous_action = [0, 0, 0, 0]
        self.save_circ = 0

        if "non_local_opt" in conf:
            opt = conf["non_local_opt"]
            self.global_iters = opt["global_iters"]
            self.optim_method = opt["method"]
            self.optim_alg = opt["optim_alg"]
            if "a" in opt:
                self.options = {k: opt[k] for k in ("a", "alpha", "c", "gamma", "beta_1", "beta_2")}
            if "lamda" in opt:
                self.options["lamda"] = opt["lamda"]
            if "maxfev" in opt:
                self.maxfev = {"maxfev": int(opt["maxfev"])}
            if "maxfev1" in opt:
                self.maxfevs = {k: int(opt[k]) for k in ("maxfev1", "maxfev2", "maxfev3")}
        else:
            self.global_iters = 0
            self.optim_method = None

    def _artefact(self, folder, prefix, suffix):
        """dmrg-to-qc/<folder>/<prefix><problem><suffix>, named as the reference names it
        (environment_qulacs.py:75-82,100-103)."""
        if self.ham_type not in _LATTICE_MODELS:
            stem = f"{self.ham_type}_{self.num_qubits}q_geom_{self.geometry}_{self.ham_mapping}"
        else:
            stem = f"{self.ham_type}_{self.num_qubits}q"
        return os.path.join(_data_root(), "dmrg-to-qc", folder, prefix + stem + suffix)

    def _load_hamiltonian(self):
        """Sets self.hamiltonian (+ self.weights) from the npz and returns eigvals.  The reference reloads the file
        in every reset() (environment_qulacs.py:347-351); the content is cached here, and the SAME matrix object is
        handed out each time so the shim's identity-keyed upload cache stays warm."""
        if self._ham_cache is None:
            d = np.load(self._artefact("mol_data", "", ".npz"))
            H = d["hamiltonian"]
            if not self.tn_in_agent:
                # Operator(H).reverse_qargs().to_matrix(): rows and columns bit-reversed, complex128
                H = loaders.reverse_qargs(np.asarray(H, dtype=np.complex128))
            self._ham_cache = (H, d["eigvals"], d["weights"])
        self.hamiltonian, eigvals, self.weights = self._ham_cache
        return eigvals

    def _simulate_init_circuit(self):
        """Statevector(tenor_circ).data (environment_qulacs_TN_notin_agent.py:158) through libtqsim."""
        gates = loaders.init_circuit_gatelist(self.tenor_circ, parametric=False)
        sim = Simulator(self.num_qubits, _device_index(self.device))
        try:
            sim.set_circuit(gates)
            return sim.states(np.zeros((1, 1)))[0]
        finally:
            sim.close()

    def _action_table(self):
        if self.restricted:
            return utils_topology_restrict.dictionary_of_actions_hexagon_connectivity_reverted(self.num_qubits)
        return utils.dictionary_of_actions(self.num_qubits)

    # ------------------------------------------------------------------------------------------------- step -----
    def step(self, action, train_flag=True):
        """Places the action's gate in the first free layer of its qubit(s), re-optimises the angles of the circuit
        as it was BEFORE this gate, then scores the new circuit (environment_qulacs.py:169-267)."""
        n = self.num_qubits
        next_state = self.state.clone()
        self.step_counter += 1
        first_layer = self._tn_depth if (self.tn_in_agent and self.TN_init) else 0

        ctrl, targ = action[0], (action[0] + action[1]) % n
        rot_qubit, rot_axis = action[2], action[3]
        self.action = action
        has_rot, has_cnot = rot_qubit < n, ctrl < n

        if has_rot:
            slot = self.moments[rot_qubit]
        elif has_cnot:
            slot = max(self.moments[ctrl], self.moments[targ])
        if has_cnot:
            next_state[first_layer + slot][targ][ctrl] = 1
        elif has_rot:
            next_state[first_layer + slot][n + rot_axis - 1][rot_qubit] = 1
        if has_rot:
            self.moments[rot_qubit] += 1
        elif has_cnot:
            self.moments[ctrl] = self.moments[targ] = max(self.moments[ctrl], self.moments[targ]) + 1

        self.current_action = action
        self.illegal_action_new()

        if self.optim_method in ["scipy_each_step"]:
            thetas, nfev, opt_ang = self.scipy_optim(self.optim_alg)
            next_state[:, n + 3:n + 6, :] = thetas[:, 0:3, :]
        self.opt_ang_save = opt_ang          # NameError for any other optim_method, as in the reference (Q3)
        self.state = next_state.clone()

        energy, energy_noiseless = self.get_energy()
        if self.noise_flag == False:  # noqa: E712
            energy = energy_noiseless
        self.energy = energy
        if energy < self.curriculum.lowest_energy and train_flag:
            self.curriculum.lowest_energy = copy.copy(energy)

        self.error = float(abs(self.min_eig - energy))
        self.error_noiseless = float(abs(self.min_eig - energy_noiseless))
        rwd = self.reward_fn(energy)
        self.prev_energy = np.copy(energy)
        self.rwd = rwd

        energy_done = int(self.error < self.done_threshold)
        layers_done = self.step_counter == (self.num_layers_termination - 1)
        done = int(energy_done or layers_done)

        self.previous_action = copy.deepcopy(action)
        self.nfev = nfev
        self.save_circ = 0

        if self.random_halt and self.step_counter == self.halting_step:
            done = 1
        if done:
            self.curriculum.update_threshold(energy_done=energy_done)
            self.done_threshold = self.curriculum.get_current_threshold()
            self.curriculum_dict[self.current_prob] = copy.deepcopy(self.curriculum)

        reward = torch.tensor(rwd, dtype=torch.float32, device=self.device)
        if self.state_with_angles:
            return next_state.view(-1).to(self.device), reward, done
        return next_state[:, :n + 3].reshape(-1).to(self.device), reward, done

    # ------------------------------------------------------------------------------------------------ reset -----
    def reset(self):
        """Fresh (num_layers, n+6, n) float32 encoding; rows [0,n) CNOT one-hots [targ][ctrl], rows [n,n+3) rotation
        one-hots [axis][qubit], rows [n+3,n+6) angles (environment_qulacs.py:269-362)."""
        n = self.num_qubits
        state = torch.zeros((self.num_layers, n + 3 + 3, n))
        self.state = state
        if self.tn_in_agent and self.TN_init:
            self._encode_init_circuit(state)

        if self.random_halt:
            self.halting_step = np.clip(np.random.negative_binomial(n=70, p=0.573, size=1), 25, 70)[0]

        self.current_number_of_cnots = 0
        self.current_action = [n] * 4
        self.illegal_actions = [[]] * n
        thetas = state[:, n + 3:]
        self.make_circuit(thetas)
        self.step_counter = -1
        self.moments = [0] * n
        self.current_prob = self.ham_type
        self.curriculum = copy.deepcopy(self.curriculum_dict[self.current_prob])
        self.done_threshold = copy.deepcopy(self.curriculum.get_current_threshold())
        eigvals = self._load_hamiltonian()
        self.min_eig = self.fake_min_energy if self.fake_min_energy is not None else min(eigvals)
        self.prev_energy = self.get_energy(thetas)[1]
        _say("Very first energy:", self.prev_energy)

        if self.state_with_angles:
            return state.reshape(-1).to(self.device)
        return state[:, :n + 3].reshape(-1).to(self.device)

    def _encode_init_circuit(self, state):
        """a10: ASAP layer d of the MPS circuit -> tensor slice d; qiskit qubit p -> column n-1-p; angle -> -theta
        (qulacs' rotation sign), or 0 when zero_param_init (StructureRL); cx(c, t) -> [n-1-t][n-1-c] = 1
        (environment_qulacs.py:285-328).  float32 on assignment, like the reference's default-dtype tensor."""
        n = self.num_qubits
        for depth_no, layer in enumerate(self.depth_wise_gates):
            for name, qubits, angle in layer:
                if name == "cx":
                    state[depth_no][n - 1 - qubits[1]][n - 1 - qubits[0]] = 1
                    continue
                col, row = n - 1 - qubits[0], _AXIS_ROW[name]
                state[depth_no][n + row][col] = 1
                state[depth_no][n + 3 + row][col] = 0 if self.zero_param_init else -angle

    def make_circuit(self, thetas=None):
        """The reference builds a non-parametric qulacs circuit here and its only caller drops it
        (environment_qulacs.py:338,364-404).  Kept as a cheap gate listing for API parity."""
        from ..circuit import decode_state_tensor
        state = self.state.clone()
        if thetas is not None:
            state[:, self.num_qubits + 3:] = thetas
        return decode_state_tensor(state, self.num_qubits)

    # ---------------------------------------------------------------------------------------- energy / VQE ------
    def _exp_val_args(self):
        args = [self.hamiltonian]
        if not self.tn_in_agent:
            args.append(self.TN_state)
        if self.shot_args:
            args += [self.n_shots, self.weights]
        return args

    def get_energy(self, thetas=None):
        """a6: circuit rebuilt from the float32 tensor, one evaluation; `thetas` is ignored and the value is returned
        twice, as in the reference (environment_qulacs.py:407-415)."""
        inst = self.vc.Parametric_Circuit(n_qubits=self.num_qubits, noise_models=self.noise_models,
                                          noise_values=self.noise_values)
        circ = inst.construct_ansatz(self.state)
        energy = self.vc.get_exp_val(self.num_qubits, circ, *self._exp_val_args())
        return energy, energy

    def scipy_optim(self, method, which_angles=[]):
        """a7: COBYLA (cfg `optim_alg`) over the angles of the CURRENT circuit, maxiter = cfg `global_iters`; the
        float32 angles are the start point and the result is written back as float32
        (environment_qulacs.py:417-445)."""
        n = self.num_qubits
        state = self.state.clone()
        thetas = state[:, n + 3:]
        rot_pos = (state[:, n:n + 3] == 1).nonzero(as_tuple=True)
        angles = thetas[rot_pos]

        inst = self.vc.Parametric_Circuit(n_qubits=n, noise_models=self.noise_models, noise_values=self.noise_values)
        circuit = inst.construct_ansatz(state)
        x0 = np.asarray(angles.cpu().detach())

        kwargs = dict(observable=self.hamiltonian, circuit=circuit, n_qubits=n, n_shots=int(self.n_shots),
                      phys_noise=self.phys_noise, which_angles=[])
        if not self.tn_in_agent:
            kwargs["TN_state"] = self.TN_state
        if self.shot_args:
            kwargs["weights"] = self.weights

        def cost(x):
            return self.vc.get_energy_qulacs(x, **kwargs)

        picked = list(which_angles)
        start = x0[which_angles] if picked else x0
        if start.shape[0] == 0:
            # Q19: nothing to optimise yet (the fixed environments reach this on their first steps)
            cost(start)
            x, nfev = np.zeros(0, dtype=np.float64), 1
        else:
            if getattr(self, "optimizer", None) is None:
                self.optimizer = os.environ.get("TQ_OPTIMIZER", "scipy")
            if self.optimizer == "native" and method == "COBYLA":
                # optional: the library's own ask/tell COBYLA (microseconds per iteration instead of scipy >= 1.16's
                # pure-Python milliseconds).  Not scipy's trajectory: off unless asked for.
                from .. import cobyla
                from ..VQAs import _backend
                group = getattr(_backend._ctx, "group", None)
                raw_energy = not (self.shot_args and int(self.n_shots) != 0)   # (shot noise is added after the evaluation)
                if group is not None and hasattr(group, "optimise") and raw_energy:
                    # lock-step worker: hand the whole problem to the group -- the optimisers of all environments then
                    # run in ONE host loop, one launch per round (no thread hand-over per evaluation)
                    res = group.optimise(_backend._ctx.worker, cost, start, self.global_iters)
                else:
                    res = cobyla.minimize(cost, start, maxiter=self.global_iters)
            else:
                res = scipy.optimize.minimize(cost, x0=start, method=method, options={"maxiter": self.global_iters})
            x, nfev = res["x"], res["nfev"]
        if picked:
            x0[which_angles] = x
            thetas[rot_pos] = torch.tensor(x0, dtype=torch.float)
        else:
            thetas[rot_pos] = torch.tensor(x, dtype=torch.float)
        return thetas, nfev, x

    def reward_fn(self, energy):
        """a8 (environment_qulacs.py:447-463)"""
        if self.fn_type == "incremental_with_fixed_ends":
            if self.error < self.done_threshold:
                return 5.
            if self.step_counter == (self.num_layers_termination - 1):
                return -5.
            return np.clip((self.prev_energy - energy) / abs(self.prev_energy - self.min_eig), -1, 1)
        print("Please define your own reward function!")

    # -------------------------------------------------------------------------------------- illegal actions -----
    def illegal_action_new(self):
        """Stateful action mask (environment_qulacs.py:466-591).  `self.illegal_actions` holds up to n remembered
        actions; the current action evicts remembered ones it conflicts with and is then remembered itself.  The
        bookkeeping below is order-sensitive in the same way as the reference's (slots are re-read while they are
        being rewritten), because the resulting mask feeds the agent and therefore the trajectory."""
        n = self.num_qubits
        action = self.current_action
        slots = self.illegal_actions
        ctrl, targ = action[0], (action[0] + action[1]) % n
        rot_qubit, rot_axis = action[2], action[3]

        def remember():
            for i in range(1, n):
                if len(slots[i]) == 0:
                    slots[i] = action
                    break

        def sweep(verdict):
            """verdict(old) -> 'evict' (drop old, remember the action), 'keep' (remember the action) or None"""
            if sum(sum(s) for s in slots) == 0:
                slots[0] = action
                return
            for i in range(n):
                old = slots[i]          # re-read: earlier iterations may have rewritten this slot
                if len(old) == 0:
                    continue
                what = verdict(old)
                if what == "evict":
                    slots[i] = []
                if what is not None:
                    remember()

        if ctrl < n:
            def cnot_verdict(old):
                if old[2] == n:      # remembered CNOT: conflict when it shares a qubit with the new one
                    ends = (old[0], (old[0] + old[1]) % n)
                    return "evict" if (ctrl in ends or targ in ends) else "keep"
                return "evict" if old[2] in (ctrl, targ) else "keep"   # remembered rotation under the new CNOT
            sweep(cnot_verdict)

        if rot_qubit < n:
            def rot_verdict(old):
                if old[0] == n:      # remembered rotation
                    if rot_qubit == old[2]:
                        return "evict" if rot_axis != old[3] else None
                    return "keep"
                ends = (old[0], (old[0] + old[1]) % n)
                return "evict" if rot_qubit in ends else "keep"
            sweep(rot_verdict)

        # duplicates: the first later twin of slot i decides which of the two is cleared
        for i in range(n):
            for j in range(i + 1, n):
                if slots[i] == slots[j]:
                    slots[j if j == i + 1 else i] = []
                    break
        # close gaps, one position per sweep
        for i in range(n - 1):
            if len(slots[i]) == 0:
                slots[i], slots[i + 1] = slots[i + 1], []

        decoded = [key for key, act in self._action_table().items() for s in slots if s == act]
        self.illegal_actions = slots
        return decoded


def _device_index(device):
    """CUDA ordinal libtqsim should run on: the env's torch device when it is a CUDA device, else the shim default."""
    from ..VQAs._backend import default_device
    try:
        dev = torch.device(device)
    except (TypeError, RuntimeError):
        return default_device()
    if dev.type == "cuda" and dev.index is not None:
        return dev.index
    return default_device()
