"""`Simulator`: Python handle over the C ABI of libtqsim (include/tqsim.h).

Host buffers are numpy arrays; device buffers are torch CUDA tensors (torch is the allocator / stream plumbing
only -- all arithmetic happens in libtqsim's kernels).
"""
import ctypes

import numpy as np

from . import _lib
from .circuit import GateList


class TqError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libtqsim error {code}: {message}")
        self.code = code


def _dptr(a):
    return a.ctypes.data_as(_lib.c_dbl_p)


class Simulator:
    """One handle per (n_qubits, device): tq_create / tq_destroy."""

    def __init__(self, n_qubits, device=0):
        self._L = _lib.lib()
        self.n_qubits = int(n_qubits)
        self.device = int(device)
        self._h = ctypes.c_void_p()
        rc = self._L.tq_create(self.n_qubits, self.device, ctypes.byref(self._h))
        if rc != 0:
            msg = self._L.tq_last_error(None).decode()
            self._h = None
            raise TqError(rc, msg)
        self.n_params = 0
        self.n_slots = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.tq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise TqError(rc, self._L.tq_last_error(self._h).decode())

    # ------------------------------------------------------------------ problem definition
    def set_pauli_hamiltonian(self, xmask, zmask, coeff):
        x = np.ascontiguousarray(xmask, dtype=np.uint64)
        z = np.ascontiguousarray(zmask, dtype=np.uint64)
        c = np.asarray(coeff)
        cre = np.ascontiguousarray(c.real, dtype=np.float64)
        cim = np.ascontiguousarray(c.imag, dtype=np.float64) if np.iscomplexobj(c) else None
        self._check(self._L.tq_set_pauli_hamiltonian(
            self._h, len(x), x.ctypes.data_as(_lib.c_u64_p), z.ctypes.data_as(_lib.c_u64_p), _dptr(cre),
            _dptr(cim) if cim is not None else None))

    def set_dense_hamiltonian(self, H):
        H = np.ascontiguousarray(H, dtype=np.complex128)
        dim = 1 << self.n_qubits
        if H.shape != (dim, dim):
            raise ValueError(f"Hamiltonian must be {dim}x{dim}")
        self._check(self._L.tq_set_dense_hamiltonian(self._h, H.ctypes.data_as(_lib.c_dbl_p)))

    def set_init_state(self, psi=None):
        if psi is None:
            self._check(self._L.tq_set_init_state(self._h, None))
            return
        psi = np.ascontiguousarray(psi, dtype=np.complex128).reshape(-1)
        if psi.shape[0] != 1 << self.n_qubits:
            raise ValueError("initial state has the wrong dimension")
        self._check(self._L.tq_set_init_state(self._h, psi.ctypes.data_as(_lib.c_dbl_p)))

    def set_circuit(self, gl: GateList):
        kind, q0, q1, pidx, fixed = gl.arrays()
        ip = _lib.c_int_p
        self._check(self._L.tq_set_circuit(self._h, len(gl), kind.ctypes.data_as(ip), q0.ctypes.data_as(ip),
                                           q1.ctypes.data_as(ip), pidx.ctypes.data_as(ip), _dptr(fixed), gl.n_params))
        self.n_params = gl.n_params
        self.n_slots = gl.n_slots

    # ------------------------------------------------------------------ host-buffer evaluation (copies inside)
    def _params(self, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        if p.ndim == 1:
            p = p.reshape(1, -1)
        if p.shape[1] == 0:
            p = np.zeros((p.shape[0], 1), dtype=np.float64)
        return p

    def energies(self, params):
        """[B][P] host angles -> [B] energies (tq_energy_batch_host)."""
        p = self._params(params)
        out = np.empty(p.shape[0], dtype=np.float64)
        self._check(self._L.tq_energy_batch_host(self._h, p.shape[0], _dptr(p), p.shape[1], _dptr(out)))
        return out

    def energies_traj(self, params, codes):
        p = self._params(params)
        c = np.ascontiguousarray(codes, dtype=np.uint8)
        if c.ndim == 1:
            c = c.reshape(1, -1)
        if c.shape[1] == 0:
            c = np.zeros((c.shape[0], 1), dtype=np.uint8)
        if c.shape[0] != p.shape[0]:
            raise ValueError("params and codes disagree on the batch size")
        out = np.empty(p.shape[0], dtype=np.float64)
        self._check(self._L.tq_energy_traj_batch_host(self._h, p.shape[0], _dptr(p), p.shape[1],
                                                      c.ctypes.data_as(_lib.c_u8_p), c.shape[1], _dptr(out)))
        return out

    def energies_dm(self, params):
        p = self._params(params)
        out = np.empty(p.shape[0], dtype=np.float64)
        self._check(self._L.tq_energy_dm_batch_host(self._h, p.shape[0], _dptr(p), p.shape[1], _dptr(out)))
        return out

    def states(self, params):
        p = self._params(params)
        out = np.empty((p.shape[0], 1 << self.n_qubits), dtype=np.complex128)
        self._check(self._L.tq_state_batch_host(self._h, p.shape[0], _dptr(p), p.shape[1],
                                                out.ctypes.data_as(_lib.c_dbl_p)))
        return out

    def density_matrices(self, params):
        """[B][2^n (col)][2^n (row)] : entry rho[r][c] of element b is out[b, c, r]."""
        p = self._params(params)
        dim = 1 << self.n_qubits
        out = np.empty((p.shape[0], dim, dim), dtype=np.complex128)
        self._check(self._L.tq_dm_batch_host(self._h, p.shape[0], _dptr(p), p.shape[1],
                                             out.ctypes.data_as(_lib.c_dbl_p)))
        return out

    # ------------------------------------------------------------------ device-buffer evaluation (async)
    def energies_dev(self, params, out=None, stream=None, mode="pure", codes=None):
        """params: torch float64 CUDA tensor [B][ld]; returns a torch float64 CUDA tensor [B] (no sync)."""
        import torch
        if not params.is_cuda or params.dtype != torch.float64 or not params.is_contiguous():
            raise ValueError("params must be a contiguous float64 CUDA tensor")
        batch, ld = params.shape
        if out is None:
            out = torch.empty(batch, dtype=torch.float64, device=params.device)
        s = stream if stream is not None else torch.cuda.current_stream(params.device).cuda_stream
        if mode == "pure":
            rc = self._L.tq_energy_batch(self._h, batch, params.data_ptr(), ld, out.data_ptr(), s)
        elif mode == "dm":
            rc = self._L.tq_energy_dm_batch(self._h, batch, params.data_ptr(), ld, out.data_ptr(), s)
        elif mode == "traj":
            rc = self._L.tq_energy_traj_batch(self._h, batch, params.data_ptr(), ld, codes.data_ptr(),
                                              codes.shape[1], out.data_ptr(), s)
        else:
            raise ValueError(mode)
        self._check(rc)
        return out

    def evolve_states(self, states, params=None, energies=False, stream=None):
        """states: torch CUDA tensor holding [B][2^n] complex128 (dtype complex128, or float64 with a trailing 2),
        evolved IN PLACE through the handle's circuit from whatever they hold (tq_evolve_states).  params: float64 CUDA
        tensor [B][ld] or None for a circuit without parameters.  energies=True also returns <psi_b|H|psi_b> of the
        evolved states as a float64 CUDA tensor [B].  No synchronisation."""
        import torch
        if not states.is_cuda or not states.is_contiguous() or states.dtype not in (torch.complex128, torch.float64):
            raise ValueError("states must be a contiguous complex128 / float64 CUDA tensor")
        amps = states.numel() // (1 if states.dtype == torch.complex128 else 2)
        batch, rem = divmod(amps, 1 << self.n_qubits)
        if rem or batch == 0:
            raise ValueError("states does not hold a whole number of 2^n-amplitude vectors")
        ld = 0
        if params is not None:
            if not params.is_cuda or params.dtype != torch.float64 or not params.is_contiguous():
                raise ValueError("params must be a contiguous float64 CUDA tensor")
            params = params.reshape(batch, -1)
            ld = params.shape[1]
        out = torch.empty(batch, dtype=torch.float64, device=states.device) if energies else None
        s = stream if stream is not None else torch.cuda.current_stream(states.device).cuda_stream
        self._check(self._L.tq_evolve_states(self._h, batch, params.data_ptr() if params is not None else None, ld,
                                             states.data_ptr(), out.data_ptr() if energies else None, s))
        return out

    def evolve_states_exchange(self, shard, params, n_ranks, rank, recv_ptrs, stream=None):
        """tq_evolve_states_exchange: apply the circuit to this rank's shard (torch CUDA tensor, 2^n complex128) and store
        the result straight into the ranks' receive buffers (`recv_ptrs`: n_ranks device addresses) in the layout after
        the rank bits have been swapped with the top local qubits.  No synchronisation."""
        import torch
        if not shard.is_cuda or not shard.is_contiguous() or shard.dtype not in (torch.complex128, torch.float64):
            raise ValueError("shard must be a contiguous complex128 / float64 CUDA tensor")
        if shard.numel() // (1 if shard.dtype == torch.complex128 else 2) != 1 << self.n_qubits:
            raise ValueError("shard does not hold 2^n amplitudes")
        ld = 0
        if params is not None:
            params = params.reshape(1, -1)
            ld = params.shape[1]
        ptrs = (ctypes.c_uint64 * int(n_ranks))(*[int(v) for v in recv_ptrs])
        s = stream if stream is not None else torch.cuda.current_stream(shard.device).cuda_stream
        self._check(self._L.tq_evolve_states_exchange(self._h, params.data_ptr() if params is not None else None, ld,
                                                      shard.data_ptr(), int(n_ranks), int(rank), ptrs, s))

    # ------------------------------------------------------------------ introspection
    def plan_info(self, which=0):
        info = (ctypes.c_int64 * 8)()
        self._check(self._L.tq_plan_info(self._h, which, info))
        keys = ("gate_passes", "expectation_passes", "tile_bits", "launches_per_call", "groups", "nnz", "unitary_gates",
                "rotation_gates")
        return dict(zip(keys, [int(v) for v in info]))

    def plan_counts(self, which=0):
        c = (ctypes.c_int64 * 8)()
        self._check(self._L.tq_plan_counts(self._h, which, c))
        keys = ("tensor_core_blocks", "lane_register_swaps", "gate_windows", "expectation_windows",
                "fp64_pipe_windows", "direct_expectation_passes", "tensor_core_block_tiles", "stream_launches")
        return dict(zip(keys, [int(v) for v in c]))

    @property
    def launch_count(self):
        return int(self._L.tq_launch_count(self._h))

    def plan_cache_stats(self):
        """Plan cache of tq_set_circuit: {'hits', 'misses', 'same', 'entries'} (tq_plan_cache_stats)."""
        c = (ctypes.c_int64 * 4)()
        self._check(self._L.tq_plan_cache_stats(self._h, c))
        return dict(zip(("hits", "misses", "same", "entries"), [int(v) for v in c]))

    PROFILE_KINDS = ("prep_matrices_kernel", "tile_pass_kernel", "tile_pass_mma_kernel", "tile_stream_kernel<gates>",
                     "tile_stream_kernel<gates+expect>", "tile_stream_kernel<expect>", "expect_direct_kernel",
                     "reduce_partials_kernel", "dm_expect_kernel", "tile_pass_kernel<table>")

    def profile(self, on=True):
        """Bracket every kernel launch of this handle with CUDA events (tq_profile_enable)."""
        self._check(self._L.tq_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, max_records=4096):
        """Launches recorded since profile(True) / the last read: list of dicts (kernel, ms, model_bytes, alg_bytes, dmma_flops)."""
        kind = (ctypes.c_int32 * max_records)()
        ms = (ctypes.c_float * max_records)()
        mb = (ctypes.c_double * max_records)()
        ab = (ctypes.c_double * max_records)()
        n = ctypes.c_int32(0)
        self._check(self._L.tq_profile_read(self._h, max_records, kind, ms, mb, ab, ctypes.byref(n)))
        fl = (ctypes.c_double * max_records)()
        nf = ctypes.c_int32(0)
        self._check(self._L.tq_profile_read_flops(self._h, max_records, fl, ctypes.byref(nf)))
        return [{"kernel": self.PROFILE_KINDS[kind[i]], "ms": float(ms[i]), "model_bytes": float(mb[i]),
                 "alg_bytes": float(ab[i]), "dmma_flops": float(fl[i]) if i < nf.value else 0.0} for i in range(n.value)]


def energies_multi(sims, params, codes=None):
    """One launch for len(sims) DIFFERENT problems (tq_energy_multi_host): sims[i] is a Simulator with its own circuit /
    Hamiltonian / initial state (same n_qubits, same device, n_qubits <= 12), params[i] its angle vector (any float
    dtype, promoted exactly), codes[i] an optional trajectory-noise code row.  Returns a float64 array of energies."""
    n = len(sims)
    if n == 0:
        return np.zeros(0)
    L = sims[0]._L
    handles = (ctypes.c_void_p * n)(*[s._h for s in sims])
    rows = [np.ascontiguousarray(np.asarray(p, dtype=np.float64).reshape(-1)) for p in params]
    for s, r in zip(sims, rows):
        if r.shape[0] < s.n_params:
            raise ValueError("angle vector shorter than the circuit's parameter count")
    prow = (_lib.c_dbl_p * n)(*[r.ctypes.data_as(_lib.c_dbl_p) if r.shape[0] else None for r in rows])
    crow = None
    keep = []
    if codes is not None:
        keep = [None if c is None else np.ascontiguousarray(np.asarray(c, dtype=np.uint8).reshape(-1)) for c in codes]
        crow = (_lib.c_u8_p * n)(*[None if c is None or c.shape[0] == 0 else c.ctypes.data_as(_lib.c_u8_p) for c in keep])
    out = np.empty(n, dtype=np.float64)
    rc = L.tq_energy_multi_host(n, handles, prow, crow, _dptr(out))
    if rc != 0:
        raise TqError(rc, L.tq_last_error(sims[0]._h).decode())
    return out


def fp64_peak(device=0, which=0):
    """Measured FP64 peak in TFLOP/s: which = 0 DMMA (mma.sync.m8n8k4.f64), 1 DFMA (tq_fp64_peak)."""
    out = ctypes.c_double(0.0)
    rc = _lib.lib().tq_fp64_peak(int(device), int(which), ctypes.byref(out))
    if rc != 0:
        raise TqError(rc, "tq_fp64_peak failed")
    return float(out.value)


def plan_dump(gl: GateList, which=0, tile_bits=12, low_bits=4, cover_masks=(), with_mats=False):
    """Planner dry run (no GPU): list of passes, each {'lead', 'local', 'ops': [(op, a, b, t, flags, fixed)],
    'windows': [{'wpos', 'tpos', 'ops': [(code, rb, rb2, qsel, flags, t, fixed)]}]}.  with_mats=True also returns the
    fused blocks' matrix programs: {'passes': [...], 'mats': [{'nq', 'diag', 'gates': [(kind, lq, pidx, fixed)]}]}."""
    L = _lib.lib()
    kind, q0, q1, pidx, fixed = gl.arrays()
    cm = np.ascontiguousarray(cover_masks, dtype=np.uint64)
    ip = _lib.c_int_p
    ptr = L.tq_plan_dump(gl.n_qubits, len(gl), kind.ctypes.data_as(ip), q0.ctypes.data_as(ip), q1.ctypes.data_as(ip),
                         pidx.ctypes.data_as(ip), fixed.ctypes.data_as(_lib.c_dbl_p), which, tile_bits, low_bits,
                         len(cm), cm.ctypes.data_as(_lib.c_u64_p))
    text = ctypes.string_at(ptr).decode()
    L.tq_free(ptr)
    passes = []
    mats = []
    last_store = True
    for line in text.splitlines():
        tok = line.split()
        if tok[0] == "ERROR":
            raise ValueError(line[6:])
        if tok[0] == "LASTSTORE":   # (which | 16) does the last gate pass write the state back for the expectation-only passes
            last_store = tok[1] == "1"
        elif tok[0] == "MAT":
            mats.append({"nq": int(tok[1]), "diag": int(tok[2]), "gates": []})
        elif tok[0] == "MG":  # (kind, lq, pidx, fixed)
            mats[-1]["gates"].append((int(tok[1]), int(tok[2]), int(tok[3]), float(tok[4])))
        elif tok[0] == "PASS":
            lead = int(tok[1].split("=")[1])
            local = [int(v) for v in tok[2].split("=")[1].split(",") if v != ""]
            support = int(tok[3].split("=")[1]) if len(tok) > 3 else (1 << 64) - 1
            passes.append({"lead": lead, "local": local, "support": support, "ops": [], "windows": []})
        elif tok[0] == "EXPGROUPS":   # (which | 16) Hamiltonian groups evaluated in this pass: indices into cover_masks,
            passes[-1]["exp_groups"] = [int(v) for v in tok[1:]]   # len(cover_masks) = the diagonal group
        elif tok[0] == "STREAMABLE":
            passes[-1]["stream"] = tok[1] == "1"
        elif tok[0] == "STREAM":
            kv = dict(t.split("=") for t in tok[2:])
            passes[-1].setdefault("layouts", {})[tok[1]] = {
                "live": int(kv["live"]), "ops": int(kv["ops"]), "box_bytes": int(kv["box_bytes"]),
                "box_of": [int(v) for v in kv["box_of"].split(",")],
                "dims": [tuple(int(x) for x in d.split(":")) for d in kv["dims"].split(",") if d]}
        elif tok[0] == "STREAMCHECK":
            passes[-1]["streamcheck"] = " ".join(tok[1:])
        elif tok[0] == "OP":
            passes[-1]["ops"].append((int(tok[1]), int(tok[2]), int(tok[3]), int(tok[4]), int(tok[5]), float(tok[6])))
        elif tok[0] == "WIN":
            wpos = [int(v) for v in tok[1].split("=")[1].split(",") if v != ""]
            tpos = [int(v) for v in tok[2].split("=")[1].split(",") if v != ""]
            passes[-1]["windows"].append({"wpos": wpos, "tpos": tpos, "ops": []})
        elif tok[0] == "MWIN":  # tensor-core window: r, ql, g, w (entry layout), rout, qlout (exit layout), flags
            kv = dict(t.split("=") for t in tok[1:])
            ints = lambda v: [int(x) for x in v.split(",") if x != ""]  # noqa: E731
            passes[-1]["windows"].append({"mma": True, "r": ints(kv["r"]), "ql": int(kv["ql"]), "g": ints(kv["g"]),
                                          "w": ints(kv["w"]), "rout": ints(kv["rout"]), "qlout": int(kv["qlout"]),
                                          "flags": int(kv["flags"]), "dead": int(kv.get("dead", 0)), "ops": []})
        elif tok[0] == "WOP":  # (code, rb, rb2, qsel, flags, t, fixed)
            passes[-1]["windows"][-1]["ops"].append(tuple(int(v) for v in tok[1:7]) + (float(tok[7]),))
    for p in passes:
        p["last_store_needed"] = last_store
    return {"passes": passes, "mats": mats} if with_mats else passes
