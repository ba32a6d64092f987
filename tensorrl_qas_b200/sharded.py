"""Single-state sharding: ONE n-qubit state vector spread over R = 2^g ranks (the north star's
">= 28 qubits" case).  The reference has no code for this (its README only claims the scale, README.md:29); the
semantics are those of the single-GPU path -- the same circuit (environments/VQAs/VQE_qulacs.py:12-44), the same
Pauli-sum energy (VQE_qulacs.py:79-86) -- and the tests pin it against the CPU checker and the unsharded kernels.

Layout.  A *layout* maps logical qubit q to a physical position pos[q].  Positions [0, n-g) index the amplitudes inside
a rank's shard (2^(n-g) complex128, contiguous in HBM); positions [n-g, n) are the bits of the rank number.  Every rank
drives its shard with the ordinary single-GPU engine on n-g qubits (`tq_evolve_states`, include/tqsim.h: the fused
tile-pass kernels applied in place to a device-resident state).

Schedule (host logic, identical on all ranks, `plan_state_sharding`):
  * a gate whose qubits are all local is appended to the current *segment*;
  * a CNOT whose control is a rank bit and whose target is local needs no communication: ranks with the bit set apply X
    to the target, the others skip it;
  * anything else closes the segment and exchanges the g rank bits with the top g local positions: that is exactly one
    all-to-all of contiguous chunks (chunk c of rank r goes to rank c as its chunk r).  Which logical qubits go out is
    chosen by furthest next use (Belady); they are first moved to the top g local positions by SWAPs written as three
    CNOTs and appended to the segment, where the engine's gate fusion absorbs them into its tile passes -- no separate
    permutation pass over the state;
  * the Hamiltonian's flip-mask groups are evaluated where their X/Y qubits are local (Z factors on rank bits are a sign
    per rank folded into the coefficients); groups that are not local in the final layout get further exchanges.  The
    first evaluation is fused with the last segment (energy and evolution in one call); the partial energies of all ranks
    and evaluations are added with one all-reduce.

Fused exchange.  With the GPU engine the all-to-all is not a separate pass: the last tile pass of a segment stores every
amplitude straight into the receive buffer of the rank that owns it after the exchange (`tq_evolve_states_exchange`: peer
memory mapped over CUDA IPC, 16-byte stores over NVLink issued tile by tile under the tensor-core work of the same
kernel); the ranks then meet in a one-element all-reduce.  Every rank keeps two shard buffers and alternates between
them.  `fused_exchange=False` keeps the NCCL all-to-all (the baseline the fused path is measured against, and the
default on eight ranks, where it is the faster of the two).

`comm` is either `TorchComm` (one rank per process, torch.distributed: NCCL on GPUs, gloo in the CPU tests) or
`LocalComm` (R virtual ranks inside one process: the same schedule with the all-to-all done as a tensor transpose --
used to test the whole path on a single GPU).  The engine is libtqsim (`GpuEngine`); there is no CPU engine in the
product -- the CPU tests plug in a numpy stand-in of their own.
"""
import ctypes

import numpy as np

from .circuit import KIND, GateList

_CNOT = KIND["CNOT"]


# ---------------------------------------------------------------------------------------------------------------
# schedule
# ---------------------------------------------------------------------------------------------------------------
def _bits(mask):
    mask = int(mask)
    return [q for q in range(mask.bit_length()) if (mask >> q) & 1]


def plan_state_sharding(gates, n_qubits, g, flip_masks):
    """gates: [(kind, q0, q1, param_idx, fixed)] on logical qubits (pure-state kinds only).  flip_masks: the distinct
    X/Y masks of the Hamiltonian's term groups.  Returns a list of steps:
        ("evolve", [(kind, p0, p1, param_idx, fixed)])   physical positions; a CNOT may have p0 >= n - g (rank bit)
        ("exchange",)                                     rank bits <-> top g local positions
        ("expect", [group indices], pos)                  pos[q] = physical position of logical qubit q at that point
    """
    n, nl = int(n_qubits), int(n_qubits) - int(g)
    if g < 0 or nl < 2 * g or nl < g + 2:
        raise ValueError("too few local qubits for this many ranks")
    gates = [tuple(t) for t in gates]
    for kind, q0, q1, _, _ in gates:
        if kind > KIND["Z"]:
            raise ValueError("state sharding supports the pure-state gate kinds only")
        if not (0 <= q0 < n) or (kind == _CNOT and (not (0 <= q1 < n) or q1 == q0)):
            raise ValueError("gate qubit out of range")
    flip_masks = [int(m) for m in flip_masks]
    if any(m >> n for m in flip_masks):
        raise ValueError("Hamiltonian acts outside the register")
    n_gates = len(gates)
    in_flips = [sum(1 for m in flip_masks if (m >> q) & 1) for q in range(n)]
    uses = [[] for _ in range(n)]
    for i, (kind, q0, q1, _, _) in enumerate(gates):
        uses[q0].append(i)
        if kind == _CNOT:
            uses[q1].append(i)
    cursor = [0] * n   # index into uses[q] of the first use not yet executed

    pos = list(range(n))
    steps, cur = [], []
    top = list(range(nl - g, nl))

    def exchange(chosen):
        """move the chosen (currently local) logical qubits to the top g local positions, then swap with the rank bits"""
        at = {pos[q]: q for q in range(n)}
        free = [t for t in top if at[t] not in chosen]
        for q in chosen:
            if pos[q] in top:
                continue
            t = free.pop()
            o, p = at[t], pos[q]
            cur.extend([(_CNOT, p, t, -1, 0.0), (_CNOT, t, p, -1, 0.0), (_CNOT, p, t, -1, 0.0)])
            pos[q], pos[o] = t, p
            at[t], at[p] = q, o
        if cur:
            steps.append(("evolve", list(cur)))
            cur.clear()
        steps.append(("exchange",))
        for q in range(n):
            if pos[q] >= nl:
                pos[q] -= g
            elif pos[q] >= nl - g:
                pos[q] += g

    def pick(exclude, cost):
        """g currently-local logical qubits outside `exclude` with the largest cost; ties: already on top, then higher"""
        cand = [q for q in range(n) if pos[q] < nl and q not in exclude]
        cand.sort(key=lambda q: (cost(q), pos[q] >= nl - g, pos[q]), reverse=True)
        return cand[:g]

    for i, (kind, q0, q1, pidx, fixed) in enumerate(gates):
        qs = (q0, q1) if kind == _CNOT else (q0,)
        local = all(pos[q] < nl for q in qs)
        if not local and not (kind == _CNOT and pos[q1] < nl):
            def next_use(q):
                u = uses[q]
                return u[cursor[q]] if cursor[q] < len(u) else n_gates + (0 if in_flips[q] else 1)
            exchange(pick(set(qs), next_use))
        cur.append((kind, pos[q0], pos[q1] if kind == _CNOT else 0, pidx, fixed))
        for q in qs:
            cursor[q] += 1

    remaining = list(range(len(flip_masks)))
    while True:
        now = [gi for gi in remaining if all(pos[q] < nl for q in _bits(flip_masks[gi]))]
        if now:
            if cur:
                steps.append(("evolve", list(cur)))
                cur.clear()
            steps.append(("expect", now, list(pos)))
            remaining = [gi for gi in remaining if gi not in now]
        if not remaining:
            break
        load = [sum(1 for gi in remaining if (flip_masks[gi] >> q) & 1) for q in range(n)]
        chosen = pick(set(), lambda q: -load[q])
        if not any(all(q not in chosen for q in _bits(flip_masks[gi])) for gi in remaining):
            raise ValueError("a Hamiltonian term flips more qubits than a shard holds")
        exchange(chosen)
    if cur:   # gates after which nothing is evaluated (no Hamiltonian): still part of the evolution
        steps.append(("evolve", list(cur)))
    return steps


def rank_gatelist(ops, n_local, rank, n_params):
    """The segment as rank `rank` runs it: CNOTs controlled by a rank bit become X (bit set) or nothing."""
    gl = GateList(n_local)
    for kind, p0, p1, pidx, fixed in ops:
        if kind == _CNOT and p0 >= n_local:
            if (rank >> (p0 - n_local)) & 1:
                gl.add_pauli("X", p1)
        else:
            gl._add(kind, p0, p1, pidx, fixed)
    gl.n_params = n_params
    return gl


def rank_terms(group_terms, pos, n_local, rank):
    """Pauli terms [(x, z, coeff)] of logical masks -> (x_local, z_local, coeff * sign) for this rank in layout `pos`."""
    xs, zs, cs = [], [], []
    for x, z, c in group_terms:
        xl = zl = 0
        sign = 1.0
        for q in _bits(int(x) | int(z)):
            p = pos[q]
            if p < n_local:
                xl |= ((int(x) >> q) & 1) << p
                zl |= ((int(z) >> q) & 1) << p
            else:   # a flip on a rank bit cannot occur here (the schedule makes X/Y qubits local): this is a Z factor
                if (rank >> (p - n_local)) & 1:
                    sign = -sign
        xs.append(xl)
        zs.append(zl)
        cs.append(c * sign)
    return np.asarray(xs, dtype=np.uint64), np.asarray(zs, dtype=np.uint64), np.asarray(cs)


# ---------------------------------------------------------------------------------------------------------------
# communication
# ---------------------------------------------------------------------------------------------------------------
class TorchComm:
    """One rank per process over torch.distributed (NCCL between GPUs; gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist, self.group = dist, group
        self.size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.ranks = [self.rank]
        self._spare = None

    def exchange(self, shards):
        import torch
        src = shards[0]
        if self._spare is None or self._spare.shape != src.shape or self._spare.device != src.device:
            self._spare = torch.empty_like(src)
        out = self._spare
        self._dist.all_to_all_single(out, src, group=self.group)   # equal contiguous chunks: chunk c -> rank c
        self._spare = src
        return [out]

    def total(self, partials):
        t = partials[0].clone()
        self._dist.all_reduce(t, group=self.group)
        return t

    # fused exchange: two shard buffers per rank, every rank maps all peers' buffers (CUDA IPC)
    def setup_exchange(self, engine):
        import torch
        lib, dev = engine.lib, engine.index
        mine = [engine.alloc_shard() for _ in range(2)]
        handles = torch.zeros(2, 64, dtype=torch.uint8)
        for s in range(2):
            buf = (ctypes.c_uint8 * 64)()
            if lib.tq_ipc_export(dev, ctypes.c_void_p(mine[s][1]), buf) != 0:
                raise RuntimeError("cudaIpcGetMemHandle failed")
            handles[s] = torch.frombuffer(bytearray(buf), dtype=torch.uint8)
        mine_dev = handles.to(engine.device)
        everyone = torch.empty(self.size, 2, 64, dtype=torch.uint8, device=engine.device)
        self._dist.all_gather_into_tensor(everyone, mine_dev, group=self.group)
        everyone = everyone.cpu()
        tables, self._opened = [[0] * self.size for _ in range(2)], []
        for r in range(self.size):
            for s in range(2):
                if r == self.rank:
                    tables[s][r] = mine[s][1]
                    continue
                raw = (ctypes.c_uint8 * 64)(*everyone[r, s].tolist())
                out = ctypes.c_void_p()
                if lib.tq_ipc_open(dev, raw, ctypes.byref(out)) != 0:
                    raise RuntimeError("cudaIpcOpenMemHandle failed (peer access between the GPUs of this job?)")
                tables[s][r] = out.value
                self._opened.append(out.value)
        self._token = torch.zeros(1, dtype=torch.float32, device=engine.device)
        return [[mine[0][0]], [mine[1][0]]], tables

    def meet(self):
        """stream-ordered rendezvous: returns (on the stream) once every rank's preceding kernels have finished"""
        self._dist.all_reduce(self._token, group=self.group)

    def release_exchange(self, engine):
        """unmap the peers' buffers (collective: every rank must have stopped storing into them)"""
        import torch
        opened = getattr(self, "_opened", None)
        if not opened:
            return
        torch.cuda.synchronize(engine.device)
        self._dist.barrier(group=self.group)
        try:
            for ptr in opened:
                engine.lib.tq_ipc_close(engine.index, ctypes.c_void_p(ptr))
        finally:
            self._opened = []
            # second rendezvous: no rank may free its own (exported) shard buffers while a slower peer still has them
            # mapped -- cudaFree of an allocation that is open in another process is undefined behaviour
            self._dist.barrier(group=self.group)


class LocalComm:
    """R virtual ranks in one process (one device): the all-to-all is a transpose of the [rank][chunk] grid."""

    def __init__(self, size):
        self.size = int(size)
        self.rank = 0
        self.ranks = list(range(self.size))

    def exchange(self, shards):
        import torch
        R = self.size
        grid = torch.stack([s.reshape(R, -1) for s in shards])          # [source rank][chunk][...]
        return [grid[:, r].reshape(-1).contiguous() for r in range(R)]  # rank r receives chunk r of every source

    def total(self, partials):
        t = partials[0].clone()
        for p in partials[1:]:
            t = t + p
        return t

    def setup_exchange(self, engine):
        bufs = [[engine.alloc_shard() for _ in self.ranks] for _ in range(2)]
        return [[b[0] for b in bufs[s]] for s in range(2)], [[b[1] for b in bufs[s]] for s in range(2)]

    def meet(self):
        pass   # one process, one stream: launches are already ordered

    def release_exchange(self, engine):
        pass


# ---------------------------------------------------------------------------------------------------------------
# engine
# ---------------------------------------------------------------------------------------------------------------
class _DeviceBuffer:
    """float64 view of a tq_device_alloc allocation for torch.as_tensor (CUDA array interface); frees it when dropped"""

    def __init__(self, address, n_doubles, lib, device):
        self._address, self._lib, self._device = address, lib, device
        self.__cuda_array_interface__ = {"shape": (int(n_doubles),), "typestr": "<f8", "data": (int(address), False),
                                         "version": 2, "strides": None}

    def __del__(self):
        try:
            self._lib.tq_device_free(self._device, ctypes.c_void_p(self._address))
        except Exception:
            pass


class GpuEngine:
    """libtqsim on one device: every (rank, step) owns a handle whose compiled plan is reused across evaluations."""

    def __init__(self, n_local, device):
        import torch
        self.n_local = int(n_local)
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        from . import _lib
        self.lib = _lib.lib()

    supports_fused_exchange = True

    def zeros(self):
        import torch
        return torch.zeros(2 << self.n_local, dtype=torch.float64, device=self.device)   # interleaved (re, im)

    def alloc_shard(self):
        """(tensor, address) of a shard buffer that is a cudaMalloc allocation of its own (exportable over CUDA IPC)"""
        import torch
        out = ctypes.c_void_p()
        if self.lib.tq_device_alloc(self.index, 16 << self.n_local, ctypes.byref(out)) != 0:
            raise MemoryError("tq_device_alloc failed")
        view = _DeviceBuffer(out.value, 2 << self.n_local, self.lib, self.index)
        t = torch.as_tensor(view, device=self.device)
        t._tq_owner = view     # keeps the allocation alive as long as the tensor
        return t, out.value

    def run_exchange(self, step, shard, params, recv_ptrs, rank, n_ranks):
        step.evolve_states_exchange(shard, params if step.n_params > 0 else None, n_ranks, rank, recv_ptrs)

    def params(self, values):
        import torch
        return torch.as_tensor(np.ascontiguousarray(values, dtype=np.float64).reshape(1, -1)).to(self.device)

    def make_step(self, gatelist, pauli):
        from .simulator import Simulator
        sim = Simulator(self.n_local, self.index)
        sim.set_circuit(gatelist)
        if pauli is not None:
            sim.set_pauli_hamiltonian(*pauli)
        return sim

    def run(self, step, shard, params, want_energy):
        return step.evolve_states(shard, params if step.n_params > 0 else None, energies=want_energy)


# ---------------------------------------------------------------------------------------------------------------
# driver
# ---------------------------------------------------------------------------------------------------------------
class ShardedSimulator:
    """Energy of one n-qubit circuit whose state is sharded over comm.size = 2^g ranks.  Mirrors `Simulator`'s problem
    definition calls (set_circuit / set_pauli_hamiltonian) and evaluates one parameter vector per `energy()` call."""

    def __init__(self, n_qubits, comm, device=0, engine=None, fused_exchange=None):
        self.n_qubits = int(n_qubits)
        self.comm = comm
        g = int(comm.size).bit_length() - 1
        if comm.size != 1 << g:
            raise ValueError("the number of ranks must be a power of two")
        self.g, self.n_local = g, self.n_qubits - g
        self.engine = engine if engine is not None else GpuEngine(self.n_local, device)
        self._gl = None
        self._pauli = None
        self._program = None
        self.n_exchanges = 0
        # the exchange as the write-back of the segment's last tile pass (tensor-core passes: shards of >= 2^9 amplitudes)
        # default (None): fused up to four ranks -- measured on B200 at 28 qubits: 16.7 vs 19.8 ms on two GPUs, 10.2 vs 10.5 ms
        # on four, 6.5 vs 6.1 ms on eight, where the many-to-many 16-byte peer stores fall behind NCCL's all-to-all
        if fused_exchange is None:
            fused_exchange = comm.size <= 4
        self.fused_exchange = bool(fused_exchange and getattr(self.engine, "supports_fused_exchange", False) and
                                   2 <= comm.size <= 8 and 9 <= self.n_local <= 28)
        self._xbufs = self._xtables = self._copy_steps = None

    def close(self):
        """Release the handles of the compiled schedule and the exchange buffers (collective under TorchComm)."""
        for st in self._program or []:
            if st[0] == "run":
                for step in st[1].values():
                    if hasattr(step, "close"):
                        step.close()
        for step in (self._copy_steps or {}).values():
            if hasattr(step, "close"):
                step.close()
        self._program = self._copy_steps = None
        if self._xbufs is not None:
            try:
                self.comm.release_exchange(self.engine)   # (two barriers: unmap everywhere, THEN drop the own buffers)
            finally:
                self._xbufs = self._xtables = None

    def set_circuit(self, gl):
        if gl.n_qubits != self.n_qubits:
            raise ValueError("circuit and simulator disagree on the number of qubits")
        self._gl, self._program = gl, None

    def set_pauli_hamiltonian(self, xmask, zmask, coeff):
        self._pauli = ([int(v) for v in xmask], [int(v) for v in zmask], list(np.asarray(coeff)))
        self._program = None

    # the schedule, materialised per local rank: [("run", {rank: step}, want_energy) | ("exchange",)]
    def _compile(self):
        if self._gl is None:
            raise RuntimeError("no circuit set")
        xs, zs, cs = self._pauli if self._pauli is not None else ([], [], [])
        masks = sorted(set(xs))
        group_of = {m: i for i, m in enumerate(masks)}
        terms = [[] for _ in masks]
        for x, z, c in zip(xs, zs, cs):
            terms[group_of[x]].append((x, z, c))
        steps = plan_state_sharding(self._gl.tuples(), self.n_qubits, self.g, masks)
        program, i = [], 0
        while i < len(steps):
            st = steps[i]
            if st[0] == "exchange":
                program.append(("exchange",))
                i += 1
                continue
            ops = st[1] if st[0] == "evolve" else []
            expect = None
            if st[0] == "evolve" and i + 1 < len(steps) and steps[i + 1][0] == "expect":
                expect = steps[i + 1]
                i += 1
            elif st[0] == "expect":
                expect = st
            i += 1
            per_rank = {}
            for r in self.comm.ranks:
                gl = rank_gatelist(ops, self.n_local, r, self._gl.n_params)
                pauli = None
                if expect is not None:
                    pauli = rank_terms([t for gi in expect[1] for t in terms[gi]], expect[2], self.n_local, r)
                if len(gl) or pauli is not None:
                    per_rank[r] = self.engine.make_step(gl, pauli)
            program.append(("run", per_rank, expect is not None))
        self._program = program
        self.n_exchanges = sum(1 for p in program if p[0] == "exchange")

    def energy(self, params=None):
        """<psi(params)|H|psi(params)> from |0...0>; every rank returns the same float."""
        if self._program is None:
            self._compile()
        if self.fused_exchange:
            return self._energy_fused(params)
        eng = self.engine
        shards = [eng.zeros() for _ in self.comm.ranks]
        if 0 in self.comm.ranks:
            shards[self.comm.ranks.index(0)][0] = 1.0
        p = eng.params(params if params is not None else np.zeros(max(1, self._gl.n_params)))
        partials = [None] * len(shards)
        for st in self._program:
            if st[0] == "exchange":
                shards = self.comm.exchange(shards)
                continue
            for k, r in enumerate(self.comm.ranks):
                step = st[1].get(r)
                if step is None:
                    continue
                e = eng.run(step, shards[k], p, st[2])
                if st[2]:
                    partials[k] = e if partials[k] is None else partials[k] + e
        zero = eng.params([0.0]).reshape(-1)[:1] * 0.0
        total = self.comm.total([zero if e is None else e.reshape(-1)[:1] for e in partials])
        return float(total.reshape(-1)[0].item())

    def _energy_fused(self, params):
        eng, comm = self.engine, self.comm
        if self._xbufs is None:
            self._xbufs, self._xtables = comm.setup_exchange(eng)
            self._copy_steps = {}
        p = eng.params(params if params is not None else np.zeros(max(1, self._gl.n_params)))
        cur = 0
        for k, r in enumerate(comm.ranks):
            self._xbufs[0][k].zero_()
            if r == 0:
                self._xbufs[0][k][0] = 1.0
        partials = [None] * len(comm.ranks)

        def send(per_rank):
            # every local rank pushes its shard (evolved through its part of the segment, if any) into the receive buffers
            nonlocal cur
            for k, r in enumerate(comm.ranks):
                step = per_rank.get(r)
                if step is None:   # nothing to apply on this rank: a gate-free pass that only moves the shard
                    if r not in self._copy_steps:
                        self._copy_steps[r] = eng.make_step(GateList(self.n_local), None)
                    step = self._copy_steps[r]
                eng.run_exchange(step, self._xbufs[cur][k], p, self._xtables[1 - cur], r, comm.size)
            comm.meet()
            cur = 1 - cur

        prog, i = self._program, 0
        while i < len(prog):
            st = prog[i]
            if st[0] == "exchange":
                send({})
            elif not st[2] and i + 1 < len(prog) and prog[i + 1][0] == "exchange":
                send(st[1])
                i += 1
            else:
                for k, r in enumerate(comm.ranks):
                    step = st[1].get(r)
                    if step is None:
                        continue
                    e = eng.run(step, self._xbufs[cur][k], p, st[2])
                    if st[2]:
                        partials[k] = e if partials[k] is None else partials[k] + e
            i += 1
        zero = eng.params([0.0]).reshape(-1)[:1] * 0.0
        total = comm.total([zero if e is None else e.reshape(-1)[:1] for e in partials])
        return float(total.reshape(-1)[0].item())
