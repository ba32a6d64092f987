"""Batch sharding of independent energy evaluations over the GPUs of one box.

Units of work (one parameter set -> one float64) are independent, so the only collective on the path is the
all-gather of the per-rank energies.  One process per GPU; `torch.distributed` is plumbing (NCCL on GPUs, gloo in the
CPU tests).  The reference itself is single-process (no distributed code at all)."""
import numpy as np


def shard_bounds(batch, rank, world):
    """Contiguous slice [lo, hi) of a global batch owned by `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, extra = divmod(int(batch), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_energies(local, batch, group=None):
    """All-gather the ranks' energy slices (torch tensor, CPU for gloo / CUDA for NCCL) into the global [batch] vector.
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    width = max(shard_bounds(batch, r, world)[1] - shard_bounds(batch, r, world)[0] for r in range(world))
    lo, hi = shard_bounds(batch, rank, world)
    if local.numel() != hi - lo:
        raise ValueError(f"rank {rank} holds {local.numel()} energies, expected {hi - lo}")
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[:hi - lo] = local
    out = torch.empty(world * width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = []
    for r in range(world):
        a, b = shard_bounds(batch, r, world)
        parts.append(out[r * width:r * width + (b - a)])
    return torch.cat(parts)


def sharded_energies(evaluate, params, group=None):
    """Evaluate a global [B][P] parameter matrix: every rank runs `evaluate` (-> numpy [b]) on its slice of the rows and
    all ranks receive the full [B] vector.  `evaluate` is e.g. `Simulator.energies`."""
    import torch
    import torch.distributed as dist
    params = np.asarray(params, dtype=np.float64)
    batch = params.shape[0]
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(evaluate(params))
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(batch, rank, world)
    local = np.asarray(evaluate(params[lo:hi]), dtype=np.float64) if hi > lo else np.zeros(0)
    t = torch.from_numpy(local.copy())
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    return gather_energies(t, batch, group).cpu().numpy()


class OverlappedGather:
    """The path's only collective, off the critical path: the all-gather of step i runs on a side stream while the kernels
    of step i + 1 are already executing (the energies alternate between two buffers, so the next step never overwrites
    what is being gathered).  One rank per GPU, NCCL.

        g = OverlappedGather(batch_per_rank, world, device)
        for step in ...:
            out = g.local_buffer()              # this step's energies go here (torch float64 [batch_per_rank])
            sim.energies_dev(params, out=out)
            g.submit()                          # returns at once
        all_energies = g.wait()                 # [batch_per_rank * world] of the LAST step; earlier ones via g.result(k)
    """

    def __init__(self, batch_per_rank, world, device, group=None, depth=2):
        import torch
        self.torch = torch
        self.group = group
        self.device = device
        self.side = torch.cuda.Stream(device=device)
        self.local = [torch.empty(batch_per_rank, dtype=torch.float64, device=device) for _ in range(depth)]
        self.full = [torch.empty(batch_per_rank * world, dtype=torch.float64, device=device) for _ in range(depth)]
        self.done = [None] * depth
        self.i = 0

    def local_buffer(self):
        k = self.i % len(self.local)
        if self.done[k] is not None:   # the gather that last used this pair of buffers must be over before they are rewritten
            self.torch.cuda.current_stream(self.device).wait_event(self.done[k])
        return self.local[k]

    def submit(self):
        import torch.distributed as dist
        torch = self.torch
        k = self.i % len(self.local)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            dist.all_gather_into_tensor(self.full[k], self.local[k], group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[k] = ev
        self.i += 1
        return k

    def result(self, k):
        self.done[k].synchronize()
        return self.full[k]

    def wait(self):
        """Block the current stream on every outstanding gather and return the last step's global vector."""
        cur = self.torch.cuda.current_stream(self.device)
        for ev in self.done:
            if ev is not None:
                cur.wait_event(ev)
        return self.full[(self.i - 1) % len(self.full)] if self.i else None


class HostBatchGather:
    """Host buffers in, every rank's energies out, one synchronisation per call: this rank's angle rows go from pinned host
    memory to the device, the evaluation and the all-gather run on the device (same stream), the gathered [batch_per_rank *
    world] vector comes back into pinned host memory.  (The plain host entry point followed by `gather_energies` costs a
    device->host->device round trip of the local energies and a second synchronisation before the collective.)

        g = HostBatchGather(sim, batch_per_rank, n_params, world, device)
        all_energies = g(params_host)            # numpy [batch_per_rank * world], valid until the next call
    """

    def __init__(self, sim, batch_per_rank, ld, world, device, group=None, mode="pure"):
        import torch
        self.torch, self.sim, self.group, self.device, self.mode = torch, sim, group, device, mode
        self.p_pin = torch.empty((batch_per_rank, ld), dtype=torch.float64).pin_memory()
        self.p_dev = torch.empty((batch_per_rank, ld), dtype=torch.float64, device=device)
        self.e_loc = torch.empty(batch_per_rank, dtype=torch.float64, device=device)
        self.e_all = torch.empty(batch_per_rank * world, dtype=torch.float64, device=device)
        self.e_pin = torch.empty(batch_per_rank * world, dtype=torch.float64).pin_memory()
        self.h2d_bytes = self.p_pin.numel() * 8
        self.d2h_bytes = self.e_pin.numel() * 8

    def __call__(self, params):
        import torch.distributed as dist
        torch = self.torch
        self.p_pin.copy_(torch.from_numpy(np.ascontiguousarray(params, dtype=np.float64)))
        self.p_dev.copy_(self.p_pin, non_blocking=True)
        self.sim.energies_dev(self.p_dev, out=self.e_loc, mode=self.mode)
        dist.all_gather_into_tensor(self.e_all, self.e_loc, group=self.group)
        self.e_pin.copy_(self.e_all, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.e_pin.numpy()
