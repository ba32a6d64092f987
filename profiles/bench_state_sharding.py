"""Single-state sharding measurement (not the headline bench -- that is bench.py on C5).

One n-qubit Heisenberg-chain energy evaluation (C5-shaped brickwork circuit: (n-1) bricks x 21 gates + 41 gates, seed 5)
with the state sharded over the ranks of the job:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        profiles/bench_state_sharding.py --qubits 28 --steps 5 --warmup 2
Without torchrun: --virtual R runs R virtual ranks on one GPU (same schedule, all-to-all as a transpose).
--check-oracle (n <= 22) compares with the CPU oracle, --check-unsharded with the unsharded kernels on rank 0.
Rank 0 prints one JSON line; times are max over ranks around K evaluations (each ends in a device->host read)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--virtual", type=int, default=0)
    ap.add_argument("--check-oracle", action="store_true")
    ap.add_argument("--check-unsharded", action="store_true")
    ap.add_argument("--no-fused", action="store_true", help="NCCL all-to-all between segments instead of the fused write-back")
    ap.add_argument("--fused", action="store_true", help="force the fused write-back (default: up to four ranks)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from tensorrl_qas_b200 import Simulator, loaders
    from tensorrl_qas_b200.circuit import brickwork_circuit
    from tensorrl_qas_b200.sharded import LocalComm, ShardedSimulator, TorchComm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = TorchComm()
    else:
        comm = LocalComm(max(1, args.virtual))

    n = args.qubits
    gl = brickwork_circuit(n, 21, 41, 5)
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    p = np.asarray(gl.initial_angles)
    sim = ShardedSimulator(n, comm, device=local_rank, fused_exchange=False if args.no_fused else (True if args.fused else None))
    sim.set_circuit(gl)
    sim.set_pauli_hamiltonian(x, z, w)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    e = None
    for _ in range(args.warmup):
        e = sim.energy(p)
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e = sim.energy(p)
    sync()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    ms = float(dt.item()) * 1e3 / max(1, args.steps)

    out = {"metric": "energy evals/sec (one fp64 statevector sharded over the ranks)", "qubits": n,
           "ranks": comm.size, "virtual_ranks": world == 1, "gates": len(gl), "pauli_terms": len(w),
           "exchanges_per_eval": sim.n_exchanges, "fused_exchange": sim.fused_exchange, "shard_bytes": 16 << sim.n_local,
           "exchange_bytes_per_rank_per_eval": sim.n_exchanges * (16 << sim.n_local) * (comm.size - 1) // comm.size,
           "ms_per_eval": ms, "value": 1e3 / ms, "unit": "evals/s", "steps": args.steps, "warmup": args.warmup,
           "energy": e}
    if args.check_oracle and rank == 0:
        from oracle import c_oracle
        c_oracle.build()
        out["oracle_abs_err"] = abs(e - float(c_oracle.energies(gl, p[None, :], pauli=(x, z, w))[0]))
    if args.check_unsharded and rank == 0:
        one = Simulator(n, local_rank)
        one.set_circuit(gl)
        one.set_pauli_hamiltonian(x, z, w)
        pd = torch.from_numpy(p[None, :].copy()).cuda()
        e1 = one.energies_dev(pd)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e1 = one.energies_dev(pd)
        torch.cuda.synchronize()
        out["unsharded_ms_per_eval"] = (time.perf_counter() - t0) * 1e3
        out["unsharded_abs_diff"] = abs(e - float(e1[0].item()))
        one.close()
    sim.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))
        tol = 1e-10
        if out.get("oracle_abs_err", 0.0) > tol or out.get("unsharded_abs_diff", 0.0) > tol:
            sys.exit(1)


if __name__ == "__main__":
    main()
