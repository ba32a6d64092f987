"""Run under torchrun (any world size >= 1, one rank per GPU): the two gather drivers of tensorrl_qas_b200.sharding on real
NCCL -- OverlappedGather (device-resident steps, gather under the next step) and HostBatchGather (host buffers in, all
ranks' energies out) -- against one-call-at-a-time host evaluations.  Prints GATHER_OK on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from tensorrl_qas_b200 import Simulator, loaders
    from tensorrl_qas_b200.circuit import parameter_batch, synthetic_circuit
    from tensorrl_qas_b200.sharding import HostBatchGather, OverlappedGather

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, B = 13, 6
    gl = synthetic_circuit(n, 90, 3)
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    sim = Simulator(n, local)
    sim.set_circuit(gl)
    sim.set_pauli_hamiltonian(x, z, w)
    every = [parameter_batch(gl, B, seed0=1000 + r * B) for r in range(world)]   # rank r owns rows [r*B, (r+1)*B)
    mine = every[rank]
    want_mine = sim.energies(mine)
    want_all = np.concatenate([sim.energies(p) for p in every])

    hg = HostBatchGather(sim, B, mine.shape[1], world, dev)
    for _ in range(3):
        got = hg(mine).copy()
        assert np.array_equal(got[rank * B:(rank + 1) * B], want_mine)
        assert np.abs(got - want_all).max() < 1e-12

    og = OverlappedGather(B, world, dev)
    p_dev = torch.from_numpy(mine).to(dev)
    for _ in range(5):
        sim.energies_dev(p_dev, out=og.local_buffer())
        og.submit()
    full = og.wait()
    torch.cuda.synchronize(dev)
    assert np.abs(full.cpu().numpy() - want_all).max() < 1e-12
    dist.barrier()
    if rank == 0:
        print("GATHER_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
