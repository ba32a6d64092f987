"""N > 1 path on CPU: world_size-2 gloo ranks shard a parameter batch, evaluate their slice with the oracle (standing in
for the per-GPU simulator) and all-gather the energies."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tensorrl_qas_b200 import loaders
from tensorrl_qas_b200.circuit import parameter_batch, synthetic_circuit
from tensorrl_qas_b200.sharding import gather_energies, shard_bounds, sharded_energies


def test_shard_bounds_cover_the_batch():
    for batch in (0, 1, 7, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    n = 6
    gl = synthetic_circuit(n, 30, 3)
    params = parameter_batch(gl, batch)
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    calls = []

    def evaluate(p):
        calls.append(len(p))
        return c_oracle.energies(gl, p, pauli=(x, z, w), nthreads=1)

    full = sharded_energies(evaluate, params)
    lo, hi = shard_bounds(batch, rank, world)
    assert calls == ([hi - lo] if hi > lo else [])
    # the raw collective helper on a ragged split
    local = torch.arange(lo, hi, dtype=torch.float64)
    assert torch.equal(gather_energies(local, batch), torch.arange(batch, dtype=torch.float64))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full)
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [5, 8])
def test_two_rank_gloo_sharded_energies(tmp_path, oracle, batch):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), batch, str(tmp_path)), nprocs=world, join=True)
    n = 6
    gl = synthetic_circuit(n, 30, 3)
    params = parameter_batch(gl, batch)
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    want = oracle.energies(gl, params, pauli=(x, z, w))
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert got.shape == (batch,) and np.array_equal(got, want)  # every rank sees the full, identical vector


def test_single_process_passthrough(oracle):
    gl = synthetic_circuit(4, 10, 1)
    p = parameter_batch(gl, 3)
    out = sharded_energies(lambda q: np.arange(len(q), dtype=np.float64), p)
    assert np.array_equal(out, [0.0, 1.0, 2.0])


@pytest.mark.gpu
def test_nccl_gather_drivers_on_the_gpus_of_the_box(built_lib):
    """OverlappedGather and HostBatchGather (the N > 1 legs of bench.py) on real NCCL: one rank per visible GPU, at most
    two (a one-GPU box runs a single-rank communicator -- same code path, collective included)."""
    import socket
    import subprocess
    import torch
    nproc = min(2, torch.cuda.device_count())
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "nccl_gather_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert f"GATHER_OK {nproc}" in res.stdout
