import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from tensorrl_qas_b200 import Simulator, loaders
from tensorrl_qas_b200.circuit import brickwork_circuit, parameter_batch, GateList, append_random_gates
from tensorrl_qas_b200.simulator import plan_dump
n = 12
rng = np.random.default_rng(1)
gl = GateList(n)
# many dense bricks on a fixed set of 6 qubits -> one window, hundreds of blocks
pairs = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)] if len(sys.argv) < 2 else [(0, 1), (0, 2), (0, 3), (0, 4), (0, 5)]
for rep in range(40):
    for (a, b) in pairs:
        append_random_gates(gl, 12, rng, (a, b))
plan = plan_dump(gl, 0, 12, 3)
ops = [o[0] for p in plan for w in p['windows'] for o in w['ops']]
print("windows", sum(len(p['windows']) for p in plan), "U2", ops.count(16), "SWAP", ops.count(17), "CXO", ops.count(18))
paulis, w = loaders.heisenberg_terms(n)
x, z = loaders.pauli_masks(paulis, n)
sim = Simulator(n, 0); sim.set_pauli_hamiltonian(x, z, w); sim.set_circuit(gl)
B = 4096
p = torch.from_numpy(parameter_batch(gl, 1).repeat(B, 0)).cuda()
out = torch.empty(B, dtype=torch.float64, device='cuda')
for _ in range(3): sim.energies_dev(p, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): sim.energies_dev(p, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
nU2 = ops.count(16)
fma = B * (1 << n) * nU2 * 16
print("ms", ms, "FMA/clk/SM", fma / (ms * 1e-3) / 148 / 1.965e9)
