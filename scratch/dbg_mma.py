import sys, os
sys.path.insert(0, '.')
import numpy as np
from oracle import c_oracle
from tensorrl_qas_b200 import Simulator, loaders
from tensorrl_qas_b200.circuit import synthetic_circuit, parameter_batch, GateList
c_oracle.build()
for n, G in [(9, 40), (10, 60), (12, 80), (13, 60), (14, 80)]:
    gl = synthetic_circuit(n, G, 3)
    p = parameter_batch(gl, 3)
    sim = Simulator(n, 0)
    sim.set_circuit(gl)
    st = sim.states(p)
    want = np.stack([c_oracle.state(gl, p[b]) for b in range(3)])
    print(n, G, "state err", np.abs(st - want).max())
    paulis, w = loaders.heisenberg_terms(n)
    x, z = loaders.pauli_masks(paulis, n)
    # per-term energies
    worst = []
    for sel in [slice(0, None)] + [slice(i, i + 1) for i in range(len(paulis))]:
        sim.set_pauli_hamiltonian(x[sel], z[sel], w[sel])
        e = sim.energies(p)
        ew = c_oracle.energies(gl, p, pauli=(x[sel], z[sel], w[sel]))
        err = np.abs(e - ew).max()
        if err > 1e-10:
            worst.append((paulis[sel][0] if sel.start else "ALL", err))
    print("   bad terms:", worst[:12], len(worst))
    sim.close()
