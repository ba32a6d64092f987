import os, sys, shutil
sys.path.insert(0, '.')
import tensorrl_qas_b200._lib as L
L.LIB_PATH = os.path.abspath('scratch/libtqsim_trace'+os.environ.get('TRW','3')+'.so')
import numpy as np, torch
from tensorrl_qas_b200 import Simulator, loaders
from tensorrl_qas_b200.circuit import brickwork_circuit, parameter_batch
gl = brickwork_circuit(20, 21, 41, 5)
paulis, w = loaders.heisenberg_terms(20)
x, z = loaders.pauli_masks(paulis, 20)
sim = Simulator(20, 0); sim.set_pauli_hamiltonian(x, z, w); sim.set_circuit(gl)
p = parameter_batch(gl, 64)
e = sim.energies(p); print("SECOND"); e = sim.energies(p)
print(e[:3])
