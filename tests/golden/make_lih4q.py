"""Fixture generator (build container): python tests/golden/make_lih4q.py
Copies the arrays of the reference's shipped LiH-4q file (dmrg-to-qc/mol_data/LIH_4q_geom_Li_.0_.0_.0;_H_.0_.0_3.4_parity.npz:
dense 16x16 complex64 Hamiltonian, its 100 complex128 weights, eigvals, energy_shift) into tests/golden/lih_4q_parity.npz --
the Hamiltonian of bench workload C1 (the file has no `paulis` key, cfg or init circuit)."""
import os

import numpy as np

REF = os.environ.get("TQ_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "dmrg-to-qc", "mol_data", "LIH_4q_geom_Li_.0_.0_.0;_H_.0_.0_3.4_parity.npz")
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    g = np.load(SRC, allow_pickle=False)
    np.savez_compressed(os.path.join(HERE, "lih_4q_parity.npz"), **{k: g[k] for k in g.files})
    print({k: g[k].shape for k in g.files})
