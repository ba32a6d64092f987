"""Fixture generator (run in the build container, where /root/reference exists):
    python tests/golden/make_init_circuits.py
Packs the reference's shipped MPS init circuits, dmrg-to-qc/init_state_circ/*.qpy (raw bytes) and their OpenQASM twins
(*.qasm, raw text), into tests/golden/init_circuits.npz.  tests/test_oracle.py parses the QPY bytes with
tensorrl_qas_b200.loaders and the OpenQASM text with tests/qasm2_reader.py -- two independent readers of two serialisations
the reference wrote from the same qiskit circuit (dmrg-to-qc/tnqc_ansatze.py: qpy.dump + qasm2.dump) -- and requires gate-
for-gate agreement."""
import glob
import os
import sys

import numpy as np

REF = os.environ.get("TQ_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    files = sorted(glob.glob(os.path.join(REF, "dmrg-to-qc", "init_state_circ", "*.qpy")))
    if not files:
        sys.exit(f"no QPY files under {REF}")
    out, names = {}, []
    for qpy in files:
        qasm = qpy[:-4] + ".qasm"
        if not os.path.exists(qasm):
            continue
        key = f"c{len(names):02d}"
        names.append(os.path.basename(qpy)[:-4])
        out[f"{key}/qpy"] = np.frombuffer(open(qpy, "rb").read(), dtype=np.uint8)
        out[f"{key}/qasm"] = np.frombuffer(open(qasm, "rb").read(), dtype=np.uint8)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "init_circuits.npz"), **out)
    print(f"{len(names)} circuit pairs -> tests/golden/init_circuits.npz")


if __name__ == "__main__":
    main()
