"""tensorrl_qas_b200 -- B200-native (sm_100a) back end for the TensorRL-QAS environment hot path.

The product is `libtqsim.so` (hand-written CUDA behind the C ABI in include/tqsim.h); this package is the thin
Python host layer that mirrors the reference's operator interface for that path:

  * `Simulator`                                   -- ctypes handle over the C ABI
  * `VQAs.VQE_qulacs*`                            -- `Parametric_Circuit`, `get_energy_qulacs`, `get_exp_val`
                                                    (reference: environments/VQAs/VQE_qulacs*.py)
  * `loaders`                                     -- qiskit-free readers for the shipped QPY circuits / npz Hamiltonians

There is no CPU fallback: importing works anywhere, evaluating needs a B200 and the built library.
"""
from .circuit import GateList, KIND  # noqa: F401
from .simulator import Simulator, TqError  # noqa: F401

__version__ = "0.1.0"
