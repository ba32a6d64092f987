"""CPU stand-in for `tensorrl_qas_b200.sharded.GpuEngine` (test infrastructure): the same four calls, executed with the
numpy oracle on CPU torch tensors, so that the sharding schedule, the layout bookkeeping and the collectives can be
checked without a GPU.  Never imported by the product."""
import numpy as np
import torch

from oracle import np_oracle


class NumpyEngine:
    def __init__(self, n_local):
        self.n_local = int(n_local)
        self.runs = 0

    def zeros(self):
        return torch.zeros(2 << self.n_local, dtype=torch.float64)

    def params(self, values):
        return torch.as_tensor(np.ascontiguousarray(values, dtype=np.float64).reshape(1, -1))

    def make_step(self, gatelist, pauli):
        return (gatelist, pauli)

    def run(self, step, shard, params, want_energy):
        gl, pauli = step
        self.runs += 1
        v = shard.numpy().view(np.complex128)      # shares memory with the tensor: evolved in place
        if len(gl):
            v[:] = np_oracle.run_circuit(self.n_local, gl.tuples(), params.numpy().reshape(-1), init=v.copy())
        if not want_energy:
            return None
        return torch.tensor([np_oracle.expect_pauli(v, *pauli)], dtype=torch.float64)
