"""Drop-in for the reference's environments/environment_qulacs_noise.py -- sampled depolarising noise on every gate, MPS circuit in the agent's state.
Same module name, class name and public surface; the logic lives in `_core.CircuitEnvBase`."""
from ..VQAs import VQE_qulacs_noise as vc
from ._core import CircuitEnvBase


class CircuitEnv(CircuitEnvBase):
    vc = vc
    tn_in_agent = True
    shot_args = True
    restricted = False


if __name__ == "__main__":
    pass
