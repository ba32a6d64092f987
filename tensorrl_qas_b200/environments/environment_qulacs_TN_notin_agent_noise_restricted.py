"""Drop-in for the reference's environments/environment_qulacs_TN_notin_agent_noise_restricted.py -- hexagon-connectivity action set, optional shot noise, no gate noise.
Same module name, class name and public surface; the logic lives in `_core.CircuitEnvBase`."""
from ..VQAs import VQE_qulacs_TN_notin_RL_noise_restricted as vc
from ._core import CircuitEnvBase


class CircuitEnv(CircuitEnvBase):
    vc = vc
    tn_in_agent = False
    shot_args = True
    restricted = True


if __name__ == "__main__":
    pass
