"""Drop-in for the reference's environments/environment_qulacs.py -- noiseless, MPS circuit encoded in the agent's state ("trainable" / StructureRL drivers).
Same module name, class name and public surface; the logic lives in `_core.CircuitEnvBase`."""
from ..VQAs import VQE_qulacs as vc
from ._core import CircuitEnvBase


class CircuitEnv(CircuitEnvBase):
    vc = vc
    tn_in_agent = True
    shot_args = False
    restricted = False


if __name__ == "__main__":
    pass
