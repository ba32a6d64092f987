"""`environments` package of the reference, B200-backed: `environment_qulacs*.CircuitEnv`, `utils`, `VQAs`.

`install()` registers this package under the reference's import name, so the unmodified drivers' and agents'
`from environments.environment_qulacs_TN_notin_agent import CircuitEnv` / `from environments.utils.utils import
get_config` resolve here."""
import importlib
import sys

_MODULES = (
    "environment_qulacs", "environment_qulacs_noise", "environment_qulacs_TN_notin_agent",
    "environment_qulacs_TN_notin_agent_noise", "environment_qulacs_TN_notin_agent_noise_restricted",
    "utils", "utils.utils", "utils.utils_topology_restrict", "utils.curricula",
)
_VQA_MODULES = (
    "VQE_qulacs", "VQE_qulacs_noise", "VQE_qulacs_TN_notin_RL", "VQE_qulacs_TN_notin_RL_noise",
    "VQE_qulacs_TN_notin_RL_noise_restricted",
)


def install(name="environments"):
    """Alias this package (and the VQA shims as `<name>.VQAs`) in sys.modules; returns the package."""
    me = sys.modules[__name__]
    sys.modules[name] = me
    for sub in _MODULES:
        sys.modules[f"{name}.{sub}"] = importlib.import_module(f"{__name__}.{sub}")
    vqas = importlib.import_module("tensorrl_qas_b200.VQAs")
    sys.modules[f"{name}.VQAs"] = vqas
    me.VQAs = vqas
    for sub in _VQA_MODULES:
        sys.modules[f"{name}.VQAs.{sub}"] = importlib.import_module(f"tensorrl_qas_b200.VQAs.{sub}")
    return me
