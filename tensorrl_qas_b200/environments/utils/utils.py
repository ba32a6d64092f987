"""Config reader and action dictionaries of the environments (reference: environments/utils/utils.py, byte-identical
copy in agents/utils.py).  The INI schema of configuration_files/**.cfg is frozen: this reader
yields the same typed dict, key by key, as the reference's `get_config` (utils.py:6-36)."""
import configparser
import json
from itertools import product

# keys whose values are re-typed after the generic "int if it parses, else the raw string" rule (utils.py:19-34)
FLOAT_KEYS = frozenset((
    "learning_rate", "dropout", "alpha", "beta", "beta_incr", "shift_threshold_ball", "succes_switch",
    "tolearance_to_thresh", "memory_reset_threshold", "fake_min_energy", "_true_en"))
STRING_KEYS = frozenset((
    "ham_type", "fn_type", "geometry", "method", "agent_type", "agent_class", "init_seed", "init_path",
    "init_thresh", "mapping", "optim_alg", "curriculum_type"))
JSON_KEYS = frozenset((
    "episodes", "neurons", "accept_err", "epsilon_decay", "epsilon_min", "final_gamma", "memory_clean",
    "update_target_net", "epsilon_restart", "thresholds", "switch_episodes"))


def _typed(key, raw):
    if key in FLOAT_KEYS:
        return float(raw)
    if key in STRING_KEYS:
        return str(raw)
    if key in JSON_KEYS:
        return json.loads(raw)
    try:
        return int(raw)
    except ValueError:
        return raw


def get_config(config_name, experiment_name, path="configuration_files", verbose=True):
    """Reads `<path>/<config_name><experiment_name>` (the drivers pass the experiment directory with a trailing
    slash first and `<cfg>.cfg` second, TensorRL_fixed_noiseless.py:218).  configparser lower-cases the keys."""
    parser = configparser.ConfigParser()
    parser.read("{}/{}{}".format(path, config_name, experiment_name))
    return {section: {key: _typed(key, raw) for key, raw in parser.items(section)}
            for section in parser.sections()}


def _cnot_then_rotation_actions(num_qubits, qubit_order, offset_order):
    """Action vectors [ctrl, offset, rot_qubit, rot_axis]: every CNOT (target = ctrl + offset mod n; rot_qubit = n
    marks "no rotation"), then every rotation (ctrl = n marks "no CNOT"; axis 1, 2, 3 = X, Y, Z)."""
    actions = [[c, x, num_qubits, 0] for c, x in product(qubit_order, offset_order)]
    actions += [[num_qubits, 0, r, h] for r, h in product(qubit_order, range(1, 4))]
    return actions


def dictionary_of_actions(num_qubits):
    """index -> action, ascending qubits: n(n-1) CNOT actions then 3n rotation actions (utils.py:39-57)."""
    return dict(enumerate(_cnot_then_rotation_actions(num_qubits, range(num_qubits), range(1, num_qubits))))


def dict_of_actions_revert_q(num_qubits):
    """The same action set enumerated with descending qubits and offsets (utils.py:59-78)."""
    return dict(enumerate(_cnot_then_rotation_actions(num_qubits, range(num_qubits - 1, -1, -1),
                                                      range(num_qubits - 1, 0, -1))))
