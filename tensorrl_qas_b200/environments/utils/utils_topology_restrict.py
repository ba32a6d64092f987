"""Restricted-connectivity ("hexagon") action dictionaries (reference: environments/utils/utils_topology_restrict.py).

Reference quirks kept on purpose: the connectivity filter compares (ctrl, target) of EVERY action
with the edge list, and a rotation action has ctrl = n, so all rotation actions are dropped; the forward and the
"reverted" n = 8 edge lists differ (the forward one holds both directions of each edge)."""
from .utils import _cnot_then_rotation_actions, get_config  # noqa: F401  (the reference module re-exports get_config)

_EDGES_FORWARD = {
    6: [(0, 1), (0, 2), (0, 3), (3, 4), (4, 5)],
    8: [(0, 1), (1, 0), (0, 2), (2, 0), (0, 3), (3, 0), (3, 4), (4, 3), (4, 5), (5, 4), (4, 6), (6, 4), (6, 7),
        (7, 6)],
    10: [(0, 1), (0, 2), (0, 3), (3, 4), (4, 5), (4, 6), (6, 7), (7, 8), (7, 9)],
}
_EDGES_REVERTED = dict(_EDGES_FORWARD)
_EDGES_REVERTED[8] = [(0, 1), (0, 2), (0, 3), (3, 4), (4, 5), (4, 6), (6, 7)]


def _restrict(actions, num_qubits, edges):
    """Keep the actions whose (ctrl, (ctrl + offset) mod n) is an edge; keys count DOWN in enumeration order
    (utils_topology_restrict.py:64-80)."""
    if num_qubits not in edges:
        raise UnboundLocalError(f"no hexagon connectivity defined for {num_qubits} qubits")  # NameError-like in the reference
    allowed = set(edges[num_qubits])
    kept = [a for a in actions if (a[0], (a[0] + a[1]) % num_qubits) in allowed]
    return {len(kept) - 1 - i: a for i, a in enumerate(kept)}


def dictionary_of_actions_hexagon_connectivity(num_qubits):
    """reference: utils_topology_restrict.py:40-80"""
    acts = _cnot_then_rotation_actions(num_qubits, range(num_qubits), range(1, num_qubits))
    return _restrict(acts, num_qubits, _EDGES_FORWARD)


def dictionary_of_actions_hexagon_connectivity_reverted(num_qubits):
    """reference: utils_topology_restrict.py:83-125"""
    acts = _cnot_then_rotation_actions(num_qubits, range(num_qubits - 1, -1, -1), range(num_qubits - 1, 0, -1))
    return _restrict(acts, num_qubits, _EDGES_REVERTED)
