"""A minimal OpenQASM 2 reader for the reference's init circuits (test infrastructure): `rx/ry/rz(expr) q[i];` and
`cx q[i],q[j];` on one register, angle expressions over numbers, `pi`, + - * / and parentheses.  Written independently of
tensorrl_qas_b200.loaders so that the shipped `.qasm` twins can pin the QPY reader."""
import ast
import math
import operator
import re

_OPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
        ast.USub: operator.neg, ast.UAdd: operator.pos}


def _value(node):
    if isinstance(node, ast.Expression):
        return _value(node.body)
    if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
        return float(node.value)
    if isinstance(node, ast.Name) and node.id == "pi":
        return math.pi
    if isinstance(node, ast.BinOp) and type(node.op) in _OPS:
        return _OPS[type(node.op)](_value(node.left), _value(node.right))
    if isinstance(node, ast.UnaryOp) and type(node.op) in _OPS:
        return _OPS[type(node.op)](_value(node.operand))
    raise ValueError("unsupported angle expression")


def parse_qasm2(text):
    """-> (n_qubits, [(name, (q0,) | (q0, q1), angle | None)]) in file order."""
    n, ops = None, []
    for raw in text.split(";"):
        stmt = raw.strip()
        if not stmt or stmt.startswith("OPENQASM") or stmt.startswith("include"):
            continue
        m = re.fullmatch(r"qreg\s+(\w+)\[(\d+)\]", stmt)
        if m:
            n = int(m.group(2))
            continue
        m = re.fullmatch(r"(rx|ry|rz)\s*\((.*)\)\s*\w+\[(\d+)\]", stmt, re.S)
        if m:
            ops.append((m.group(1), (int(m.group(3)),), _value(ast.parse(m.group(2).strip(), mode="eval"))))
            continue
        m = re.fullmatch(r"cx\s+\w+\[(\d+)\]\s*,\s*\w+\[(\d+)\]", stmt)
        if m:
            ops.append(("cx", (int(m.group(1)), int(m.group(2))), None))
            continue
        raise ValueError(f"unsupported OpenQASM statement: {stmt!r}")
    return n, ops
