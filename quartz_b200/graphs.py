"""Backend-neutral graph expressions.  A graph is a nested dict (the shape tests/golden/make_golden.py emits):
  {"op": "sine(440)"}                                    leaf -> str_to_net
  {"op": ">>", "n": 0, "inputs": [...]}                 connective circle (process.rs:1719-1876)
  {"op": "quantize()", "arr": [...]} / get() / wave()   array-fed constructors (process.rs:1450-1477, 1652-1667)
  {"op": "feedback()", "net": g, "delay": s|None}       process.rs:1479-1515
  {"op": "kr()"|"s()"|"reset()"|"sr()", "net": g, "n": x}   process.rs:1540-1580
  {"op": "trig_reset()"|"reset_v()", "net": g}          process.rs:1582-1613
  {"op": "seq()"|"select()", "inputs": [...]}           process.rs:1615-1650
  {"op": "branch()"|..., "str": "sine(#)", "arr": [...]}    process.rs:1669-1717
`build(expr, B)` evaluates it with backend class B (tests.oracle_ffi.ONet or quartz_b200.Net)."""

CONNECTIVE = {">>": ">>", "|": "|", "&": "&", "^": "^", "+": "+", "*": "*", "-": "-", "!": "!",
              "PIP": ">>", "STA": "|", "BUS": "&", "BRA": "^", "SUM": "+", "PRO": "*", "SUB": "-", "THR": "!"}
ARRAY_OPS = {"branch()", "bus()", "pipe()", "stack()", "sum()", "product()"}


def build(e, B, node_limit=500):
    op = e["op"]
    if op in CONNECTIVE:
        kids = [build(k, B, node_limit) for k in e.get("inputs", [])]
        return B.connect(CONNECTIVE[op], kids, e.get("n", 0.0), node_limit)
    if op == "quantize()":
        return B.quantize(e["arr"])
    if op == "get()":
        return B.get(e["arr"])
    if op == "wave()":
        return B.wave(e["arr"])
    if op == "feedback()":
        return B.feedback(build(e["net"], B, node_limit), e.get("delay"))
    if op == "kr()":
        return B.kr(build(e["net"], B, node_limit), e["n"], False)
    if op == "s()":
        return B.kr(build(e["net"], B, node_limit), e["n"], True)
    if op == "reset()":
        return B.reset_every(build(e["net"], B, node_limit), e["n"])
    if op == "sr()":
        return build(e["net"], B, node_limit).set_sample_rate(e["n"])
    if op == "trig_reset()":
        return B.trig_reset(build(e["net"], B, node_limit))
    if op == "reset_v()":
        return B.reset_v(build(e["net"], B, node_limit))
    if op == "seq()":
        return B.seq([build(k, B, node_limit) for k in e["inputs"]])
    if op == "select()":
        return B.select([build(k, B, node_limit) for k in e["inputs"]])
    if op in ARRAY_OPS:
        return B.array_op(op, e["str"], e["arr"])
    if op in ("in()", "adc()", "buffin()", "buffout()", "monitor()"):
        return B.live_io(op)
    if op == "var()":
        return B.var(e.get("n", 0.0))
    return B.str_to_net(op)


def L(op):
    return {"op": op}


def pipe(*xs, n=0.0):
    return {"op": ">>", "n": n, "inputs": [L(x) if isinstance(x, str) else x for x in xs]}


def stack(*xs, n=0.0):
    return {"op": "|", "n": n, "inputs": [L(x) if isinstance(x, str) else x for x in xs]}


def branch(*xs, n=0.0):
    return {"op": "^", "n": n, "inputs": [L(x) if isinstance(x, str) else x for x in xs]}


def bus(*xs, n=0.0):
    return {"op": "&", "n": n, "inputs": [L(x) if isinstance(x, str) else x for x in xs]}


def add(*xs, n=0.0):
    return {"op": "+", "n": n, "inputs": [L(x) if isinstance(x, str) else x for x in xs]}


def mul(*xs, n=0.0):
    return {"op": "*", "n": n, "inputs": [L(x) if isinstance(x, str) else x for x in xs]}


def sub(a, b):
    return {"op": "-", "n": 0.0, "inputs": [L(x) if isinstance(x, str) else x for x in (a, b)]}


def thru(a):
    return {"op": "!", "n": 0.0, "inputs": [L(a) if isinstance(a, str) else a]}
