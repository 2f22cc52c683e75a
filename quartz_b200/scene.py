"""Scene (RON) importer: turns a saved quartz patch into graph expressions that `quartz_b200.graphs.build` / `Net` evaluate.

A scene is the Bevy RON dump written by `save_scene` (/root/reference/src/main.rs:324-370): per circle the components
`Op`, `Number`, `Arr`, `Order`, `Holes`; per hole a `WhiteHole{bh_parent, link_types, open}` or `BlackHole`
(/root/reference/src/components.rs:20-151).  The audio graph of a circle is re-derived exactly the way the patch
interpreter does it every frame: connective circles read their inputs in white-hole link-index order
(/root/reference/src/process.rs:1730-1734, 1811-1815), array-fed constructors read the linked circle's `Arr`
(:1453, :1466, :1655, :1677), `kr/s/reset/sr` read their own `Number` (:1559), `feedback` reads an optional delay from a
`(-1, 2)` link (:1489), `var()` holds the circle's `Number` (:1382).  Control-plane ops (targets, arrays, input,
colours...) are not audio and are ignored."""
import re

NUM = r'-?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|inf|NaN)'
CONNECTIVE = {">>": ">>", "|": "|", "&": "&", "^": "^", "+": "+", "*": "*", "-": "-", "!": "!",
              "PIP": ">>", "STA": "|", "BUS": "&", "BRA": "^", "SUM": "+", "PRO": "*", "SUB": "-", "THR": "!"}
ARRAY_OPS = {"branch()", "bus()", "pipe()", "stack()", "sum()", "product()"}


def _nums(s):
    return [float(x) for x in re.findall(NUM, s)]


class Scene:
    def __init__(self, text):
        self.circles = {}    # id -> {"op", "number", "arr", "order", "holes"}
        self.white = {}      # hole id -> (bh_parent, (lt0, lt1))
        for m in re.finditer(r'\n    (\d+): \(\n      components: \{\n(.*?)\n      \},\n    \),', text, re.S):
            eid, comp = int(m.group(1)), {}
            for cm in re.finditer(r'"quartz::components::(\w+)": (.*?)(?=\n        "|\Z)', m.group(2), re.S):
                comp[cm.group(1)] = cm.group(2).strip().rstrip(',')
            if "Op" in comp:
                op = re.match(r'\("(.*)"\)$', comp["Op"], re.S)
                self.circles[eid] = {
                    "op": op.group(1) if op else "",
                    "number": (_nums(comp.get("Number", "(0.0)")) or [0.0])[0],
                    "arr": _nums(comp.get("Arr", "")),
                    "order": int((_nums(comp.get("Order", "(0)")) or [0])[0]),
                    "holes": [int(x) for x in re.findall(r'\d+', comp.get("Holes", ""))],
                }
            elif "WhiteHole" in comp:
                w = comp["WhiteHole"]
                bhp = int(re.search(r'bh_parent:\s*(\d+)', w).group(1))
                lt = tuple(int(x) for x in re.search(r'link_types:\s*\(\s*(-?\d+),\s*(-?\d+)', w).groups())
                self.white[eid] = (bhp, lt)

    @staticmethod
    def load(path):
        return Scene(open(path).read())

    def inputs_of(self, eid):
        """[(link_types, parent circle id)] for every white hole on circle `eid`"""
        return [(self.white[h][1], self.white[h][0]) for h in self.circles[eid]["holes"] if h in self.white]

    def find(self, op):
        return [e for e, c in self.circles.items() if c["op"].replace(" ", "") == op]

    def _net_input(self, eid, lt=(0, 1)):
        for l, p in self.inputs_of(eid):
            if l == lt and p in self.circles:
                return p
        return None

    def _arr_input(self, eid):
        for l, p in self.inputs_of(eid):
            if l == (-13, 1) and p in self.circles:
                return self.circles[p]["arr"]
        return None

    def expr(self, eid, depth=0):
        """graph expression for the Net held by circle `eid` (None when the circle holds no audio graph)"""
        if depth > 64 or eid not in self.circles:
            return None
        c = self.circles[eid]
        op = c["op"].replace(" ", "")
        sub = lambda p: self.expr(p, depth + 1) if p is not None else None   # noqa: E731
        if op in CONNECTIVE:
            slots = {}
            for (l, p) in self.inputs_of(eid):
                if l[0] == 0 and p in self.circles:
                    slots[max(l[1], 0)] = p
            kids = [sub(slots[k]) for k in sorted(slots)]
            kids = [k for k in kids if k is not None]
            if CONNECTIVE[op] == "-":
                lhs, rhs = sub(self._net_input(eid, (0, 1))), sub(self._net_input(eid, (0, 2)))
                kids = [k for k in (lhs, rhs) if k is not None]
            return {"op": CONNECTIVE[op], "n": c["number"], "inputs": kids}
        if op in ("quantize()", "get()", "wave()"):
            arr = self._arr_input(eid)
            return {"op": op, "arr": arr} if arr else None
        if op == "feedback()":
            net = sub(self._net_input(eid))
            delay = None
            for l, p in self.inputs_of(eid):
                if l == (-1, 2) and p in self.circles:
                    delay = self.circles[p]["number"]
            return {"op": op, "net": net, "delay": delay} if net else None
        if op in ("kr()", "s()", "reset()", "sr()"):
            net = sub(self._net_input(eid))
            return {"op": op, "net": net, "n": c["number"]} if net else None
        if op in ("trig_reset()", "reset_v()"):
            net = sub(self._net_input(eid))
            return {"op": op, "net": net} if net else None
        if op in ("seq()", "select()"):
            slots = {}
            for (l, p) in self.inputs_of(eid):
                if l[0] == 0 and p in self.circles:
                    slots[max(l[1], 0)] = p
            kids = [sub(slots[k]) for k in sorted(slots)]
            return {"op": op, "inputs": [k for k in kids if k is not None]}
        if op in ARRAY_OPS:
            arr = self._arr_input(eid)
            src = self._net_input(eid, (0, 2))
            if arr is None or src is None:
                return None
            return {"op": op, "str": self.circles[src]["op"], "arr": arr}
        if op == "var()":
            return {"op": "var()", "n": c["number"]}
        if op in ("in()", "adc()", "monitor()", "buffin()", "buffout()"):
            return {"op": "monitor()" if op == "monitor()" else op}
        if op.startswith("swap("):
            return sub(self._net_input(eid))          # SwapUnit forwards to the net it was given (nodes.rs:548-555)
        if op in ("out()", "dac()"):
            return sub(self._net_input(eid))
        if op in ("render", "apply"):
            return sub(self._net_input(eid))
        if "(" in op and op.endswith(")"):
            return {"op": c["op"]}
        return None

    def unsupported_ops(self, e, B):
        """names of ops in expression `e` that backend B cannot lower"""
        bad = set()

        def walk(x):
            if x is None:
                return
            for k in x.get("inputs", []):
                walk(k)
            if "net" in x:
                walk(x["net"])
            if x["op"] not in CONNECTIVE and "(" in x["op"] and "inputs" not in x and "net" not in x and "arr" not in x \
                    and x["op"] not in ("var()", "in()", "adc()", "monitor()", "buffin()", "buffout()"):
                try:
                    n = B.str_to_net(x["op"])
                    u = n.unsupported() if hasattr(n, "unsupported") else None
                    if u:
                        bad.add(u)
                except Exception:
                    bad.add(x["op"].split("(")[0])
        walk(e)
        return sorted(bad)
