#!/usr/bin/env python3
"""Extract the known-answer vectors that the reference ships inside its example scenes.

quartz has no unit tests; the only pinned numbers are arrays that were saved inside
`render` / `apply` circles of the Bevy RON scenes under /root/reference/assets.  This script walks the
scene graph (Op / Number / Arr / Holes / WhiteHole{bh_parent, link_types}; format per
/root/reference/src/main.rs:330-353 and src/components.rs:20-151), re-creates the audio-graph expression
that fed each circle (ordered by white-hole link index exactly like src/process.rs:1730-1734,
1811-1815) and writes tests/golden/quartz_assets.json.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
"""
import json, os, re, struct, sys

REF = "/root/reference/assets"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quartz_assets.json")

NUM = r'-?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|inf|NaN)'


def parse_scene(path):
    txt = open(path).read()
    ents = {}
    for m in re.finditer(r'\n    (\d+): \(\n      components: \{\n(.*?)\n      \},\n    \),', txt, re.S):
        comp = {}
        for cm in re.finditer(r'"quartz::components::(\w+)": (.*?)(?=\n        "|\Z)', m.group(2), re.S):
            comp[cm.group(1)] = cm.group(2).strip().rstrip(',')
        ents[int(m.group(1))] = comp
    return ents


def nums(s):
    return [float(x) for x in re.findall(NUM, s)]


def op_of(c):
    return re.match(r'\("(.*)"\)$', c['Op'], re.S).group(1)


def number_of(c):
    return nums(c.get('Number', '(0.0)'))[0]


def inputs_of(ents, eid):
    """[(link_types, parent id)] for every white hole on circle `eid`."""
    res = []
    for h in [int(x) for x in re.findall(r'\d+', ents[eid].get('Holes', ''))]:
        w = ents.get(h, {}).get('WhiteHole')
        if w:
            bhp = int(re.search(r'bh_parent:\s*(\d+)', w).group(1))
            lt = tuple(int(x) for x in re.search(r'link_types:\s*\(\s*(-?\d+),\s*(-?\d+)', w).groups())
            res.append((lt, bhp))
    return res


CONNECTIVE = {">>", "|", "&", "^", "+", "*", "-", "!", "PIP", "STA", "BUS", "BRA", "SUM", "PRO", "SUB", "THR"}


def expr(ents, eid, depth=0):
    """Audio-graph expression tree for the Net held by circle `eid`."""
    c = ents[eid]
    op = op_of(c)
    if depth > 32:
        raise RuntimeError("cycle")
    ins = inputs_of(ents, eid)
    if op in CONNECTIVE:
        # process.rs:1730-1734 / 1811-1815: inputs[max(lt.1,0)] = parent, for link_types.0 == 0
        slots = {}
        for (lt, p) in ins:
            if lt[0] == 0:
                slots[max(lt[1], 0)] = p
        kids = [expr(ents, slots[k], depth + 1) for k in sorted(slots)]
        return {"op": op, "n": number_of(c), "inputs": kids}
    if op in ("quantize()", "get()", "wave()"):
        arr = None
        for (lt, p) in ins:
            if lt == (-13, 1):
                arr = nums(ents[p].get('Arr', ''))
        return {"op": op, "arr": arr}
    return {"op": op}


def f32hex(x):
    return struct.unpack('<I', struct.pack('<f', x))[0]


def main():
    gold = {"source": "syther-labs/quartz assets (scene-embedded arrays)", "cases": []}

    # ---- apply circles: one Net::tick frame (process.rs:1311-1330)
    for scene in ("wip", "curve", "distance"):
        ents = parse_scene(os.path.join(REF, scene))
        for eid, c in ents.items():
            if 'Op' not in c or op_of(c) != "apply":
                continue
            net = inp = None
            for (lt, p) in inputs_of(ents, eid):
                if lt == (0, 1):
                    net = expr(ents, p)
                if lt == (-13, 2):
                    inp = nums(ents[p].get('Arr', ''))
            out = nums(c.get('Arr', ''))
            gold["cases"].append({"name": f"{scene}:apply:{eid}", "kind": "apply", "asset": f"assets/{scene}",
                                  "net": net, "input": inp, "output": out,
                                  "output_bits": [f32hex(v) for v in out]})

    # ---- render circle: 512-point Hann window (process.rs:1332-1357).  The generating graph is left in the
    # scene next to the comment circle "this mess here is how i generated the window function"; the render
    # circle itself was disconnected before saving, so the chain is re-assembled from the `>>` circle that
    # owns dc(44100).
    for scene in ("spectral-gate", "spectral-delay"):
        ents = parse_scene(os.path.join(REF, scene))
        render = [e for e, c in ents.items() if 'Op' in c and op_of(c) == "render"
                  and number_of(c) == 512.0 and len(nums(c.get('Arr', ''))) == 512]
        if not render:
            continue
        arr = nums(ents[render[0]]['Arr'])
        chain = None
        for eid, c in ents.items():
            if 'Op' in c and op_of(c) == ">>":
                e = expr(ents, eid)
                flat = json.dumps(e)
                if "dc(44100)" in flat and "ramp()" in flat and "cos()" in flat:
                    if chain is None or len(flat) > len(json.dumps(chain)):
                        chain = e
        gold["cases"].append({"name": f"{scene}:render:hann512", "kind": "render", "asset": f"assets/{scene}",
                              "net": chain, "sample_rate": 44100.0, "len": 512, "output": arr,
                              "output_bits": [f32hex(v) for v in arr]})
    json.dump(gold, open(OUT, "w"), indent=1)
    print("wrote", OUT, len(gold["cases"]), "cases")
    for c in gold["cases"]:
        print(c["name"], json.dumps(c["net"])[:300])


if __name__ == "__main__":
    main()


def scene_graphs():
    """out() graphs of the example scenes that only use ops with a GPU lowering -> tests/golden/scenes/out_graphs.json
    (produced by quartz_b200/scene.py, the importer the product ships)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from quartz_b200.scene import Scene
    res = {}
    for name in ["ar", "circular-scope", "pd", "scope", "shift_reg", "spectral-delay", "spectral-gate"]:
        sc = Scene.load(os.path.join(REF, name))
        res[name] = sc.expr((sc.find("out()") + sc.find("dac()"))[0])
    os.makedirs(os.path.join(os.path.dirname(OUT), "scenes"), exist_ok=True)
    json.dump(res, open(os.path.join(os.path.dirname(OUT), "scenes", "out_graphs.json"), "w"))


if __name__ == "__main__":
    scene_graphs()
