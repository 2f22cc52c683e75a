"""Regenerates tests/golden/oracle_exact_digest.json: SHA-1 of the oracle's output for every bit-exact render case of
tests/cases.py (integer / trigger / table / IEEE add-mul-div work: no libm transcendentals, so the digests do not depend on
the host).  The oracle defines parity for the GPU path; this fixture makes any change of its behaviour visible in the CPU
suite.  Run from the repo root:  python tests/golden/make_oracle_digest.py"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

from tests import cases  # noqa: E402
from tests.graphs import build  # noqa: E402
from tests.oracle_ffi import ONet  # noqa: E402


def digest(expr, n):
    out = np.ascontiguousarray(build(expr, ONet).set_salt(7).render(n), dtype=np.float32)
    return hashlib.sha1(out.tobytes()).hexdigest(), list(out.shape)


if __name__ == "__main__":
    table = {}
    for name, expr, n, tol in cases.RENDER:
        if tol == "exact":
            h, shape = digest(expr, n)
            table[name] = {"sha1": h, "shape": shape}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_exact_digest.json")
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print(len(table), "cases ->", path)
