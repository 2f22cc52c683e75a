"""The backend-neutral graph expressions moved into the package (quartz_b200/graphs.py: bench.py and the scene
importer build graphs too); tests keep importing them from here."""
from quartz_b200.graphs import *  # noqa: F401,F403
from quartz_b200.graphs import ARRAY_OPS, CONNECTIVE, L, add, branch, build, bus, mul, pipe, stack, sub, thru  # noqa: F401
