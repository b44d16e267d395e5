"""oracle/tracegen.py -- the trace generator lives in the package (cueball_b200.tracegen) so
bench.py and tools/ can make inputs without importing anything from oracle/; the test
infrastructure keeps importing it under this name."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cueball_b200  # noqa: E402,F401  (registers the hyphenated package directory)
from cueball_b200.tracegen import *  # noqa: E402,F401,F403
from cueball_b200.tracegen import SEED, M32, _hash32_np  # noqa: E402,F401
