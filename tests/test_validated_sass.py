"""Which device code has actually run on hardware: the SASS of every kernel in the current build is
compared with the hashes recorded at the last GPU-validated commit (profiles/validated_sass.json).
A kernel listed there must be byte-identical (refactors of host code, comments or test hooks must
not perturb it); kernels not listed are reported as not yet validated on hardware."""
import json
import os
import shutil
import sys

import pytest

import util

sys.path.insert(0, os.path.join(util.ROOT, "tools"))
import sass_hashes  # noqa: E402

SO = os.path.join(util.ROOT, "node-cueball_b200", "libcueball_b200.so")
UNVALIDATED_OK = ()  # kernels allowed to exist without a hardware run (none: every kernel of the build is pinned)


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_validated_kernels_are_byte_identical():
    if not os.path.exists(SO):
        import __graft_entry__ as g
        g.build()
    with open(os.path.join(util.ROOT, "profiles", "validated_sass.json")) as f:
        want = json.load(f)["kernels"]
    got = sass_hashes.kernel_hashes(SO)
    changed = [k for k in want if got.get(k) != want[k]]
    assert not changed, "device code changed since the last hardware validation: %s" % changed
    new = [k for k in got if k not in want]
    assert all(any(u in k for u in UNVALIDATED_OK) for k in new), "unexpected unvalidated kernels: %s" % new
