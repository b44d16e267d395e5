"""The N-API addon source (shim/napi/cueball_b200_napi.c) cannot be built here -- Node.js and its
headers are not installed -- but it can be type-checked: `gcc -fsyntax-only` against the real C ABI
header and a declarations-only stand-in for <node_api.h> (shim/napi/stub).  This catches a binding
that drifts from include/cueball_b200.h."""
import os
import shutil
import subprocess

import pytest

import util


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not on PATH")
def test_napi_shim_type_checks_against_the_c_abi():
    src = os.path.join(util.ROOT, "shim", "napi", "cueball_b200_napi.c")
    out = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror",
                          "-I", os.path.join(util.ROOT, "shim", "napi", "stub"),
                          "-I", os.path.join(util.ROOT, "include"), src], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
