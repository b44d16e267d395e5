"""Static guard for the scripts that only run on the GPU box (bench.py, __graft_entry__.py, the
package, the tools): every name a function reads is defined somewhere it can see.  A NameError in
bench.py would otherwise surface only at round end."""
import ast
import builtins
import glob
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted([os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] +
               glob.glob(os.path.join(ROOT, "node-cueball_b200", "*.py")) +
               glob.glob(os.path.join(ROOT, "tools", "*.py")) + [os.path.join(ROOT, "tests", "ring_check.py")])


def _stores(node):
    out = set()
    for n in ast.walk(node):
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            out.add(n.id)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                out.add((a.asname or a.name).split(".")[0])
        elif isinstance(n, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
            out.add(n.name)
            if not isinstance(n, ast.ClassDef):
                a = n.args
                for x in a.args + a.kwonlyargs + a.posonlyargs + [y for y in (a.vararg, a.kwarg) if y]:
                    out.add(x.arg)
        elif isinstance(n, ast.Lambda):
            a = n.args
            for x in a.args + a.kwonlyargs + [y for y in (a.vararg, a.kwarg) if y]:
                out.add(x.arg)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            out.add(n.name)
    return out


@pytest.mark.parametrize("path", FILES, ids=[os.path.relpath(p, ROOT) for p in FILES])
def test_no_undefined_names(path):
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    star = any(isinstance(n, ast.ImportFrom) and any(a.name == "*" for a in n.names) for n in ast.walk(tree))
    if star:
        pytest.skip("star import")
    known = _stores(tree) | set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    missing = sorted({n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)} - known)
    assert not missing, "undefined names: %s" % missing
