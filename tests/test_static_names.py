"""CPU: bench.py, __graft_entry__.py and the package's Python files only run on the GPU box - catch a name that is read
but never bound anywhere in its file (a slip of an edit) here, where there is no GPU to run them."""
import ast
import builtins
import glob
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _unbound(path):
    tree = ast.parse(open(path).read())
    bound = {"__file__", "__name__", "__doc__"}
    for n in ast.walk(tree):
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            bound.add(n.id)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            bound.update((a.asname or a.name).split(".")[0] for a in n.names)
        elif isinstance(n, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
            bound.add(n.name)
        elif isinstance(n, ast.arg):
            bound.add(n.arg)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            bound.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            bound.update(n.names)
    return sorted({(n.id, n.lineno) for n in ast.walk(tree)
                   if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in bound and not hasattr(builtins, n.id)})


def test_no_name_is_read_without_being_bound():
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] + sorted(glob.glob(os.path.join(ROOT, "bbm_b200", "*.py"))) \
        + sorted(glob.glob(os.path.join(ROOT, "tools", "*.py")))
    bad = {os.path.relpath(f, ROOT): _unbound(f) for f in files}
    bad = {k: v for k, v in bad.items() if v}
    assert not bad, bad
