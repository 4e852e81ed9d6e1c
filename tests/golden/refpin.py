"""Build-container helper: import the REFERENCE tree (/root/reference) next to this repo.

The reference's `brdf/` and `third_party/` directories have no __init__.py (namespace packages),
and Python prefers a regular package anywhere on sys.path over a namespace portion that comes
first -- so the repo-root drop-in stub `brdf/` (regular package) would shadow the reference's own
`brdf` even with /root/reference first on sys.path.  `pin()` binds those names to the reference's
directories explicitly; `unpin()` forgets every reference module again."""
import importlib.machinery
import importlib.util
import sys

REF = '/root/reference'
_NAMES = ('brdf', 'third_party')


def pin(ref=REF):
    for name in _NAMES:
        mod = sys.modules.get(name)
        if mod is not None and any(str(p).startswith(ref) for p in getattr(mod, '__path__', [])):
            continue
        for k in [k for k in sys.modules if k == name or k.startswith(name + '.')]:
            del sys.modules[k]
        spec = importlib.machinery.PathFinder.find_spec(name, [ref])
        if spec is None:
            continue
        sys.modules[name] = importlib.util.module_from_spec(spec)


def unpin(ref=REF):
    for k in list(sys.modules):
        mod = sys.modules[k]
        f = getattr(mod, '__file__', None) or ''
        paths = [str(p) for p in getattr(mod, '__path__', [])] if hasattr(mod, '__path__') else []
        if f.startswith(ref) or any(p.startswith(ref) for p in paths):
            del sys.modules[k]
