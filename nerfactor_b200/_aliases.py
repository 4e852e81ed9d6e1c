"""Import-root drop-in: the reference's own import spellings resolve to this package.

The reference is launched as `python $REPO/nerfactor/trainvali.py` with PYTHONPATH=$REPO
(nerfactor/trainvali_run.sh:29-33), so its files import BOTH `nerfactor.models.X` / `brdf.renderer`
(via $REPO) and bare `models.X`, `datasets.X`, `networks`, `util`, `losses` (via $REPO/nerfactor:
models/__init__.py:19, datasets/__init__.py:19, models/base.py:17-19).  `install()` puts one
meta-path finder in front of the import system that maps every such name onto the SAME module
object as its `nerfactor_b200.*` twin (no second copy of any module, one ctypes library, one
context).  The repo-root stubs `nerfactor/` and `brdf/` call it; nothing is aliased before a
caller asks for the reference's names."""
import importlib
import importlib.abc
import importlib.util
import sys

ROOTS = {'nerfactor': 'nerfactor_b200', 'brdf': 'nerfactor_b200.brdf'}
BARE = {'models': 'nerfactor_b200.models', 'datasets': 'nerfactor_b200.datasets',
        'networks': 'nerfactor_b200.networks', 'util': 'nerfactor_b200.util',
        'losses': 'nerfactor_b200.losses'}


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, table):
        self.table = dict(table)

    def _real(self, fullname):
        head, _, rest = fullname.partition('.')
        if head not in self.table:
            return None
        return self.table[head] + ('.' + rest if rest else '')

    def find_spec(self, fullname, path=None, target=None):
        real = self._real(fullname)
        if real is None:
            return None
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except (ImportError, AttributeError, ValueError):
            return None
        spec = importlib.util.spec_from_loader(fullname, self)
        spec._nf_real = real
        return spec

    def create_module(self, spec):
        return importlib.import_module(spec._nf_real)      # the twin itself, not a copy

    def exec_module(self, module):
        pass


_finder = None


def install(bare=False):
    """Idempotent.  bare=True additionally maps the bare names the reference uses when
    $REPO/nerfactor is on sys.path (`models`, `datasets`, `networks`, `util`, `losses`)."""
    global _finder
    table = dict(ROOTS)
    if bare or (_finder is not None and any(k in _finder.table for k in BARE)):
        table.update(BARE)
    if _finder is None:
        _finder = _AliasFinder(table)
        sys.meta_path.insert(0, _finder)
    else:
        _finder.table.update(table)
    for name in list(table):
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, '__name__', name) == name and \
                not getattr(mod, '_nf_stub', False):
            continue                                  # a foreign module of that name is loaded
        sys.modules[name] = importlib.import_module(table[name])
    return _finder


def uninstall():
    global _finder
    if _finder is not None:
        if _finder in sys.meta_path:
            sys.meta_path.remove(_finder)
        for name in list(sys.modules):
            head = name.partition('.')[0]
            if head in _finder.table and not name.startswith('nerfactor_b200'):
                del sys.modules[name]
        _finder = None
