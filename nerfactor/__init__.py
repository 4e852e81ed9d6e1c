"""Drop-in import root: `import nerfactor.models.nerfactor`, `from nerfactor.networks import mlp`,
... resolve to the B200-native package `nerfactor_b200` (same module objects).  When this directory
itself is on sys.path -- the reference's launch convention `python $REPO/nerfactor/trainvali.py`,
nerfactor/trainvali_run.sh:29-33 -- the bare spellings `models.X`, `datasets.X`, `networks`,
`util`, `losses` resolve as well."""
import os as _os
import sys as _sys

_nf_stub = True
from nerfactor_b200 import _aliases as _aliases  # noqa: E402

_here = _os.path.dirname(_os.path.abspath(__file__))
_aliases.install(bare=any(_os.path.abspath(p or '.') == _here for p in _sys.path))
