"""`python $REPO/nerfactor/geometry_from_nerf.py ...` -- the reference's launch spelling
(nerfactor/geometry_from_nerf_run.sh) -- runs nerfactor_b200.geometry_from_nerf with the reference's flag names."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfactor_b200 import _aliases  # noqa: E402

_aliases.install(bare=True)
from nerfactor_b200.geometry_from_nerf import main  # noqa: E402

if __name__ == '__main__':
    main()
