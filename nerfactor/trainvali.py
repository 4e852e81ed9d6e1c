"""`python $REPO/nerfactor/trainvali.py ...` -- the reference's launch spelling
(nerfactor/trainvali_run.sh) -- runs nerfactor_b200.trainvali with the reference's flag names."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfactor_b200 import _aliases  # noqa: E402

_aliases.install(bare=True)
from nerfactor_b200.trainvali import main  # noqa: E402

if __name__ == '__main__':
    main()
