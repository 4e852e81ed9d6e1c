"""`python $REPO/nerfactor/test.py ...` -- the reference's launch spelling
(nerfactor/test_run.sh) -- runs nerfactor_b200.test with the reference's flag names."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfactor_b200 import _aliases  # noqa: E402

_aliases.install(bare=True)
from nerfactor_b200.test import main  # noqa: E402

if __name__ == '__main__':
    main()
