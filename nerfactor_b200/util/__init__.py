from . import math, img  # noqa: F401
