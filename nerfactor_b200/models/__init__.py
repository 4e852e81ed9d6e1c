"""Mirror of nerfactor/models/__init__.py:18-20 (name -> class registry)."""
from importlib import import_module


def get_model_class(name):
    return import_module('nerfactor_b200.models.' + name).Model
