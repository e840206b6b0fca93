"""B200-native drop-in for the reference's ``models`` package (toandaominh1997/EfficientDet.Pytorch).

Put this directory's parent (``efficientdet.pytorch_b200/``) in front of the reference checkout on
``sys.path`` and ``train.py`` / ``eval.py`` / ``demo.py`` import these classes unchanged:
same class names, constructor signatures, attribute names and state-dict schema; every forward /
backward runs hand-written sm_100a kernels from ``csrc/libeffdet_b200.so`` (no CPU fallback).
"""
from .efficientdet import EfficientDet  # noqa: F401  (reference: models/__init__.py:1)
