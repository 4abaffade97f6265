"""combat_b200 -- B200-native (sm_100a) implementation of COMBAT's alternated generator/surrogate training step.

Importing the package loads the in-tree C-ABI library (combat_b200/libcombat_b200.so); there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises if the CUDA library is missing)

__all__ = ["_lib"]
