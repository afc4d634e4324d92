"""pymes_b200 -- B200-native coupled-cluster contraction engine behind the pymes call surface.

Sub-packages mirror the reference layout (``solver``, ``mixer``, ``model``,
``integral``, ``mean_field``, ``util``) so that a pymes driver runs after changing
its imports from ``pymes`` to ``pymes_b200``.  All arithmetic on the hot path runs
in ``libpymes_b200.so`` (hand-written sm_100a CUDA behind the C ABI declared in
``include/pymes_b200.h``); there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (does not dlopen until first use)
