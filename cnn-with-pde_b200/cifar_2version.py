"""Drop-in for cifar_2version.LearnableDiffusionLayer (cifar_2version.py:20-187)."""
from .cifar10 import EnhancedDiffusionLayer as _Enhanced


class LearnableDiffusionLayer(_Enhanced):
    """Same parameters as the cifar10 layer but Lie splitting: x(dt/2) then y(dt/2), no closing
    x half-sweep (cifar_2version.py:93,99)."""

    _lie = True


from .classifiers import (CIFAR10HybridPDEModel, HamiltonianBlock, HybridPDEExtractor, NonConvSpatialAttention,  # noqa: E402,F401
                          ParabolicBlock, SymmetricLayer)
from .classifiers import HybridFC as PDEClassifier  # noqa: E402,F401  (cifar_2version.py:336-372)
