"""Model plugins: nn.Module taking a list of tensors and returning a list of tensors
(contract of /root/reference/onssen/nn/__init__.py:1-5)."""
from .chimera import chimera
from .deep_clustering import deep_clustering

__all__ = ["chimera", "deep_clustering"]
