"""Model plugins: nn.Module taking a list of tensors and returning a list of tensors
(contract of /root/reference/onssen/nn/__init__.py:1-5). ConvTasNet is out of scope."""
from .chimera import chimera
from .deep_clustering import deep_clustering
from .enhancement import enhance
from .phase_network import phase_net

__all__ = ["chimera", "deep_clustering", "enhance", "phase_net"]
