"""Loss plugins: fn(output_list, label_list) -> tensor (contract of /root/reference/onssen/loss/__init__.py:1-7)."""
from .loss_dc import loss_dc
from .loss_chimera import loss_chimera_msa, loss_chimera_psa

__all__ = ["loss_dc", "loss_chimera_msa", "loss_chimera_psa"]
