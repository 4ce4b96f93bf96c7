"""Loss plugins: fn(output_list, label_list) -> tensor (contract of /root/reference/onssen/loss/__init__.py:1-7).
SI-SNR losses (loss_e2e.py) are TasNet-only and out of scope."""
from .loss_dc import loss_dc
from .loss_chimera import loss_chimera_msa, loss_chimera_psa
from .loss_mask import loss_mask_msa, loss_mask_psa
from .loss_phase import loss_phase

__all__ = ["loss_dc", "loss_chimera_msa", "loss_chimera_psa", "loss_mask_msa", "loss_mask_psa", "loss_phase"]
