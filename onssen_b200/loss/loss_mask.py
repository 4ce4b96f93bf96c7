"""Enhancement mask losses -- drop-in for /root/reference/onssen/loss/loss_mask.py:6-40."""
from .. import _lib


def loss_mask_msa(output, label):
    [clean_est] = output
    [mag_clean, cos_diff] = label
    return _lib.loss_mse_fwd(clean_est.float().contiguous(), mag_clean.float().contiguous())


def loss_mask_psa(output, label):
    [mask] = output
    [mag_noisy, mag_clean, cos_diff] = label
    c = lambda t: t.float().contiguous()
    return _lib.loss_l1_psa_fwd(c(mask), c(mag_noisy), c(mag_clean), c(cos_diff))
