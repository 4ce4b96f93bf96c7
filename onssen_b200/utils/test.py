"""tester + per-model estimators -- /root/reference/onssen/utils/test.py:7-41 and the `get_est_sig` hooks of
egs/wsj0-2mix/{deep_clustering,chimera}/evaluate.py.  mask x mixture STFT -> waveform runs on the device
(onssen_istft_masked) and so does the VAD + K-means on the active-bin embeddings (onssen_kmeans_masks: Lloyd with a
deterministic seeding instead of sklearn's RNG-driven k-means++; same partition up to the label order, which the
permutation-invariant SI-SDR ignores).  SI-SDR is a few lines of torch (sdr.py is out of scope)."""
import itertools
import os

import torch

from .. import _lib
from .basic import AverageMeter


def batch_si_sdr(est, ref):
    """Permutation-invariant SI-SDR in dB (behaviour of onssen/evaluate/sdr.py:40-87). est/ref (B,S,n)."""
    est = est.double() - est.double().mean(-1, keepdim=True)
    ref = ref.double() - ref.double().mean(-1, keepdim=True)
    S = est.shape[1]
    best = None
    for perm in itertools.permutations(range(S)):
        e = est[:, list(perm)]
        proj = (e * ref).sum(-1, keepdim=True) / (ref * ref).sum(-1, keepdim=True).clamp_min(1e-12) * ref
        sdr = 10 * torch.log10((proj ** 2).sum(-1) / ((e - proj) ** 2).sum(-1).clamp_min(1e-12))
        val = sdr.mean(1)
        best = val if best is None else torch.maximum(best, val)
    return float(best.mean().item())


class tester:
    def __init__(self, args):
        self.model_name = args["model_name"]
        self.test_loader = args["test_loader"]
        self.device = args["device"]
        self.model = args["model"]
        self.hop_size = args["feature_options"]["hop_size"] if "feature_options" in args else 64
        self.window_size = args["feature_options"]["window_size"] if "feature_options" in args else 256
        saved = torch.load(os.path.join(args["checkpoint_path"], "final.mdl"), weights_only=False)
        self.model.load_state_dict(saved["model"])
        self.model = self.model.to(self.device)

    def get_est_sig(self, input, label, output):
        raise NotImplementedError

    def masked_istft(self, stft_r, stft_i, masks, nsample):
        """masks (B,S,frames,F) on the device -> (B,S,nsample) (evaluate.py:42-45)."""
        return _lib.istft_masked(stft_r.float().contiguous(), stft_i.float().contiguous(), masks.float().contiguous(),
                                 self.window_size, self.hop_size, nsample)

    def eval(self):
        sdrs = AverageMeter()
        self.model = self.model.eval()
        with torch.no_grad():
            for input, label in self.test_loader:
                output = self.model(input)
                sig_est, sig_ref = self.get_est_sig(input, label, output)
                sdrs.update(batch_si_sdr(sig_est, sig_ref))
        return sdrs.avg


class tester_dc(tester):
    """egs/wsj0-2mix/deep_clustering/evaluate.py:10-47"""

    def get_est_sig(self, input, label, output):
        feature_mix, = input
        embedding, = output
        stft_r, stft_i, sig_ref = label
        num_spk, nsample = sig_ref.shape[1], sig_ref.shape[2]
        # evaluate.py:34-41 (batch 1): VAD at max - 40/20, cluster the active embeddings, mask[0] = label,
        # mask[1] = 1 - label -- one C-ABI call, nothing leaves the device
        masks = _lib.kmeans_masks(embedding[0].float().contiguous(), feature_mix[0].float().contiguous(), num_spk, 40.0)
        return self.masked_istft(stft_r, stft_i, masks.unsqueeze(0), nsample), sig_ref


class tester_chimera(tester):
    """egs/wsj0-2mix/chimera/evaluate.py:13-46"""

    def get_est_sig(self, input, label, output):
        _, mask_A, mask_B = output
        stft_r, stft_i, sig_ref = label
        masks = torch.stack([mask_A[0], mask_B[0]], 0).unsqueeze(0)
        return self.masked_istft(stft_r, stft_i, masks, sig_ref.shape[2]), sig_ref
