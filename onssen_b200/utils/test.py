"""tester + per-model estimators -- /root/reference/onssen/utils/test.py:7-41 and the `get_est_sig` hooks of
egs/wsj0-2mix/{deep_clustering,chimera}/evaluate.py.  mask x mixture STFT -> waveform runs on the device
(onssen_istft_masked) and so does the VAD + K-means on the active-bin embeddings (onssen_kmeans_masks: Lloyd with a
deterministic seeding instead of sklearn's RNG-driven k-means++; same partition up to the label order, which the
permutation-invariant SI-SDR ignores)."""
import itertools
import os

import torch

from .. import _lib
from .basic import AverageMeter


def batch_si_sdr(est, ref, return_perm=False):
    """`batch_SDR_torch` of onssen/evaluate/sdr.py:40-87 restated for the device (pinned by tests/golden/sdr.npz,
    generated from the live function): zero-mean both, SDR[i,j] of estimate i against reference j with the
    reference's 1e-8 regularisers (sdr.py:24-32), best sum over source permutations (sorted order, first maximum
    wins, sdr.py:71-82), divided by the number of sources.  est/ref (B,S,n) -> (B,) in the promoted dtype."""
    dt = torch.promote_types(est.dtype, ref.dtype)
    est, ref = est.to(dt), ref.to(dt)
    B, S, n = est.shape
    assert ref.shape == est.shape, "Estimation and original sources should have same shape."
    assert S < n, "Axis 1 should be the number of sources, and axis 2 should be the signal."
    est = est - est.mean(2, keepdim=True)
    ref = ref - ref.mean(2, keepdim=True)
    e, o = est[:, :, None, :], ref[:, None, :, :]                      # (B,S,1,n) x (B,1,S,n)
    origin_power = (o * o).sum(-1, keepdim=True) + 1e-8
    est_true = (o * e).sum(-1, keepdim=True) / origin_power * o
    est_res = e - est_true
    sdr = 10 * torch.log10((est_true ** 2).sum(-1) + 1e-8) - 10 * torch.log10((est_res ** 2).sum(-1) + 1e-8)   # (B,S,S)
    perms = sorted(set(itertools.permutations(range(S))))
    idx = torch.arange(S, device=est.device)
    per = torch.stack([sdr[:, idx, torch.tensor(p, device=est.device)].sum(1) for p in perms], 1)      # (B, S!)
    best, which = per.max(1)
    return (best / S, which) if return_perm else best / S


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
                B = input[0].shape[0]
                if B == 1:                                                    # the reference's loop (batch 1)
                    output = self.model(input)
                    sig_est, sig_ref = self.get_est_sig(input, label, output)
                    sdr = batch_si_sdr(sig_est, sig_ref)                      # (B,) like batch_SDR_torch
                    sdrs.update(float(sdr.mean().item()), sdr.numel())
                    continue
                # zero-padded batch of utterances of different lengths (eval_batch_size > 1): one model call with the
                # per-utterance frame counts in the recurrence, then every utterance at its own length
                ns = label[3]
                frames = 1 + ns.to(torch.int64) // self.hop_size
                self.model.frame_lengths = frames.to(torch.int32)
                try:
                    output = self.model(input)
                finally:
                    self.model.frame_lengths = None
                for b in range(B):
                    fb, nb = int(frames[b]), int(ns[b])
                    inp_b = [x[b:b + 1, :fb] for x in input]
                    lab_b = [label[0][b:b + 1, :fb], label[1][b:b + 1, :fb], label[2][b:b + 1, :, :nb]]
                    out_b = [o[b:b + 1, :fb] for o in output]
                    sig_est, sig_ref = self.get_est_sig(inp_b, lab_b, out_b)
                    sdrs.update(float(batch_si_sdr(sig_est, sig_ref).mean().item()), 1)
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
