"""enhance: BLSTM -> BatchNorm1d -> sigmoid mask, with "restoration" layers fc_pre / fc_post.

Drop-in for /root/reference/onssen/nn/enhancement.py:5-52 (state_dict keys rnn.*, bn.*, fc_mi.*, fc_pre.*,
fc_post.*; two-tensor input [x, mag_noisy]; returns [clean])."""
import torch
import torch.nn as nn

from .. import _lib
from ._blstm import PackCache, blstm_forward, bn_sync


class enhance(nn.Module):
    def __init__(self, input_dim, hidden_dim=300, num_layers=3, dropout=0.3):
        super().__init__()
        self.add_module("rnn", nn.LSTM(input_dim, hidden_dim, num_layers, dropout=dropout, bidirectional=True,
                                       batch_first=True))
        self.add_module("bn", nn.BatchNorm1d(hidden_dim * 2))
        self.add_module("fc_mi", nn.Linear(hidden_dim * 2, input_dim))
        self.add_module("fc_pre", nn.Linear(input_dim, input_dim))
        self.add_module("fc_post", nn.Linear(input_dim, input_dim))
        self.input_dim, self.hidden_dim = input_dim, hidden_dim
        self._rnn_cache, self._mi, self._pre, self._post = PackCache(), PackCache(), PackCache(), PackCache()
        self.use_tensor_cores = True

    def forward(self, input):
        assert len(input) == 2, "There must be two tensors in the input for the enhance network"
        x, mag_noisy = input
        x = x.float()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if not self.training:
                raise NotImplementedError("enhance: gradients are implemented for train() mode (batch-statistics "
                                          "BatchNorm), call model.train()")
            from ._train import enhance_backward, enhance_forward_train, run_model
            return [run_model(self, enhance_forward_train, enhance_backward, (x, mag_noisy))]
        B, T, F = x.shape
        H = self.hidden_dim
        M = T * B
        _, y_f = blstm_forward(self.rnn, self._rnn_cache, x, self.training, want_f32=True, want_f16=False,
                               use_tensor_cores=self.use_tensor_cores)
        bn = self.bn
        a_h, _, _ = _lib.bn_forward_f16(y_f, M, H, bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                        bn.running_var, bn.eps, bn.momentum, self.training, sync=bn_sync(self))
        if self.training:
            bn.num_batches_tracked += 1
        w_mi = self._mi.get([self.fc_mi.weight], lambda: _lib.pack_linear_f16(self.fc_mi.weight, True, H))
        w_pre = self._pre.get([self.fc_pre.weight], lambda: _lib.pack_linear_f16(self.fc_pre.weight, False))
        w_post = self._post.get([self.fc_post.weight], lambda: _lib.pack_linear_f16(self.fc_post.weight, False))
        dev = x.device
        mask = torch.empty(M, F, device=dev, dtype=torch.float32)          # time-major
        _lib.gemm_f16(a_h, w_mi, self.fc_mi.bias.detach(), mask, M, F, a_h.shape[1], F, epi=1)
        noisy_h = _lib.pack_input_f16(mag_noisy.float().contiguous())       # time-major fp16
        pre = torch.empty(M, F, device=dev, dtype=torch.float32)
        _lib.gemm_f16(noisy_h, w_pre, self.fc_pre.bias.detach(), pre, M, F, noisy_h.shape[1], F, epi=2)
        est_h = _lib.mul_pack_f16(pre, mask)
        clean = torch.empty(B, T, F, device=dev, dtype=torch.float32)
        _lib.gemm_f16(est_h, w_post, self.fc_post.bias.detach(), clean, M, F, est_h.shape[1], F, epi=2,
                      remap_inner=B, remap_outer=T)
        return [clean]
