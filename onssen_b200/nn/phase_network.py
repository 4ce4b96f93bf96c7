"""phase_net: chimera + a second BLSTM (run once per speaker) + BN + phase head.

REPAIRED restatement of /root/reference/onssen/nn/phase_network.py:7-67, which is not constructible as written
(`output_dim` undefined at :28).  Repair: fc_phase = Linear(2H, num_speaker*input_dim) so the
reshape at :55,63 yields (B,T,F,2) and can be added to x_phase.  State_dict keys: chimera.*, rnn.*, bn.*,
fc_phase.*.  No reference oracle exists for this model; parity is checked against the oracle's repaired
restatement only."""
import torch
import torch.nn as nn

from .. import _lib
from ._blstm import PackCache, bn_sync
from .chimera import chimera


class phase_net(nn.Module):
    def __init__(self, input_dim, hidden_dim=300, num_layers=3, embedding_dim=20, dropout=0.3, num_speaker=2):
        super().__init__()
        assert num_speaker == 2, "the (re,im) residual add of phase_network.py:63-64 needs a last dim of 2"
        self.add_module("rnn", nn.LSTM(input_dim * 3, hidden_dim, num_layers, dropout=dropout, bidirectional=True,
                                       batch_first=True))
        self.add_module("bn", nn.BatchNorm1d(hidden_dim * 2))
        self.add_module("fc_phase", nn.Linear(hidden_dim * 2, num_speaker * input_dim))
        self.add_module("chimera", chimera(input_dim, hidden_dim, num_layers, embedding_dim, dropout, num_speaker))
        self.input_dim, self.hidden_dim = input_dim, hidden_dim
        self._rnn_cache, self._ph = PackCache(), PackCache()

    def _branch(self, xin_h, B, T, x_phase):
        """second BLSTM on a pre-packed time-major fp16 input -> BN -> fc_phase -> + x_phase -> normalise"""
        from ._blstm import blstm_forward_packed
        H, F = self.hidden_dim, self.input_dim
        M = T * B
        _, y_f = blstm_forward_packed(self.rnn, self._rnn_cache, xin_h, B, T, self.training, True, False)
        bn = self.bn
        a_h, _, _ = _lib.bn_forward_f16(y_f, M, H, bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                        bn.running_var, bn.eps, bn.momentum, self.training, sync=bn_sync(self))
        if self.training:
            bn.num_batches_tracked += 1
        w = self._ph.get([self.fc_phase.weight], lambda: _lib.pack_linear_f16(self.fc_phase.weight, True, H))
        ph = torch.empty(B, T, F, 2, device=xin_h.device, dtype=torch.float32)
        _lib.gemm_f16(a_h, w, self.fc_phase.bias.detach(), ph, M, 2 * F, a_h.shape[1], 2 * F, remap_inner=B,
                      remap_outer=T)
        return _lib.add_l2norm_pairs(ph, x_phase)

    def forward(self, input):
        assert len(input) == 2, "There must be 2 tensors in the input for phase network"
        x_mag, x_phase = input
        x_mag = x_mag.float().contiguous()
        x_phase = x_phase.float().contiguous()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if not self.training:
                raise NotImplementedError("phase_net: gradients are implemented for train() mode (batch-statistics "
                                          "BatchNorm), call model.train()")
            from ._train import phase_backward, phase_forward_train, run_model
            emb, masks, p_a, p_b = run_model(self, phase_forward_train, phase_backward, (x_mag, x_phase))
            return [emb, masks[:, :, :, 0], masks[:, :, :, 1], p_a, p_b]
        embedding, mask_A, mask_B = self.chimera([x_mag])
        B, T, F = mask_A.shape
        outs = []
        for mk in (mask_A, mask_B):
            xin = _lib.pack_phase_input_f16(x_mag, mk, mk.stride(-1), x_phase)
            outs.append(self._branch(xin, B, T, x_phase))
        return [embedding, mask_A, mask_B, outs[0], outs[1]]
