"""deep_clustering: BLSTM -> BatchNorm1d -> Linear -> unit-norm embedding, on sm_100a kernels.

Drop-in for /root/reference/onssen/nn/deep_clustering.py:5-43 (same ctor kwargs, list-in/list-out forward,
same submodule names => same state_dict keys: rnn.*, bn.*, fc_dc.*).
"""
import torch
import torch.nn as nn

from .. import _lib
from ._blstm import PackCache, blstm_forward, bn_sync, eval_lengths


class deep_clustering(nn.Module):
    def __init__(self, input_dim, hidden_dim=300, num_layers=3, embedding_dim=20, dropout=0.3):
        super().__init__()
        # parameter containers only (never called): identical names/shapes/init to the reference
        self.add_module("rnn", nn.LSTM(input_dim, hidden_dim, num_layers, dropout=dropout, bidirectional=True,
                                       batch_first=True))
        self.add_module("bn", nn.BatchNorm1d(hidden_dim * 2))
        self.add_module("fc_dc", nn.Linear(hidden_dim * 2, embedding_dim * input_dim))
        self.input_dim, self.hidden_dim, self.embedding_dim = input_dim, hidden_dim, embedding_dim
        self._rnn_cache, self._fc_cache = PackCache(), PackCache()
        self.use_tensor_cores = True

    def forward(self, input):
        assert len(input) == 1, "There must be one tensor in the input for the deep clustering model"
        x = input[0].float()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if not self.training:
                raise NotImplementedError("deep_clustering: gradients are implemented for train() mode "
                                          "(batch-statistics BatchNorm), call model.train()")
            from ._train import dc_backward, dc_forward_train, run_model
            return [run_model(self, dc_forward_train, dc_backward, x)]
        B, T, F = x.shape
        H, D = self.hidden_dim, self.embedding_dim
        M = T * B
        _, y_f = blstm_forward(self.rnn, self._rnn_cache, x, self.training, want_f32=True, want_f16=False,
                               use_tensor_cores=self.use_tensor_cores, lengths=eval_lengths(self, B, x.device))
        bn = self.bn
        a_h, _, _ = _lib.bn_forward_f16(y_f, M, H, bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                        bn.running_var, bn.eps, bn.momentum, self.training, sync=bn_sync(self))
        if self.training:
            bn.num_batches_tracked += 1
        w_p = self._fc_cache.get([self.fc_dc.weight], lambda: _lib.pack_linear_f16(self.fc_dc.weight, True, H))
        emb = torch.empty(B, T, F, D, device=x.device, dtype=torch.float32)
        if not _lib.gemm_l2norm_supported(D):
            raise _lib.OnssenB200Error(f"embedding_dim={D}: fused normalise epilogue supports 4/8/12/16/20/24/32/40")
        _lib.gemm_f16(a_h, w_p, self.fc_dc.bias.detach(), emb, M, F * D, a_h.shape[1], F * D, epi=3, group=D,
                      remap_inner=B, remap_outer=T)
        return [emb]
