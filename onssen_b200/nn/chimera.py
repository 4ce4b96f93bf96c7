"""chimera / chimera++: BLSTM -> {Linear + unit-norm embedding, Linear + sigmoid masks}.

Drop-in for /root/reference/onssen/nn/chimera.py:5-46 (state_dict keys rnn.*, fc_dc.*, fc_mi.*).
"""
import torch
import torch.nn as nn

from .. import _lib
from ._blstm import eval_lengths, PackCache, blstm_forward


class chimera(nn.Module):
    def __init__(self, input_dim, hidden_dim=300, num_layers=3, embedding_dim=20, dropout=0.3, num_speaker=2):
        super().__init__()
        self.add_module("rnn", nn.LSTM(input_dim, hidden_dim, num_layers, dropout=dropout, bidirectional=True,
                                       batch_first=True))
        self.add_module("fc_dc", nn.Linear(hidden_dim * 2, input_dim * embedding_dim))
        self.add_module("fc_mi", nn.Linear(hidden_dim * 2, input_dim * num_speaker))
        self.input_dim, self.hidden_dim = input_dim, hidden_dim
        self.embedding_dim, self.num_speaker = embedding_dim, num_speaker
        self._rnn_cache, self._dc_cache, self._mi_cache = PackCache(), PackCache(), PackCache()
        self.use_tensor_cores = True

    def forward(self, input):
        assert len(input) == 1, "There must be one tensor in the input for the chimera network"
        x = input[0].float()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from ._train import chimera_backward, chimera_forward_train, run_model
            embedding, masks = run_model(self, chimera_forward_train, chimera_backward, x)
            return [embedding, masks[:, :, :, 0], masks[:, :, :, 1]]
        B, T, F = x.shape
        H, D, S = self.hidden_dim, self.embedding_dim, self.num_speaker
        M = T * B
        y_h, _ = blstm_forward(self.rnn, self._rnn_cache, x, self.training, want_f32=False, want_f16=True,
                               use_tensor_cores=self.use_tensor_cores, lengths=eval_lengths(self, x.shape[0], x.device))
        wdc = self._dc_cache.get([self.fc_dc.weight], lambda: _lib.pack_linear_f16(self.fc_dc.weight, True, H))
        wmi = self._mi_cache.get([self.fc_mi.weight], lambda: _lib.pack_linear_f16(self.fc_mi.weight, True, H))
        emb = torch.empty(B, T, F, D, device=x.device, dtype=torch.float32)
        if not _lib.gemm_l2norm_supported(D):
            raise _lib.OnssenB200Error(f"embedding_dim={D}: fused normalise epilogue supports 4/8/12/16/20/24/32/40")
        _lib.gemm_f16(y_h, wdc, self.fc_dc.bias.detach(), emb, M, F * D, y_h.shape[1], F * D, epi=3, group=D,
                      remap_inner=B, remap_outer=T)
        masks = torch.empty(B, T, F, S, device=x.device, dtype=torch.float32)
        _lib.gemm_f16(y_h, wmi, self.fc_mi.bias.detach(), masks, M, F * S, y_h.shape[1], F * S, epi=1,
                      remap_inner=B, remap_outer=T)
        mask_A = masks[:, :, :, 0]
        mask_B = masks[:, :, :, 1]
        return [emb, mask_A, mask_B]


