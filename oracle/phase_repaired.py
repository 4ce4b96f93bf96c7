"""Torch restatement WITH REPAIRS of the two reference pieces that raise as written (test infrastructure only):
`phase_net` (onssen/nn/phase_network.py:28 uses an undefined `output_dim`) and `loss_phase`
(onssen/loss/loss_phase.py:7-13 unpacks 5 of 6 and calls loss_dc with the wrong lists).  Built on the LIVE reference
`chimera` and `loss_dc`, so everything except the repaired lines is the reference's own code.  Repairs (SURVEY.md
8a-14 / a18): `fc_phase = Linear(2H, 2F)`; 5 outputs; `loss_dc([embedding], [one_hot_label, mag_mix])`.
These rows stay "parity unpinned by nature": there is no running reference to compare with."""


def build(onssen):
    import torch
    import torch.nn.functional as Fn

    class PhaseNetRepaired(torch.nn.Module):
        def __init__(self, F, H, L, D):
            super().__init__()
            self.rnn = torch.nn.LSTM(3 * F, H, L, dropout=0.0, bidirectional=True, batch_first=True)
            self.bn = torch.nn.BatchNorm1d(2 * H)
            self.fc_phase = torch.nn.Linear(2 * H, 2 * F)         # repair: output_dim := num_speaker * input_dim
            self.chimera = onssen.nn.chimera(F, H, L, D, dropout=0.0)

        def forward(self, inp):
            x_mag, x_phase = inp
            emb, m_a, m_b = self.chimera([x_mag])
            B, T, F = m_a.shape
            outs, self.pre_norms = [], []
            for m in (m_a, m_b):
                y, _ = self.rnn(torch.cat((x_mag * m, x_phase.reshape(B, T, -1)), 2))
                y = self.bn(y.permute(0, 2, 1)).permute(0, 2, 1)
                v = self.fc_phase(y).reshape(B, T, F, -1) + x_phase
                self.pre_norms.append(v.detach().norm(dim=-1))   # conditioning of the normalisation, for the tests
                outs.append(Fn.normalize(v, p=2, dim=-1))
            return [emb, m_a, m_b] + outs

    def loss_phase_repaired(output, label):
        emb, m_a, m_b, p_a, p_b = output                           # repair: 5 outputs
        oh, mix, s1, s2, ph1, ph2 = label
        B = mix.shape[0]
        l_emb = onssen.loss.loss_dc([emb], [oh, mix])              # repair: mag_mix belongs to the label list
        l1n = lambda x: x.abs().reshape(B, -1).sum(1)
        lm1 = l1n(m_a * mix - s1) + l1n(m_b * mix - s2)
        lm2 = l1n(m_b * mix - s1) + l1n(m_a * mix - s2)
        cs = lambda a, b: Fn.cosine_similarity(a, b, dim=3)
        lp1 = (-mix * cs(p_a, ph1) - mix * cs(p_b, ph2)).reshape(B, -1).sum(1)
        lp2 = (-mix * cs(p_b, ph1) - mix * cs(p_a, ph2)).reshape(B, -1).sum(1)
        first = lm1 < lm2
        return l_emb * 0.975 + torch.where(first, lm1, lm2) * 0.025 + torch.where(first, lp1, lp2) * 0.025

    return PhaseNetRepaired, loss_phase_repaired
