"""Parity at the EXACT BASELINE configurations against the reference itself (oracle/_ref: the unmodified
speechLabBcCuny/onssen modules staged by oracle/make_ref.py), run in fp32 on the GPU box's host cores:

  cfg2  deep_clustering(129, 600, 3, 40), B=32, T=400, train-mode BatchNorm, dropout 0, loss_dc
  cfg3  chimera(129, 600, 4, 20) ("chimera++": 6 labels, loss_chimera_psa), B=64, T=400
  cfg5  enhance(513, 600, 3) per-GPU shard B=32, T=400, loss_mask_msa          (Edinburgh-TTS shape)
  cfg4  phase_net(257, 300, 3, 20) repaired, per-GPU shard B=16, T=400         (repaired restatement, unpinned)

forward outputs, loss, and EVERY parameter gradient of torch.mean(loss) (the trainer's scalar, train.py:79-82)
are compared: the reference's autograd at H=600 is the oracle of the persistent BPTT kernel at the configuration the
bench times.  Tolerances are written next to each assert (north_star: <=1e-4 relative on the dpcl loss)."""
import os

import numpy as np
import pytest
import torch

from oracle import onssen_oracle as O
from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not staged")]

GRAD_RTOL = 5e-3          # relative L2 error per parameter tensor (measured values are printed)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def synth_inputs(model_name, B, n_fft, hop, nsample, T, dev, first=0):
    import onssen_b200 as ob
    utts = [O.synth_utterance(first + i, nsample) for i in range(B)]
    cu = lambda k: torch.from_numpy(np.stack([u[k] for u in utts])).to(dev)
    hi = O.num_crop_starts(nsample, hop, T)
    starts = torch.from_numpy(np.array([np.random.RandomState(77 + i).randint(hi) for i in range(B)], dtype=np.int32))
    with torch.no_grad():
        return ob.data.featurize_batch(cu(0), cu(1), cu(2), model_name, n_fft, hop, T, 40, crop_start=starts)


def compare(ours, ref, loss_ours, loss_ref, inp, lab, out_tols, loss_rtol, grad_rtol=GRAD_RTOL, tag="", weights=None):
    """runs both in train mode (dropout 0), compares outputs / loss / all parameter gradients."""
    torch.set_num_threads(os.cpu_count() or 8)
    ref.load_state_dict({k: v.detach().cpu() for k, v in ours.state_dict().items()})
    ours.train(); ref.train()
    from onssen_b200 import _lib
    _lib.bptt_saturation_count(reset=True)
    out = ours(inp)
    lo = loss_ours(out, lab)
    torch.mean(lo).backward()
    nsat = _lib.bptt_saturation_count(reset=True)
    print(f"{tag} BPTT exchange values clamped: {nsat}")
    out_r = ref([t.cpu() for t in inp])
    lr = loss_ref(out_r, [t.cpu() for t in lab])
    torch.mean(lr).backward()
    assert lo.shape == lr.shape
    weights = weights(ref) if weights is not None else [None] * len(out)
    for i, (a, b, tol, w) in enumerate(zip(out, out_r, out_tols, weights)):
        d = (a.detach().cpu() - b.detach()).abs()
        if w is not None:
            d = d * w                                   # ill-conditioned entries weighted down (stated at the caller)
        err, lim = float(d.max()), tol * max(1.0, float(b.detach().abs().max()))
        print(f"{tag} output[{i}] max abs err {err:.3e} (limit {lim:g})")
        assert a.shape == b.shape and err < lim, (i, err)
    lrel = float((lo.detach().cpu() - lr.detach()).abs().max() / lr.detach().abs().max())
    print(f"{tag} loss rel dev {lrel:.3e} (limit {loss_rtol:g})")
    assert lrel < loss_rtol
    ref_grads = dict(ref.named_parameters())
    errs = []
    for k, p in ours.named_parameters():
        assert p.grad is not None, k
        errs.append((rel_err(p.grad.cpu(), ref_grads[k].grad), k))
    errs.sort(reverse=True)
    print(f"{tag} worst relative gradient errors: " + ", ".join(f"{k} {e:.3e}" for e, k in errs[:4]))
    assert errs[0][0] < grad_rtol, errs[:4]
    # running statistics of train-mode BatchNorm follow the reference (momentum 0.1, unbiased variance)
    for k, v in ours.state_dict().items():
        if "running_" in k:
            torch.testing.assert_close(v.cpu(), ref.state_dict()[k], rtol=2e-3, atol=2e-4)
    return lrel, errs[0][0]


def test_cfg2_deep_clustering_b32_vs_reference(cuda_device):
    import onssen_b200 as ob
    R = ref_loader.import_reference()
    torch.manual_seed(11)
    ours = ob.nn.deep_clustering(129, 600, 3, 40, dropout=0.0).to(cuda_device)
    ref = R.nn.deep_clustering(129, 600, 3, 40, dropout=0.0)
    inp, lab = synth_inputs("dc", 32, 256, 64, 32000, 400, cuda_device)
    compare(ours, ref, ob.loss.loss_dc, R.loss.loss_dc, inp, lab, out_tols=[2e-3], loss_rtol=1e-4, tag="cfg2")


def test_cfg3_chimera_pp_b64_vs_reference(cuda_device):
    import onssen_b200 as ob
    R = ref_loader.import_reference()
    torch.manual_seed(12)
    ours = ob.nn.chimera(129, 600, 4, 20, dropout=0.0).to(cuda_device)
    ref = R.nn.chimera(129, 600, 4, 20, dropout=0.0)
    inp, lab = synth_inputs("chimera++", 64, 256, 64, 32000, 400, cuda_device, first=100)
    compare(ours, ref, ob.loss.loss_chimera_psa, R.loss.loss_chimera_psa, inp, lab, out_tols=[2e-3, 1e-3, 1e-3],
            loss_rtol=1e-4, tag="cfg3")


def test_cfg5_enhance_f513_vs_reference(cuda_device):
    """Edinburgh-TTS shape: 16 kHz, n_fft 1024 / hop 256 (513 bins), 251 frames <= 400 -> the tiling branch."""
    import onssen_b200 as ob
    R = ref_loader.import_reference()
    torch.manual_seed(13)
    ours = ob.nn.enhance(513, 600, 3, dropout=0.0).to(cuda_device)
    ref = R.nn.enhance(513, 600, 3, dropout=0.0)
    inp, lab = synth_inputs("chimera++", 32, 1024, 256, 64000, 400, cuda_device, first=200)
    # Edinburgh layout (edinburgh_tts.py:84-97): input [feature, mag_noisy], label [mag_clean, cos_diff]
    e_inp, e_lab = [inp[0], lab[1]], [lab[2], lab[4]]
    compare(ours, ref, ob.loss.loss_mask_msa, R.loss.loss_mask_msa, e_inp, e_lab, out_tols=[2e-3], loss_rtol=1e-4,
            tag="cfg5")


def test_cfg4_phase_net_f257_vs_repaired_restatement(cuda_device):
    """phase_net / loss_phase raise in the reference; oracle/phase_repaired.py wraps the live chimera + loss_dc with the
    listed repairs.  16 kHz, n_fft 512 / hop 128 (257 bins); egs/wsj0-2mix/phase-net/config.json shape, B=16/GPU."""
    import onssen_b200 as ob
    from oracle.phase_repaired import build
    R = ref_loader.import_reference()
    PhaseNetRepaired, loss_phase_repaired = build(R)
    torch.manual_seed(14)
    F, H, L, D, B = 257, 300, 3, 20, 16
    ours = ob.nn.phase_net(F, H, L, D, dropout=0.0).to(cuda_device)
    ref = PhaseNetRepaired(F, H, L, D)
    # A freshly initialised mask head gives mask_A ~ mask_B ~ 0.5, which makes the two PIT assignments of the mask loss
    # tie to ~1e-5 relative: the arg-min (and with it the whole phase-loss gradient) is then decided by rounding.  The
    # loss is continuous there but its gradient is not, so the comparison is made at a point where the permutation is
    # well separated: speaker-0 mask biased up, speaker-1 mask biased down (output index f*2 + s, chimera.py:42-45).
    with torch.no_grad():
        b = ours.chimera.fc_mi.bias.view(F, 2)
        b[:, 0] += 1.0
        b[:, 1] -= 1.0
        # The (re, im) normalisation v/|v| with v = fc_phase(y) + x_phase has a 1/|v| gradient.  With a default-size
        # fc_phase output (|ph| ~ 0.3) a few of the 1.6 M bins land within 1e-3 of ph = -x_phase; their 1/|v| ~ 1e3
        # contributions dominate the whole gradient and flip with a 1e-4 change of ph (measured: the SAME kernels give
        # 4e-2 at T=40 and 1.0 at T=400 relative error against fp32 torch, while d loss / d phase matches to 5e-4) --
        # the repaired model is ill-conditioned there, in any arithmetic.  A small phase head keeps |v| ~ |x_phase|,
        # where mag_mix / |v| is bounded, so the comparison checks the kernels rather than the conditioning.
        ours.fc_phase.weight.mul_(0.01)
        ours.fc_phase.bias.mul_(0.01)
    inp, lab = synth_inputs("phase", B, 512, 128, 64000, 400, cuda_device, first=300)
    # gradients: the (re,im) normalisation divides by |v|, tiny on quiet bins -> looser limit
    # phase outputs are v/|v|: an absolute error e on v moves the unit vector by ~e/|v| -> weight by min(1, |v|)
    wts = lambda r: [None, None, None] + [n.clamp(max=1.0)[..., None] for n in r.pre_norms]
    compare(ours, ref, ob.loss.loss_phase, loss_phase_repaired, inp, lab, out_tols=[2e-3, 1e-3, 1e-3, 3e-3, 3e-3],
            loss_rtol=2e-4, grad_rtol=3e-2, tag="cfg4", weights=wts)
    # the permutation margin of the reference at this point (relative gap of the two PIT sums per utterance)
    with torch.no_grad():
        _, m_a, m_b, _, _ = ref([t.cpu() for t in inp])
        mix, s1, s2 = (t.cpu() for t in lab[1:4])
        l1n = lambda x: x.abs().reshape(B, -1).sum(1)
        lm1 = l1n(m_a * mix - s1) + l1n(m_b * mix - s2)
        lm2 = l1n(m_b * mix - s1) + l1n(m_a * mix - s2)
        margin = ((lm1 - lm2).abs() / torch.minimum(lm1, lm2)).min().item()
    print(f"cfg4 smallest relative PIT margin {margin:.3e}")
    assert margin > 1e-3
