"""enhance (restoration layers), phase_net (repaired) and their losses through the plugin API."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import onssen_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["small", "mid"])
def test_enhance_vs_reference_fixture(cuda_device, name):
    import onssen_b200 as ob
    p, g = load_golden(f"enhance_{name}.npz")
    B, T, F, H, L, D = [int(v) for v in g["cfg"]]
    model = ob.nn.enhance(F, H, L, dropout=0.0).to(cuda_device)
    assert sorted(k for k in model.state_dict() if "num_batches" not in k) == sorted(p)
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()}, strict=False)
    cu = lambda a: torch.from_numpy(a).to(cuda_device)
    scale = np.abs(g["clean_eval"]).max()
    with torch.no_grad():
        model.eval()
        clean, = model([cu(g["feature"]), cu(g["mag_noisy"])])
        assert clean.shape == (B, T, F)
        np.testing.assert_allclose(clean.cpu().numpy(), g["clean_eval"], atol=3e-3 * scale)
        model.train()
        clean_t, = model([cu(g["feature"]), cu(g["mag_noisy"])])
        np.testing.assert_allclose(clean_t.cpu().numpy(), g["clean_train"], atol=3e-3 * scale)
        # loss kernels on the reference's own outputs (fp32 arithmetic)
        ce = cu(g["clean_eval"])
        l_msa = ob.loss.loss_mask_msa([ce], [cu(g["mag_clean"]), cu(g["cos_diff"])])
        np.testing.assert_allclose(l_msa.item(), g["loss_msa"], rtol=1e-5)
        l_psa = ob.loss.loss_mask_psa([torch.sigmoid(ce)], [cu(g["mag_noisy"]), cu(g["mag_clean"]), cu(g["cos_diff"])])
        np.testing.assert_allclose(l_psa.cpu().numpy(), g["loss_psa"], rtol=1e-5)


def test_phase_net_vs_repaired_oracle(cuda_device):
    import onssen_b200 as ob
    B, T, F, H, L, D = 3, 24, 33, 40, 2, 20
    torch.manual_seed(5)
    model = ob.nn.phase_net(F, H, L, D, dropout=0.0).to(cuda_device).eval()
    with torch.no_grad():
        model.bn.running_mean.normal_(0, 0.1); model.bn.running_var.uniform_(0.5, 1.5)
    rng = np.random.RandomState(1)
    x_mag = np.abs(rng.standard_normal((B, T, F))).astype(np.float32)
    x_ph = rng.standard_normal((B, T, F, 2)).astype(np.float32)
    cu = lambda a: torch.from_numpy(a).to(cuda_device)
    with torch.no_grad():
        out = model([cu(x_mag), cu(x_ph)])
    params = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    ref = O.phase_net_forward(params, [x_mag, x_ph], L, training=False)
    assert len(out) == 5
    for o, r, tol in zip(out, ref, (2e-3, 1e-3, 1e-3, 3e-3, 3e-3)):
        np.testing.assert_allclose(o.cpu().numpy(), r, atol=tol)
    # repaired loss_phase vs the oracle's repaired restatement, on identical (reference-layout) inputs
    mags = [np.abs(rng.standard_normal((B, T, F))).astype(np.float32) for _ in range(3)]
    ph = [O.l2_normalize(rng.standard_normal((B, T, F, 2)).astype(np.float32)) for _ in range(2)]
    oh = np.zeros((B, T, F, 2)); oh[..., 0] = rng.uniform(size=(B, T, F)) > 0.5; oh[..., 1] = 1 - oh[..., 0]
    lab_np = [oh, mags[0], mags[1], mags[2], ph[0], ph[1]]
    with torch.no_grad():
        got = ob.loss.loss_phase(out, [cu(a) for a in lab_np])
    want = O.loss_phase([o.cpu().numpy() for o in out], lab_np)
    assert got.shape == (B, B)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=5e-5)


def test_l2norm_pairs_backward_matches_autograd(cuda_device):
    """(ph + x_phase) / |ph + x_phase| over (re, im) pairs (phase_network.py:55-56,63-66): the hand-written backward
    against torch autograd of F.normalize on the same fp32 inputs, at the cfg4 bin count; elementwise, so tight."""
    import torch.nn.functional as Fn
    from onssen_b200 import _lib
    B, T, F = 4, 50, 257
    g = torch.Generator(device="cpu").manual_seed(3)
    ph = torch.randn(B, T, F, 2, generator=g).to(cuda_device)
    xp = torch.randn(B, T, F, 2, generator=g).to(cuda_device)
    dy = torch.randn(B, T, F, 2, generator=g).to(cuda_device)
    v = (ph.clone().requires_grad_(True))
    out = Fn.normalize(v + xp, p=2, dim=-1)
    out.backward(dy)
    dz, sc = _lib.l2norm_pairs_bwd(dy, ph, xp)                     # time-major [T*B][2F]
    want = v.grad.permute(1, 0, 2, 3).reshape(T * B, 2 * F)
    scale = want.abs().max().item()
    assert (dz - want).abs().max().item() < 1e-5 * scale
    assert torch.allclose(_lib.add_l2norm_pairs(ph, xp), out.detach(), atol=1e-6)
    assert abs(sc[0].item() * sc[1].item() - 1.0) < 1e-6 and 256 <= scale * sc[0].item() <= 2048   # power-of-two scale
