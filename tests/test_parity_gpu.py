"""Model-level parity of the CUDA path (through the plugin API) against the reference's own outputs
(golden fixtures) and against the oracle on seeded inputs.

Tolerances (north_star: <=1e-4 relative deviation of the dpcl loss vs the fp32 reference; embeddings to a
stated tolerance): the GEMMs use fp16 operands with fp32 accumulation, which the design measured at ~3e-4
absolute on unit-norm embeddings and ~2e-6 relative on the loss at BASELINE size.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import onssen_oracle as O

pytestmark = pytest.mark.gpu

EMB_ATOL = 2e-3      # unit-norm embedding entries, fp16-operand GEMMs
LOSS_RTOL = 1e-4     # north_star tolerance on the dpcl loss


def load_state(model, params):
    sd = {k: torch.from_numpy(v.copy()) for k, v in params.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("num_batches" in m for m in missing), (missing, unexpected)


@pytest.mark.parametrize("name", ["small", "mid"])
@pytest.mark.parametrize("tc", [True, False])
def test_deep_clustering_vs_reference_fixture(cuda_device, name, tc):
    import onssen_b200 as ob
    p, g = load_golden(f"dc_{name}.npz")
    B, T, F, H, L, D = [int(v) for v in g["cfg"]]
    model = ob.nn.deep_clustering(F, H, L, D, dropout=0.0).to(cuda_device)
    model.use_tensor_cores = tc
    assert sorted(k for k in model.state_dict() if "num_batches" not in k) == sorted(p)   # drop-in key names
    load_state(model, p)
    cu = lambda a: torch.from_numpy(a).to(cuda_device)
    with torch.no_grad():
        model.eval()
        emb, = model([cu(g["feature"])])
        loss = ob.loss.loss_dc([emb], [cu(g["one_hot"]), cu(g["mag_mix"])])
        assert emb.shape == (B, T, F, D) and loss.shape == (B, B)
        np.testing.assert_allclose(emb.cpu().numpy(), g["emb_eval"], atol=EMB_ATOL)
        np.testing.assert_allclose(loss.cpu().numpy(), g["loss_eval"], rtol=20 * LOSS_RTOL)  # tiny T*F: little averaging
        model.train()
        emb_t, = model([cu(g["feature"])])
        np.testing.assert_allclose(emb_t.cpu().numpy(), g["emb_train"], atol=2 * EMB_ATOL)  # B*T=48 samples: batch stats amplify rounding
        np.testing.assert_allclose(model.bn.running_mean.cpu().numpy(), g["bn_rm_after"], atol=1e-4)
        np.testing.assert_allclose(model.bn.running_var.cpu().numpy(), g["bn_rv_after"], rtol=1e-3)
    # the loss kernel alone, on the reference's own embedding: fp32 arithmetic, tight tolerance
    loss_ref_in = ob.loss.loss_dc([cu(g["emb_eval"])], [cu(g["one_hot"]), cu(g["mag_mix"])])
    np.testing.assert_allclose(loss_ref_in.cpu().numpy(), g["loss_eval"], rtol=2e-5)


@pytest.mark.parametrize("name", ["small", "mid"])
def test_chimera_vs_reference_fixture(cuda_device, name):
    import onssen_b200 as ob
    p, g = load_golden(f"chimera_{name}.npz")
    B, T, F, H, L, D = [int(v) for v in g["cfg"]]
    model = ob.nn.chimera(F, H, L, D, dropout=0.0).to(cuda_device).eval()
    load_state(model, p)
    cu = lambda a: torch.from_numpy(a).to(cuda_device)
    with torch.no_grad():
        e, ma, mb = model([cu(g["feature"])])
        assert ma.shape == (B, T, F) and ma.stride(-1) == 2           # strided views like the reference
        np.testing.assert_allclose(e.cpu().numpy(), g["emb"], atol=EMB_ATOL)
        np.testing.assert_allclose(ma.cpu().numpy(), g["mask_a"], atol=1e-3)
        np.testing.assert_allclose(mb.cpu().numpy(), g["mask_b"], atol=1e-3)
        lab = [cu(g[k]) for k in ("one_hot", "mag_mix", "mag_s1", "mag_s2", "cos_s1", "cos_s2")]
        l_msa = ob.loss.loss_chimera_msa([e, ma, mb], lab[:4])
        l_psa = ob.loss.loss_chimera_psa([e, ma, mb], lab)
        np.testing.assert_allclose(l_msa.cpu().numpy(), g["loss_msa"], rtol=20 * LOSS_RTOL)
        np.testing.assert_allclose(l_psa.cpu().numpy(), g["loss_psa"], rtol=20 * LOSS_RTOL)
        # loss kernels alone on the reference's outputs
        eo, mao, mbo = cu(g["emb"]), cu(g["mask_a"]), cu(g["mask_b"])
        np.testing.assert_allclose(ob.loss.loss_chimera_psa([eo, mao, mbo], lab).cpu().numpy(), g["loss_psa"], rtol=2e-5)


def test_deep_clustering_baseline_shape_vs_oracle(cuda_device):
    """BASELINE cfg2 shape (3x600 BLSTM, 129 bins, D=40, T=400) at a batch the numpy oracle finishes in
    seconds; inputs from the device featurizer on seeded synthetic mixtures."""
    import onssen_b200 as ob
    B, T, F, H, L, D = 4, 400, 129, 600, 3, 40
    torch.manual_seed(1)
    model = ob.nn.deep_clustering(F, H, L, D).to(cuda_device).eval()
    with torch.no_grad():
        model.bn.running_mean.normal_(0, 0.1); model.bn.running_var.uniform_(0.5, 1.5)
        model.bn.weight.uniform_(0.5, 1.5); model.bn.bias.normal_(0, 0.1)
    utts = [O.synth_utterance(i) for i in range(B)]
    cu = lambda k: torch.from_numpy(np.stack([u[k] for u in utts])).to(cuda_device)
    starts = torch.tensor([0, 100, 37, 64], dtype=torch.int32)
    with torch.no_grad():
        inp, lab = ob.data.featurize_batch(cu(0), cu(1), cu(2), "dc", 256, 64, T, 40, crop_start=starts)
        emb, = model(inp)
        loss = ob.loss.loss_dc([emb], lab)
    params = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    feat = inp[0].cpu().numpy()
    emb_ref, = O.deep_clustering_forward(params, [feat], L, training=False)
    loss_ref = O.loss_dc([emb_ref], [lab[0].cpu().numpy(), lab[1].cpu().numpy()])
    err = np.abs(emb.cpu().numpy() - emb_ref).max()
    rel = np.abs(loss.cpu().numpy() - loss_ref).max() / np.abs(loss_ref).max()
    print(f"cfg2-shape parity: max|emb err|={err:.3e}  loss rel dev={rel:.3e}")
    assert err < EMB_ATOL
    assert rel < LOSS_RTOL
    # size-independent properties: unit norm, loss invariant under a joint permutation of the utterances
    n = emb.norm(dim=-1)
    assert (n - 1).abs().max().item() < 1e-5
    perm = torch.tensor([2, 0, 3, 1], device=cuda_device)
    with torch.no_grad():
        emb_p, = model([inp[0][perm].contiguous()])
    assert (emb_p - emb[perm]).abs().max().item() < 5e-4   # batch-slice assignment changes only rounding order


def test_chimera_pp_baseline_shape_vs_oracle(cuda_device):
    """BASELINE cfg3 shape (chimera++ 4x600 BLSTM, mask + embed heads, PSA / W_MR loss, T=400) at a batch the
    numpy oracle finishes in seconds; inputs from the device featurizer on seeded synthetic mixtures."""
    import onssen_b200 as ob
    B, T, F, H, L, D = 3, 400, 129, 600, 4, 20
    torch.manual_seed(2)
    model = ob.nn.chimera(F, H, L, D).to(cuda_device).eval()
    utts = [O.synth_utterance(20 + i) for i in range(B)]
    cu = lambda k: torch.from_numpy(np.stack([u[k] for u in utts])).to(cuda_device)
    starts = torch.tensor([3, 50, 99], dtype=torch.int32)
    with torch.no_grad():
        inp, lab = ob.data.featurize_batch(cu(0), cu(1), cu(2), "chimera++", 256, 64, T, 40, crop_start=starts)
        out = model(inp)
        loss = ob.loss.loss_chimera_psa(out, lab)
    params = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    ref = O.chimera_forward(params, [inp[0].cpu().numpy()], L)
    loss_ref = O.loss_chimera_psa(ref, [t.cpu().numpy() for t in lab])
    err_e = np.abs(out[0].cpu().numpy() - ref[0]).max()
    err_m = max(np.abs(out[1].cpu().numpy() - ref[1]).max(), np.abs(out[2].cpu().numpy() - ref[2]).max())
    rel = np.abs(loss.cpu().numpy() - loss_ref).max() / np.abs(loss_ref).max()
    print(f"cfg3-shape parity: max|emb err|={err_e:.3e} max|mask err|={err_m:.3e} loss rel dev={rel:.3e}")
    assert err_e < EMB_ATOL and err_m < 1e-3 and rel < LOSS_RTOL
