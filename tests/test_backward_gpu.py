"""Training path: gradients of torch.mean(loss_dc(model(x))) from the hand-written backward vs the gradients
the REFERENCE's autograd produced (golden fixtures tests/golden/dcgrad_*.npz)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_loss_dc_backward_vs_reference(cuda_device):
    import onssen_b200 as ob
    for name in ("small", "mid"):
        _, g = load_golden(f"dc_{name}.npz")
        gz = np.load(f"tests/golden/dcgrad_{name}.npz")
        emb = torch.from_numpy(g["emb_train"]).to(cuda_device).requires_grad_(True)
        loss = ob.loss.loss_dc([emb], [torch.from_numpy(g["one_hot"]).to(cuda_device),
                                       torch.from_numpy(g["mag_mix"]).to(cuda_device)])
        torch.mean(loss).backward()
        # the fixture's gradient was taken at the reference's own train-mode embedding
        assert rel_err(emb.grad.cpu().numpy(), gz["g:embedding"]) < 1e-4


@pytest.mark.parametrize("name", ["small", "mid"])
def test_deep_clustering_gradients_vs_reference(cuda_device, name):
    import onssen_b200 as ob
    p, g = load_golden(f"dc_{name}.npz")
    gz = np.load(f"tests/golden/dcgrad_{name}.npz")
    B, T, F, H, L, D = [int(v) for v in g["cfg"]]
    model = ob.nn.deep_clustering(F, H, L, D, dropout=0.0).to(cuda_device).train()
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()}, strict=False)
    cu = lambda a: torch.from_numpy(a).to(cuda_device)
    emb, = model([cu(g["feature"])])
    assert emb.requires_grad
    torch.mean(ob.loss.loss_dc([emb], [cu(g["one_hot"]), cu(g["mag_mix"])])).backward()
    worst = 0.0
    for k, v in model.named_parameters():
        assert v.grad is not None, k
        e = rel_err(v.grad.cpu().numpy(), gz["g:" + k])
        worst = max(worst, e)
        print(f"{k:28s} rel err {e:.3e}  |g|={np.linalg.norm(gz['g:' + k]):.3e}")
        assert e < 3e-2, (k, e)
    print("worst relative gradient error", worst)


def test_training_step_reduces_loss(cuda_device):
    """trainer-style optimisation steps on one synthetic batch: the loss must go down."""
    import onssen_b200 as ob
    from oracle import onssen_oracle as O
    torch.manual_seed(0)
    B, T = 4, 60
    model = ob.nn.deep_clustering(129, 64, 2, 20, dropout=0.3).to(cuda_device).train()
    utts = [O.synth_utterance(i, 8000) for i in range(B)]
    cu = lambda k: torch.from_numpy(np.stack([u[k] for u in utts])).to(cuda_device)
    inp, lab = ob.data.featurize_batch(cu(0), cu(1), cu(2), "dc", 256, 64, T, 40, crop_start=torch.zeros(B, dtype=torch.int32))
    opt = ob.utils.build_optimizer(model.parameters(), {"name": "adam", "lr": 1e-3})
    losses = []
    for _ in range(12):
        loss = torch.mean(ob.loss.loss_dc(model(inp), lab))
        opt.zero_grad()
        loss.backward()
        ob.utils.clip_grad_norm_(model.parameters(), 5)
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and losses[-1] < 0.9 * losses[0], losses


@pytest.mark.parametrize("name", ["small", "mid"])
def test_chimera_gradients_vs_reference(cuda_device, name):
    import onssen_b200 as ob
    p, g = load_golden(f"chimera_{name}.npz")
    gz = np.load(f"tests/golden/chimeragrad_{name}.npz")
    B, T, F, H, L, D = [int(v) for v in g["cfg"]]
    model = ob.nn.chimera(F, H, L, D, dropout=0.0).to(cuda_device).train()
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()}, strict=False)
    cu = lambda a: torch.from_numpy(a).to(cuda_device)
    out = model([cu(g["feature"])])
    lab = [cu(g[k]) for k in ("one_hot", "mag_mix", "mag_s1", "mag_s2", "cos_s1", "cos_s2")]
    torch.mean(ob.loss.loss_chimera_psa(out, lab)).backward()
    worst = 0.0
    for k, v in model.named_parameters():
        e = rel_err(v.grad.cpu().numpy(), gz["g:" + k])
        worst = max(worst, e)
        assert e < 3e-2, (k, e)
    print("chimera++ worst relative gradient error", worst)


@pytest.mark.parametrize("name", ["small", "mid"])
def test_enhance_gradients_vs_reference(cuda_device, name):
    import onssen_b200 as ob
    p, g = load_golden(f"enhance_{name}.npz")
    gz = np.load(f"tests/golden/enhancegrad_{name}.npz")
    B, T, F, H, L, D = [int(v) for v in g["cfg"]]
    model = ob.nn.enhance(F, H, L, dropout=0.0).to(cuda_device).train()
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()}, strict=False)
    cu = lambda a: torch.from_numpy(a).to(cuda_device)
    clean, = model([cu(g["feature"]), cu(g["mag_noisy"])])
    ob.loss.loss_mask_msa([clean], [cu(g["mag_clean"]), cu(g["cos_diff"])]).backward()
    worst = 0.0
    for k, v in model.named_parameters():
        e = rel_err(v.grad.cpu().numpy(), gz["g:" + k])
        worst = max(worst, e)
        assert e < 3e-2, (k, e)
    print("enhance worst relative gradient error", worst)


@pytest.mark.parametrize("name", ["small", "mid"])
def test_phase_net_gradients_vs_torch_restatement(cuda_device, name):
    """phase_net + loss_phase (both REPAIRED): forward and every parameter gradient against
    the fixtures of the torch restatement with the same repairs (oracle/make_golden.py main_phase)."""
    import onssen_b200 as ob
    p, g = load_golden(f"phase_{name}.npz")
    gz = np.load(f"tests/golden/phasegrad_{name}.npz")
    B, T, F, H, L, D = [int(v) for v in g["cfg"]]
    model = ob.nn.phase_net(F, H, L, D, dropout=0.0).to(cuda_device).train()
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()}, strict=False)
    cu = lambda a: torch.from_numpy(a).to(cuda_device)
    out = model([cu(g["feature"]), cu(g["x_phase"])])
    # the (re,im) normalisation amplifies errors by 1/|v| where the pre-normalisation vector is short (x_phase is a raw
    # STFT, so it is on quiet bins): weight the phase error by min(1, |v|) of the restatement
    one = np.ones(g["mask_a"].shape, np.float32)
    for o, key, w in zip(out, ["emb", "mask_a", "mask_b", "phase_a", "phase_b"],
                         [one[..., None], one, one, np.minimum(1, g["norm_a"])[..., None], np.minimum(1, g["norm_b"])[..., None]]):
        assert (np.abs(o.detach().cpu().numpy() - g[key]) * w).max() < 3e-3, key
    loss = ob.loss.loss_phase(out, [cu(g[k]) for k in ("one_hot", "mag_mix", "mag_s1", "mag_s2", "phase_s1", "phase_s2")])
    assert np.allclose(loss.detach().cpu().numpy(), g["loss"], rtol=2e-3, atol=1e-2)
    torch.mean(loss).backward()
    worst = 0.0
    for k, v in model.named_parameters():
        e = rel_err(v.grad.cpu().numpy(), gz["g:" + k])
        worst = max(worst, e)
        assert e < 3e-2, (k, e)
    print("phase_net worst relative gradient error", worst)


def test_loss_mask_psa_gradient(cuda_device):
    import onssen_b200 as ob
    from oracle import onssen_oracle as O
    _, g = load_golden("enhance_mid.npz")
    rng = np.random.RandomState(3)
    mask = rng.uniform(0, 1, g["mag_noisy"].shape).astype(np.float32)
    cu = lambda a: torch.from_numpy(a).to(cuda_device)
    m = cu(mask).requires_grad_(True)
    loss = ob.loss.loss_mask_psa([m], [cu(g["mag_noisy"]), cu(g["mag_clean"]), cu(g["cos_diff"])])
    w = rng.uniform(0.5, 2, loss.shape[0]).astype(np.float32)
    (loss * cu(w)).sum().backward()
    want = O.loss_mask_psa_grad(mask, g["mag_noisy"], g["mag_clean"], g["cos_diff"], w)
    assert np.abs(m.grad.cpu().numpy() - want).max() < 1e-6


def test_clip_grad_norm_and_adam_kernels_vs_oracle(cuda_device):
    """multi-tensor clip_grad_norm_ + Adam (csrc/optim.cu) against the oracle restatement, clipping and not clipping"""
    import onssen_b200 as ob
    from oracle import onssen_oracle as O
    rng = np.random.RandomState(9)
    shapes = [(2400, 129), (70001,), (5, 3, 7), (1,), (65536,), (65537,)]
    ps = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    tp = [torch.nn.Parameter(torch.from_numpy(p.copy()).to(cuda_device)) for p in ps]
    opt = ob.utils.build_optimizer(tp, {"name": "adam", "lr": 1e-3})
    assert isinstance(opt, ob.utils.Adam)
    m = [np.zeros_like(p) for p in ps]
    v = [np.zeros_like(p) for p in ps]
    for step in range(1, 5):
        gs = [(rng.standard_normal(s) * (1.0 if step % 2 else 1e-3)).astype(np.float32) for s in shapes]
        for t, g in zip(tp, gs):
            t.grad = torch.from_numpy(g.copy()).to(cuda_device)
        ver = tp[0]._version
        tn = float(ob.utils.clip_grad_norm_(tp, 5))
        opt.step()
        assert tp[0]._version > ver
        n = O.clip_adam_step(ps, [g.copy() for g in gs], m, v, step)
        assert abs(n - tn) < 1e-5 * max(n, 1e-3), (n, tn)
        for a, t in zip(ps, tp):
            assert np.abs(a - t.detach().cpu().numpy()).max() < 2e-6


def test_sync_batchnorm_kernels_equal_full_batch(cuda_device):
    """cross-rank BatchNorm (SURVEY.md 8e parity mode): two half-batch 'ranks' whose column sums are added (the
    all-reduce, emulated in-process) must reproduce the full-batch forward, statistics and input gradient; d_gamma /
    d_beta stay per-rank sums that add up to the full-batch ones."""
    from onssen_b200 import _lib
    H, M = 40, 512
    Hp = _lib.hp_of(H)
    g = torch.Generator(device="cpu").manual_seed(4)
    y = (torch.randn(M, 2 * Hp, generator=g) * 1.7 + 0.3).to(cuda_device)
    d_out = torch.randn(M, 2 * Hp, generator=g).to(cuda_device)
    gamma = (torch.rand(2 * H, generator=g) + 0.5).to(cuda_device)
    beta = torch.randn(2 * H, generator=g).to(cuda_device)
    stats = lambda: (torch.zeros(2 * H, device=cuda_device), torch.ones(2 * H, device=cuda_device))
    rm, rv = stats()
    full_h, mean, invstd = _lib.bn_forward_f16(y, M, H, gamma, beta, rm, rv, 1e-5, 0.1, True, save_stats=True)
    full_dy, full_dg, full_db = _lib.bn_backward(d_out, y, M, H, gamma, mean, invstd)
    halves = [slice(0, M // 2), slice(M // 2, M)]

    def two_pass(call):
        local = []
        for sl in halves:                                   # pass 1: collect every rank's local sums
            call(sl, (lambda t: local.append(t.clone()), 2))
        total = local[0] + local[1]
        return [call(sl, (lambda t: t.copy_(total), 2)) for sl in halves]   # pass 2: 'all-reduced' sums

    outs = two_pass(lambda sl, sync: _lib.bn_forward_f16(y[sl].contiguous(), M // 2, H, gamma, beta, *stats(), 1e-5, 0.1,
                                                       True, save_stats=True, sync=sync))
    for sl, (o_h, m2, i2) in zip(halves, outs):
        assert torch.allclose(m2, mean, atol=1e-6) and torch.allclose(i2, invstd, rtol=1e-6)
        assert (o_h.float() - full_h[sl].float()).abs().max() < 4e-3          # one fp16 ulp at |x| < 8
    grads = two_pass(lambda sl, sync: _lib.bn_backward(d_out[sl].contiguous(), y[sl].contiguous(), M // 2, H, gamma, mean,
                                                       invstd, sync=sync))
    for sl, (dy, _, _) in zip(halves, grads):
        assert torch.allclose(dy, full_dy[sl], atol=2e-6, rtol=1e-5)
    assert torch.allclose(grads[0][1] + grads[1][1], full_dg, rtol=1e-5, atol=1e-5)
    assert torch.allclose(grads[0][2] + grads[1][2], full_db, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,H,T", [(20, 600, 5), (32, 600, 7), (64, 600, 5), (96, 600, 4), (5, 40, 9), (3, 8, 6),
                                   (12, 300, 6)])
def test_bptt_persistent_kernels_match_per_step_kernel(cuda_device, B, H, T):
    """differential test of the three BPTT implementations on the same random state: the tcgen05 cluster kernel
    (mode 2: 16 / 32 column slices, pad columns, a partly padded last unit block at H=600, single unit block at small
    H) and the mma.sync persistent kernel (mode 1) against the one-launch-per-step kernel (mode 0)"""
    from onssen_b200 import _lib
    lib = _lib.load()
    Hp, M = _lib.hp_of(H), T * B
    g = torch.Generator(device="cpu").manual_seed(B)
    k = 1 / np.sqrt(H)
    mk = lambda *s: ((torch.rand(*s, generator=g) * 2 - 1) * k).to(cuda_device)
    whh_t = _lib.lstm_pack_whh_t(mk(4 * H, H), mk(4 * H, H), H)
    act0 = torch.rand(M, 8 * Hp, generator=g).to(cuda_device)
    c = torch.randn(M, 2 * Hp, generator=g).to(cuda_device)
    dy = (torch.randn(M, 2 * Hp, generator=g) * 1e-3).to(cuda_device)
    sc = _lib.amax_scale(dy, target=0.0625)
    outs = []
    try:
        for mode in (0, 1, 2):
            lib.onssen_blstm_rec_bwd_set_persistent(mode)
            act = act0.clone()
            dg16 = torch.zeros(M, 8 * Hp, device=cuda_device, dtype=torch.float16)
            _lib.blstm_rec_bwd(act, dg16, c, dy, whh_t, sc, B, T, H, 0.3, 7, 1)
            torch.cuda.synchronize()
            outs.append((act, dg16))
    finally:
        lib.onssen_blstm_rec_bwd_set_persistent(2)
    a0, h0 = outs[0]
    scale = a0.abs().max().item()
    for a1, h1 in outs[1:]:
        assert scale > 0 and torch.isfinite(a1).all()
        assert (a0 - a1).abs().max().item() < 2e-4 * scale            # same operands, different summation order
        assert (h0.float() - h1.float()).abs().max().item() <= 2e-3 * h0.float().abs().max().item()


def test_second_backward_through_the_same_graph_is_refused(cuda_device):
    """BPTT overwrites the saved forward state in place: a second backward through the same graph must say so instead
    of returning garbage (and the outputs are saved through save_for_backward, not in a dict that closes a cycle)."""
    import onssen_b200 as ob
    torch.manual_seed(0)
    model = ob.nn.deep_clustering(33, 16, 2, 8, dropout=0.0).to(cuda_device).train()
    x = torch.randn(2, 12, 33, device=cuda_device)
    emb, = model([x])
    loss = emb.square().mean()
    loss.backward(retain_graph=True)
    g1 = model.fc_dc.weight.grad.clone()
    assert torch.isfinite(g1).all() and g1.abs().max() > 0
    with pytest.raises(RuntimeError, match="already consumed"):
        loss.backward()
