"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only).

The reference (/root/reference) is imported unmodified; `librosa` and `attrdict` (not installed here) are
replaced by EMPTY placeholder modules solely to get past two import lines (onssen/data/feature_utils.py:1,
onssen/utils/train.py:1) -- they contain no arithmetic and nothing below touches onssen.data / onssen.utils.
Fixtures pin: nn.deep_clustering / nn.chimera / nn.enhance forward (eval and train-mode BN) and
loss.loss_dc / loss_chimera_msa / loss_chimera_psa / loss_mask_msa / loss_mask_psa.

feat_*.npz pin the featurizer helpers (rows a2-a7) and sdr.npz pins SI-SDR, both from the live reference too.

Usage:  python oracle/make_golden.py [--only models|phase|feat|sdr]
"""
import os
import sys

import numpy as np

REF = "/root/reference"


def import_reference():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import ref_loader
    return ref_loader.import_reference()


def sd_to_np(sd):
    return {k: v.detach().cpu().numpy().copy() for k, v in sd.items() if "num_batches" not in k}


def rand_bn(model, torch):
    with torch.no_grad():
        model.bn.running_mean.normal_(0, 0.2)
        model.bn.running_var.uniform_(0.5, 1.5)
        model.bn.weight.uniform_(0.5, 1.5)
        model.bn.bias.normal_(0, 0.2)


def make_labels(rng, B, T, F):
    mag1 = np.abs(rng.standard_normal((B, T, F))).astype(np.float32) * rng.uniform(0.1, 2, (B, 1, F)).astype(np.float32)
    mag2 = np.abs(rng.standard_normal((B, T, F))).astype(np.float32) * rng.uniform(0.1, 2, (B, 1, F)).astype(np.float32)
    cos1 = np.cos(rng.uniform(-np.pi, np.pi, (B, T, F))).astype(np.float32)
    cos2 = np.cos(rng.uniform(-np.pi, np.pi, (B, T, F))).astype(np.float32)
    mix = (mag1 + mag2 * rng.uniform(0.5, 1.0, (B, T, F))).astype(np.float32)
    feat = np.log10(mix + 1e-7).astype(np.float32)
    oh = np.zeros((B, T, F, 2))
    first = mag1 >= mag2
    oh[..., 0] = first
    oh[..., 1] = ~first
    for b in range(B):
        oh[b][feat[b] < feat[b].max() - 0.6] = 0      # a tight threshold so that silence actually occurs
    return feat, oh, mix, mag1, mag2, cos1, cos2


def main():
    import torch
    onssen = import_reference()
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.manual_seed(20260924)
    rng = np.random.RandomState(7)
    cases = {"small": dict(B=3, T=16, F=9, H=8, L=2, D=4), "mid": dict(B=2, T=40, F=33, H=40, L=3, D=20)}
    for name, c in cases.items():
        B, T, F, H, L, D = (c[k] for k in "BTFHLD")
        feat, oh, mix, mag1, mag2, cos1, cos2 = make_labels(rng, B, T, F)
        tt = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        # ---- deep clustering
        dc = onssen.nn.deep_clustering(F, H, L, D, dropout=0.0)
        rand_bn(dc, torch)
        sd = sd_to_np(dc.state_dict())
        with torch.no_grad():
            dc.eval()
            emb_eval, = dc([tt(feat)])
            loss_eval = onssen.loss.loss_dc([emb_eval], [tt(oh), tt(mix)])
            dc.train()
            emb_train, = dc([tt(feat)])
            loss_train = onssen.loss.loss_dc([emb_train], [tt(oh), tt(mix)])
            sd_after = sd_to_np(dc.state_dict())
        # gradients of torch.mean(loss_dc) in train mode (train.py:77-82) w.r.t. every parameter and the embedding
        dc.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()}, strict=False)
        dc.train()
        dc.zero_grad()
        emb_g, = dc([tt(feat)])
        emb_g.retain_grad()
        torch.mean(onssen.loss.loss_dc([emb_g], [tt(oh), tt(mix)])).backward()
        grads = {"g:" + k: v.grad.detach().numpy().copy() for k, v in dc.named_parameters()}
        grads["g:embedding"] = emb_g.grad.detach().numpy().copy()
        np.savez_compressed(os.path.join(out_dir, f"dcgrad_{name}.npz"), **grads)
        np.savez_compressed(os.path.join(out_dir, f"dc_{name}.npz"), cfg=np.array([B, T, F, H, L, D]), feature=feat,
                            one_hot=oh, mag_mix=mix, emb_eval=emb_eval.numpy(), loss_eval=loss_eval.numpy(),
                            emb_train=emb_train.numpy(), loss_train=loss_train.numpy(),
                            bn_rm_after=sd_after["bn.running_mean"], bn_rv_after=sd_after["bn.running_var"],
                            **{"p:" + k: v for k, v in sd.items()})
        # ---- chimera / chimera++
        ch = onssen.nn.chimera(F, H, L, D, dropout=0.0).eval()
        sd = sd_to_np(ch.state_dict())
        with torch.no_grad():
            e, ma, mb = ch([tt(feat)])
            l_msa = onssen.loss.loss_chimera_msa([e, ma, mb], [tt(oh), tt(mix), tt(mag1), tt(mag2)])
            l_psa = onssen.loss.loss_chimera_psa([e, ma, mb], [tt(oh), tt(mix), tt(mag1), tt(mag2), tt(cos1), tt(cos2)])
        # gradients of torch.mean(loss_chimera_psa) (train mode == eval mode here: no BN, dropout 0)
        ch.zero_grad()
        e_g, ma_g, mb_g = ch([tt(feat)])
        torch.mean(onssen.loss.loss_chimera_psa([e_g, ma_g, mb_g], [tt(oh), tt(mix), tt(mag1), tt(mag2), tt(cos1),
                                                                    tt(cos2)])).backward()
        np.savez_compressed(os.path.join(out_dir, f"chimeragrad_{name}.npz"),
                            **{"g:" + k: v.grad.detach().numpy().copy() for k, v in ch.named_parameters()})
        np.savez_compressed(os.path.join(out_dir, f"chimera_{name}.npz"), cfg=np.array([B, T, F, H, L, D]), feature=feat,
                            one_hot=oh, mag_mix=mix, mag_s1=mag1, mag_s2=mag2, cos_s1=cos1, cos_s2=cos2,
                            emb=e.numpy(), mask_a=ma.numpy(), mask_b=mb.numpy(), loss_msa=l_msa.numpy(),
                            loss_psa=l_psa.numpy(), **{"p:" + k: v for k, v in sd.items()})
        # ---- enhancement + restoration layers
        en = onssen.nn.enhance(F, H, L, dropout=0.0)
        rand_bn(en, torch)
        sd = sd_to_np(en.state_dict())
        with torch.no_grad():
            en.eval()
            clean_eval, = en([tt(feat), tt(mix)])
            en.train()
            clean_train, = en([tt(feat), tt(mix)])
            l_msa = onssen.loss.loss_mask_msa([clean_eval], [tt(mag1), tt(cos1)])
            l_psa = onssen.loss.loss_mask_psa([torch.sigmoid(clean_eval)], [tt(mix), tt(mag1), tt(cos1)])
        # gradients of loss_mask_msa (scalar MSE) in train mode
        en.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()}, strict=False)
        en.train()
        en.zero_grad()
        cl_g, = en([tt(feat), tt(mix)])
        onssen.loss.loss_mask_msa([cl_g], [tt(mag1), tt(cos1)]).backward()
        np.savez_compressed(os.path.join(out_dir, f"enhancegrad_{name}.npz"),
                            **{"g:" + k: v.grad.detach().numpy().copy() for k, v in en.named_parameters()})
        np.savez_compressed(os.path.join(out_dir, f"enhance_{name}.npz"), cfg=np.array([B, T, F, H, L, D]), feature=feat,
                            mag_noisy=mix, mag_clean=mag1, cos_diff=cos1, clean_eval=clean_eval.numpy(),
                            clean_train=clean_train.numpy(), loss_msa=l_msa.numpy(), loss_psa=l_psa.numpy(),
                            **{"p:" + k: v for k, v in sd.items()})
        print("wrote", name)


def main_phase():
    """phase_net / loss_phase fixtures.  The reference classes raise as written (phase_network.py:28,
    loss_phase.py:7-13), so these come from a torch RESTATEMENT WITH THE REPAIRS of SURVEY.md 8a-14/a18 built on the
    live reference chimera + loss_dc: they pin our CUDA path and the numpy oracle to each other and to torch autograd,
    not to the (non-running) reference -> "parity unpinned" for these two rows."""
    import torch
    onssen = import_reference()
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
    torch.manual_seed(20260925)
    rng = np.random.RandomState(11)

    from oracle.phase_repaired import build as build_phase
    PhaseNetRepaired, loss_phase_repaired = build_phase(onssen)

    cases = {"small": dict(B=3, T=16, F=9, H=8, L=2, D=4), "mid": dict(B=2, T=24, F=33, H=40, L=2, D=20)}
    for name, c in cases.items():
        B, T, F, H, L, D = (c[k] for k in "BTFHLD")
        feat, oh, mix, mag1, mag2, _, _ = make_labels(rng, B, T, F)
        x_phase = (rng.standard_normal((B, T, F, 2)) * mix[..., None]).astype(np.float32)
        ph1 = (rng.standard_normal((B, T, F, 2)) * mag1[..., None]).astype(np.float32)
        ph2 = (rng.standard_normal((B, T, F, 2)) * mag2[..., None]).astype(np.float32)
        tt = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        net = PhaseNetRepaired(F, H, L, D)
        rand_bn(net, torch)
        sd = sd_to_np(net.state_dict())
        net.train()
        out = net([tt(feat), tt(x_phase)])
        loss = loss_phase_repaired(out, [tt(oh), tt(mix), tt(mag1), tt(mag2), tt(ph1), tt(ph2)])
        torch.mean(loss).backward()
        np.savez_compressed(os.path.join(out_dir, f"phasegrad_{name}.npz"),
                            **{"g:" + k: v.grad.detach().numpy().copy() for k, v in net.named_parameters()})
        np.savez_compressed(os.path.join(out_dir, f"phase_{name}.npz"), cfg=np.array([B, T, F, H, L, D]), feature=feat,
                            x_phase=x_phase, one_hot=oh, mag_mix=mix, mag_s1=mag1, mag_s2=mag2, phase_s1=ph1,
                            phase_s2=ph2, emb=out[0].detach().numpy(), mask_a=out[1].detach().numpy(),
                            mask_b=out[2].detach().numpy(), phase_a=out[3].detach().numpy(),
                            phase_b=out[4].detach().numpy(), loss=loss.detach().numpy(),
                            norm_a=net.pre_norms[0].numpy(), norm_b=net.pre_norms[1].numpy(),
                            **{"p:" + k: v for k, v in sd.items()})
        print("wrote phase", name)


class _Opts(dict):
    __getattr__ = dict.__getitem__


def main_feat():
    """Rows a2-a7: the reference's OWN `wsj0_2mix_dataset.get_feature` (onssen/data/wsj0_2mix.py:103-158: tile,
    np.random crop, get_log_magnitude, np.abs, get_one_hot + VAD, get_cos_difference, get_phase) run live on synthetic
    waveforms.  The only substituted piece is `get_stft` (librosa, absent) -> oracle.stft of the in-memory signal
    (oracle/ref_loader.patch_stft).  Stored: waveforms, the numpy seed, the crop start the reference drew, and every
    returned tensor, for the 256/64 case (frames > frame_length) and the 1024/256 tiling case (frames <= frame_length)."""
    import torch  # noqa: F401
    onssen = import_reference()
    from oracle import ref_loader, onssen_oracle as O
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
    cases = {"crop": dict(nsample=6400, n_fft=256, hop=64, T=60, seed=314),        # 101 frames -> crop 60
             "tile": dict(nsample=7000, n_fft=1024, hop=256, T=40, seed=2718)}     # 28 frames <= 40 -> tile x2 -> crop
    for name, c in cases.items():
        mix, s1, s2 = O.synth_utterance(4000 + c["seed"], c["nsample"])
        sigs = {"/x/mix/u.wav": mix, "/x/s1/u.wav": s1, "/x/s2/u.wav": s2}
        ref_loader.patch_stft(onssen, sigs)
        out = dict(mix=mix, s1=s1, s2=s2, cfg=np.array([c["nsample"], c["n_fft"], c["hop"], c["T"], c["seed"]]))
        got = {}
        for model_name in ("dc", "chimera", "chimera++", "phase"):
            opts = _Opts(data_path="/nonexistent", batch_size=1, frame_length=c["T"], sampling_rate=8000,
                         window_size=c["n_fft"], hop_size=c["hop"], db_threshold=40)
            ds = onssen.data.wsj0_2mix.wsj0_2mix_dataset(model_name, opts, "tr")
            np.random.seed(c["seed"])
            inp, lab = ds.get_feature("/x/mix/u.wav")
            got[model_name] = ([t.numpy() for t in inp], [t.numpy() for t in lab])
        # the four label layouts share their leading entries (wsj0_2mix.py:137-152): store each array once
        full_in, full_lab = got["chimera++"]
        same = lambda a, b: a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
        assert len(got["dc"][1]) == 2 and len(got["chimera"][1]) == 4 and len(got["phase"][1]) == 6
        assert all(same(a, b) for a, b in zip(got["dc"][1], full_lab[:2])) and same(got["dc"][0][0], full_in[0])
        assert all(same(a, b) for a, b in zip(got["chimera"][1], full_lab[:4]))
        assert all(same(a, b) for a, b in zip(got["phase"][1][:4], full_lab[:4])) and same(got["phase"][0][0], full_in[0])
        for k, a in zip(("feature", "one_hot", "mag_mix", "mag_s1", "mag_s2", "cos_s1", "cos_s2"), full_in + full_lab):
            out[k] = a
        out["phase_mix"], out["phase_s1"], out["phase_s2"] = got["phase"][0][1], got["phase"][1][4], got["phase"][1][5]
        np.random.seed(c["seed"])
        out["crop_start"] = np.array(np.random.randint(O.num_crop_starts(c["nsample"], c["hop"], c["T"])))
        np.savez_compressed(os.path.join(out_dir, f"feat_{name}.npz"), **out)
        print("wrote feat", name, "crop_start", int(out["crop_start"]))


def main_sdr():
    """batch_SDR_torch (onssen/evaluate/sdr.py:40-87) on seeded signals: pins utils.test.batch_si_sdr."""
    import torch
    onssen = import_reference()
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
    rng = np.random.RandomState(99)
    B, S, n = 5, 2, 2000
    ref = rng.standard_normal((B, S, n)).astype(np.float32)
    est = ref[:, ::-1].copy() * rng.uniform(0.3, 2, (B, S, 1)).astype(np.float32)
    est[::2] = ref[::2]                                     # some batches un-permuted
    est += rng.standard_normal((B, S, n)).astype(np.float32) * rng.uniform(0.01, 1.0, (B, 1, 1)).astype(np.float32)
    sdr, perm = onssen.evaluate.batch_SDR_torch(torch.from_numpy(est), torch.from_numpy(ref), return_perm=True)
    np.savez_compressed(os.path.join(out_dir, "sdr.npz"), est=est, ref=ref, sdr=sdr.numpy(), perm=perm.numpy())
    print("wrote sdr", sdr.numpy())


if __name__ == "__main__":
    # `--only X` regenerates one family (each is seeded independently, the others stay bit-identical)
    only = sys.argv[2] if sys.argv[1:2] == ["--only"] else None
    if only in (None, "models"):
        main()
    if only in (None, "phase"):
        main_phase()
    if only in (None, "feat"):
        main_feat()
    if only in (None, "sdr"):
        main_sdr()
