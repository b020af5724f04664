"""Pins oracle/fusion_head_oracle.py to the reference: the golden fixtures were produced by the reference's
own modules (oracle/make_golden.py, run where /root/reference exists).  CPU only."""
import glob
import os

import pytest
import torch

from oracle import fusion_head_oracle as O
from oracle import synth

TOL = 2e-5   # fp32 CPU vs fp32 CPU, different op order (explicit MHA vs fused kernels)


def rel(a, b):
    a, b = a.double(), b.double()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


def _train_cases(golden_dir=None):
    root = os.path.join(os.path.dirname(__file__), "golden")
    return sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(root, "train_*.pt")))


def _check_grads(gold, weights):
    checked = 0
    for key, summ in gold["grads"].items():
        grp, name = key.split("/", 1)
        g = weights[grp][name].grad
        if summ is None:                      # anchor temperature: reference leaves .grad = None
            assert g is None or float(g.abs().max()) == 0.0, key
            continue
        assert g is not None, key
        scale = summ["norm"] / max(1.0, g.numel() ** 0.5) + 1e-12
        if summ["norm"] == 0.0:               # anchor parameters: exactly-zero gradients
            assert float(g.abs().max()) == 0.0, key
            continue
        if summ["norm"] < 1e-6:               # mathematically zero (e.g. MHA key bias): rounding noise only
            assert g.double().norm().item() < 1e-5, key
            continue
        assert abs(g.double().norm().item() - summ["norm"]) <= 5 * TOL * summ["norm"], key
        probe = g.reshape(-1)[summ["idx"]]
        assert (probe.double() - summ["vals"].double()).abs().max().item() <= 50 * TOL * max(scale, summ["vals"].abs().max().item()), key
        if "full" in summ:
            assert rel(g, summ["full"]) < 10 * TOL, key
        checked += 1
    return checked


@pytest.mark.parametrize("case", _train_cases())
def test_oracle_matches_reference_train(case, golden_dir):
    gold = torch.load(os.path.join(golden_dir, f"{case}.pt"), weights_only=False)
    cfg = gold["config"]
    weights = synth.head_weights(cfg["C"], cfg["num_layers"], seed=0)
    for grp in weights.values():
        for k, v in grp.items():
            if v.is_floating_point() and k not in synth.CLASSIFIER_BUFFERS:
                v.requires_grad_(True)
    a, t, am, tm, labels = synth.make_inputs(cfg["B"], cfg["Ta"], cfg["Tt"], cfg["C"], cfg["seed"], cfg["with_masks"])
    out = O.head_forward(a, t, am, tm, labels, weights, cfg["C"], cfg["num_layers"])
    out["loss"].backward()

    for k in ("logits", "unc", "fused", "a_vec", "t_vec"):
        assert rel(out[k], gold[k]) < TOL, k
    assert rel(out["a_enh"][:, :4], gold["a_enh_head"]) < TOL
    assert rel(out["t_enh"][:, :4], gold["t_enh_head"]) < TOL
    assert abs(out["a_enh"].double().norm().item() - gold["a_enh_norm"]) / gold["a_enh_norm"] < TOL
    for k, gk in (("ce", "ce"), ("focal", "focal"), ("unc_loss", "unc_loss"), ("proto", "proto"), ("loss", "loss")):
        assert abs(out[k].item() - gold[gk]) <= TOL * max(1.0, abs(gold[gk])), k
    assert out["anchor"].item() == 0.0 and gold["anchor"] == 0.0
    # bit-exact argmax
    assert torch.equal(out["logits"].argmax(1), gold["logits"].argmax(1))

    assert _check_grads(gold, weights) > 300


def test_oracle_matches_reference_mixed_dims(golden_dir):
    """CrossModalAttention(768, 1024) / FusionLayer(1536, 2048, 512) of the reference (unequal audio / text widths:
    cross_attention.py:7-30, fusion.py:6-16) against the oracle's restatement on the fixture's own inputs."""
    gold = torch.load(os.path.join(golden_dir, "mixed_dims_small.pt"), weights_only=False)
    cfg, ins = gold["config"], gold["inputs"]
    weights = {"cross": synth.cross_weights(audio_dim=cfg["audio_dim"], text_dim=cfg["text_dim"]),
               "fusion": synth.fusion_weights(audio_dim=2 * cfg["audio_dim"], text_dim=2 * cfg["text_dim"])}
    for grp in weights.values():
        for v in grp.values():
            v.requires_grad_(True)
    x = {k: ins[k].clone().requires_grad_(True) for k in ("a", "t", "av", "tv")}
    ea, et = O.cross_attention(x["a"], x["t"], ins["a_mask"], ins["t_mask"], weights["cross"])
    fused = O.fusion(x["av"], x["tv"], weights["fusion"])
    ((ea * ins["ua"]).sum() + (et * ins["ut"]).sum() + (fused * ins["up"]).sum()).backward()
    assert rel(ea, gold["audio_enh"]) < TOL and rel(et, gold["text_enh"]) < TOL and rel(fused, gold["fused"]) < TOL
    for k in ("a", "t", "av", "tv"):
        assert rel(x[k].grad, gold["din"][k]) < 10 * TOL, k
    assert _check_grads(gold, weights) >= 40


def test_oracle_matches_reference_train_dropout(golden_dir):
    """Training mode with every dropout ACTIVE: the fixture holds the keep masks the reference drew (recorded through a
    patched torch.nn.functional.dropout, oracle/make_golden.py); the oracle, given the same masks at its named sites,
    must reproduce outputs and gradients -- this pins WHERE each dropout acts (attention weights after the softmax,
    branch output before the residual add, post-ReLU hidden units, post-dropout features feeding both heads)."""
    gold = torch.load(os.path.join(golden_dir, "dropout_train_small.pt"), weights_only=False)
    cfg = gold["config"]
    weights = synth.head_weights(cfg["C"], cfg["num_layers"], seed=0)
    for grp in weights.values():
        for k, v in grp.items():
            if v.is_floating_point() and k not in synth.CLASSIFIER_BUFFERS:
                v.requires_grad_(True)
    a, t, am, tm, labels = synth.make_inputs(cfg["B"], cfg["Ta"], cfg["Tt"], cfg["C"], cfg["seed"], cfg["with_masks"])
    masks = {k: v["keep"].float() / (1.0 - v["p"]) for k, v in gold["masks"].items()}
    assert len(masks) == 4 + 2 + 1 + 2 * cfg["num_layers"] + 1 + 2
    with O.dropout_masks(masks):
        out = O.head_forward(a, t, am, tm, labels, weights, cfg["C"], cfg["num_layers"])
        out["loss"].backward()
    for k in ("logits", "unc", "fused", "a_vec", "t_vec"):
        assert rel(out[k], gold[k]) < TOL, k
    assert abs(out["a_enh"].double().norm().item() - gold["a_enh_norm"]) / gold["a_enh_norm"] < TOL
    assert abs(out["t_enh"].double().norm().item() - gold["t_enh_norm"]) / gold["t_enh_norm"] < TOL
    for k in ("ce", "focal", "unc_loss", "proto", "loss"):
        assert abs(out[k].item() - gold[k]) <= TOL * max(1.0, abs(gold[k])), k
    assert torch.equal(out["logits"].argmax(1), gold["logits"].argmax(1))
    assert _check_grads(gold, weights) > 300
    # and the masks matter: without them the same inputs give a different loss
    with torch.no_grad():
        plain = O.head_forward(a, t, am, tm, labels, weights, cfg["C"], cfg["num_layers"])
    assert abs(plain["loss"].item() - gold["loss"]) > 1e-3


def test_oracle_matches_reference_eval(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "eval_cfg5_small.pt"), weights_only=False)
    cfg = gold["config"]
    C = cfg["C"]
    w = synth.classifier_weights(C, 35, seed=0)
    g = torch.Generator().manual_seed(cfg["seed"])
    val_fused = torch.randn(64, 512, generator=g)
    val_labels = torch.randint(0, C, (64,), generator=g)
    with torch.no_grad():
        f = O.classifier_features(val_fused, w)
        assert rel(f, gold["val_features"]) < TOL
        w.update(O.fit_weibull(f, val_labels, C, w))
        for k in synth.CLASSIFIER_BUFFERS:
            assert rel(w[k], gold["weibull"][k]) < 1e-4, k
        fused_views = torch.randn(cfg["views"], cfg["B"], 512, generator=g)
        labels = torch.randint(0, C, (cfg["B"],), generator=g)
        lv = torch.stack([O.classifier(fused_views[v], w, use_openmax=True, training=False) for v in range(cfg["views"])])
        lp = torch.stack([O.classifier(fused_views[v], w, use_openmax=False) for v in range(cfg["views"])])
        assert rel(lv, gold["logits_views"]) < 1e-4
        assert rel(lp, gold["logits_plain"]) < TOL
        mean_logits = O.tta_mean(lv)
        T = O.find_optimal_temperature(lp[0], labels)
        assert abs(T - gold["temperature"]) < 1e-4 * gold["temperature"]
        scaled = mean_logits / T
        probs = torch.softmax(scaled, -1)
        assert rel(probs, gold["probs"]) < 1e-4
        assert torch.equal(probs.argmax(-1), gold["preds"])
        assert rel(O.energy_score(scaled), gold["energy"]) < 1e-4


def test_oracle_matches_reference_supcon(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "supcon_small.pt"), weights_only=False)
    cfg = gold["config"]
    g = torch.Generator().manual_seed(cfg["seed"])
    f = (torch.randn(cfg["B"], cfg["D"], generator=g) * 2.0).requires_grad_(True)
    labels = torch.randint(0, cfg["C"], (cfg["B"],), generator=g)
    labels[-1] = cfg["C"]
    assert torch.equal(labels, gold["labels"])
    loss = O.supcon_loss(f, labels, cfg["temperature"])
    loss.backward()
    assert abs(loss.item() - gold["loss"]) <= TOL * abs(gold["loss"])
    assert rel(f.grad, gold["grad"]) < 10 * TOL


@pytest.mark.parametrize("name", ["feature_fusion_combined", "feature_fusion_asr"])
def test_oracle_matches_reference_feature_fusion(golden_dir, name):
    """the per-utterance feature fusion of the encoders, eval and train (recorded masks)."""
    gold = torch.load(os.path.join(golden_dir, f"{name}.pt"), weights_only=False)
    cfg = gold["config"]
    w = synth.feature_fusion_weights(cfg["attr"], cfg["F"], seed=0, hidden=cfg["hid"])
    g = torch.Generator().manual_seed(cfg["seed"])
    seq = torch.randn(cfg["B"], cfg["T"], cfg["hid"], generator=g)
    feats = torch.rand(cfg["B"], cfg["F"], generator=g) * 2.0 - 0.5
    up = torch.randn(cfg["B"], cfg["T"], cfg["hid"], generator=g)
    for mode in ("eval", "train"):
        ws = {k: v.clone().requires_grad_(True) for k, v in w.items()}
        x = seq.clone().requires_grad_(True)
        masks = {"feat.out": gold["train"]["mask"]} if mode == "train" else {}
        with O.dropout_masks(masks):
            y = O.utterance_feature_fusion(x, feats, ws)
        (y * up).sum().backward()
        ref = gold[mode]
        assert rel(y, ref["y"]) < TOL
        assert rel(x.grad, ref["dx"]) < 10 * TOL
        assert rel(ws["0.weight"].grad, ref["dw"]) < 10 * TOL
        assert rel(ws["0.bias"].grad, ref["db"]) < 10 * TOL


def test_feature_fusion_decomposition_identity():
    """The algebra csrc/featfuse.cu relies on: with features constant over the frames of an utterance, the concatenated
    Linear equals a 768-wide GEMM plus a per-utterance bias c[u] = W[:, D:] f[u] + b, and the weight gradient splits into
    dW[:, :D] = dz^T x, dW[:, D:] = (sum_t dz[u, t])^T f, db = sum dz.  Checked in fp64 against the oracle's
    concatenation form."""
    torch.manual_seed(3)
    B, T, D, F = 3, 5, 16, 4
    x = torch.randn(B, T, D, dtype=torch.float64)
    f = torch.randn(B, F, dtype=torch.float64)
    up = torch.randn(B, T, D, dtype=torch.float64)
    w = {"0.weight": torch.randn(D, D + F, dtype=torch.float64).requires_grad_(True),
         "0.bias": torch.randn(D, dtype=torch.float64).requires_grad_(True)}
    xr = x.clone().requires_grad_(True)
    y = O.utterance_feature_fusion(xr, f, w)
    (y * up).sum().backward()
    W, b = w["0.weight"].detach(), w["0.bias"].detach()
    c = f @ W[:, D:].T + b                                     # [B, D]
    z = x @ W[:, :D].T + c[:, None, :]
    y2 = torch.relu(z)
    dz = up * (y2 > 0)
    assert torch.allclose(y2, y.detach(), atol=1e-12)
    assert torch.allclose(dz @ W[:, :D], xr.grad, atol=1e-12)
    dW = w["0.weight"].grad
    assert torch.allclose(dz.reshape(-1, D).T @ x.reshape(-1, D), dW[:, :D], atol=1e-12)
    assert torch.allclose(dz.sum(1).T @ f, dW[:, D:], atol=1e-12)
    assert torch.allclose(dz.sum((0, 1)), w["0.bias"].grad, atol=1e-12)


def test_oracle_matches_reference_late_ood(golden_dir):
    """energy + prototype-distance OOD scoring against the reference's LateStageOODDetector."""
    gold = torch.load(os.path.join(golden_dir, "late_ood_small.pt"), weights_only=False)
    s = O.late_ood_scores(gold["logits"], gold["features"], gold["state"])
    assert rel(s["energy"], gold["energy"]) < TOL
    assert rel(s["distances"], gold["distances"]) < TOL and rel(s["min_distance"], gold["min_distance"]) < TOL
    r = gold["result"]
    assert abs(s["energy"].mean().item() - r["energy_score"]) <= TOL * abs(r["energy_score"])
    assert abs(s["min_distance"].mean().item() - r["prototype_distance"]) <= TOL * abs(r["prototype_distance"])
    assert abs(s["combined"].mean().item() - r["combined_score"]) <= TOL * abs(r["combined_score"])
    assert bool((s["combined"] < 0.5).any()) == r["is_ood"]


def test_quirks():
    """SURVEY.md 8(a) quirks the CUDA path must reproduce."""
    torch.manual_seed(0)
    C = 4
    w = synth.classifier_weights(C, 2, seed=0)
    f = torch.randn(8, 256)
    _, al = O.anchor_clustering(f, w)
    assert al.item() == 0.0
    # uncertainty term = mean(unc) * mean(correct)  ([B,1]*[B] broadcast)
    unc = torch.rand(8, 1)
    corr = (torch.rand(8) > 0.5).float()
    assert abs((unc * corr).mean().item() - unc.mean().item() * corr.mean().item()) < 1e-6
    # fully padded key row -> NaN
    cw = synth.cross_weights(0)
    a = torch.randn(2, 5, 768); t = torch.randn(2, 3, 768)
    tm = torch.tensor([[1., 1., 0.], [0., 0., 0.]])
    ae, te = O.cross_attention(a, t, None, tm, cw)
    assert torch.isfinite(ae[0]).all() and torch.isnan(ae[1]).all()
    pw = synth.pool_weights("pool_a", 0)
    pooled = O.attentive_stats_pooling(torch.randn(2, 3, 768), tm, pw)
    assert torch.isfinite(pooled[0]).all() and torch.isnan(pooled[1]).all()
