"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules (imported by file path from
/root/reference, CPU fp32) on the deterministic synthetic weights/inputs of oracle/synth.py.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
The fixtures hold outputs and gradient summaries -- never weights (those are regenerated from
synth.py) -- so they stay small.  tests/test_oracle_golden.py pins oracle/fusion_head_oracle.py to them.
"""
from __future__ import annotations

import importlib.util
import os
import re
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

REF = os.environ.get("SER_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def load_ref(name: str):
    """Load src/models/<name>.py by path; the package __init__ needs librosa."""
    path = os.path.join(REF, "src", "models", f"{name}.py")
    spec = importlib.util.spec_from_file_location(f"ref_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_reference_head(C: int, num_layers: int, weights):
    ca = load_ref("cross_attention"); po = load_ref("pooling"); fu = load_ref("fusion")
    cl = load_ref("classifier"); pr = load_ref("prototypes"); lo = load_ref("losses")
    m = {
        # adapters exactly as src/models/audio_encoder.py:19-21 / text_encoder.py:17-19
        "adapter_a": nn.Sequential(nn.Linear(768, 256), nn.ReLU(), nn.Linear(256, 768)),
        "adapter_t": nn.Sequential(nn.Linear(768, 256), nn.ReLU(), nn.Linear(256, 768)),
        "cross": ca.CrossModalAttention(768, 768, shared_dim=256, num_heads=8),
        "pool_a": po.AttentiveStatsPooling(768),
        "pool_t": po.AttentiveStatsPooling(768),
        "fusion": fu.FusionLayer(1536, 1536, 512),
        "classifier": cl.AdvancedOpenMaxClassifier(input_dim=512, num_labels=C, num_layers=num_layers, base_dim=512,
                                                   dropout=0.15),
        "prototypes": pr.PrototypeMemory(C, 512),
    }
    for k, mod in m.items():
        missing = mod.load_state_dict(weights[k], strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        mod.eval()      # dropout off; nothing else differs between train/eval on this path
    losses = {"ce": lo.LabelSmoothingCrossEntropy(0.1),
              "focal": lo.ClassBalancedFocalLoss(beta=0.9999, gamma=2.0, num_classes=C)}
    return m, losses


def summarize(t: torch.Tensor, name: str):
    flat = t.detach().reshape(-1).double()
    g = torch.Generator().manual_seed(abs(hash(name)) % (2 ** 31))
    n = flat.numel()
    # fixed pseudo-random probe positions (stored, so the hash seed does not matter to readers)
    idx = torch.randint(0, n, (min(32, n),), generator=g)
    out = {"shape": tuple(t.shape), "norm": flat.norm().item(), "sum": flat.sum().item(),
           "idx": idx, "vals": flat[idx].float()}
    m = re.search(r"(residual_layers|layer_norms)\.(\d+)\.", name)
    keep_full = n <= 4096 and (m is None or int(m.group(2)) in (0, 17, 34))
    if keep_full:
        out["full"] = t.detach().clone().float()
    return out


def run_train_case(name, B, Ta, Tt, C, seed, with_masks=True, num_layers=35):
    weights = synth.head_weights(C, num_layers, seed=0)
    m, L = build_reference_head(C, num_layers, weights)
    a_hid, t_hid, a_mask, t_mask, labels = synth.make_inputs(B, Ta, Tt, C, seed, with_masks)
    # src/train.py:145-168 (adapters applied as in the encoders' forward)
    a_seq = a_hid + m["adapter_a"](a_hid)
    t_seq = t_hid + m["adapter_t"](t_hid)
    a_enh, t_enh = m["cross"](a_seq, t_seq, a_mask, t_mask)
    a_vec = m["pool_a"](a_enh, a_mask)
    t_vec = m["pool_t"](t_enh, t_mask)
    fused = m["fusion"](a_vec, t_vec)
    logits, unc, anchor_loss = m["classifier"](fused, use_openmax=False, return_uncertainty=True)
    ce = L["ce"](logits, labels)
    focal = L["focal"](logits, labels)
    loss = ce + 0.3 * focal
    loss = loss + 0.1 * anchor_loss
    unc_loss = torch.mean(unc * (labels == logits.argmax(dim=1)).float())
    loss = loss + 0.05 * unc_loss
    proto = m["prototypes"].prototype_loss(fused, labels)
    loss = loss + 0.01 * proto
    loss.backward()

    grads = {}
    for k, mod in m.items():
        for pn, p in mod.named_parameters():
            grads[f"{k}/{pn}"] = None if p.grad is None else summarize(p.grad, f"{k}/{pn}")
    gold = {
        "config": dict(B=B, Ta=Ta, Tt=Tt, C=C, seed=seed, with_masks=with_masks, num_layers=num_layers),
        "logits": logits.detach(), "unc": unc.detach(), "fused": fused.detach(),
        "a_vec": a_vec.detach(), "t_vec": t_vec.detach(),
        "a_enh_head": a_enh[:, :4].detach().clone(), "t_enh_head": t_enh[:, :4].detach().clone(),
        "a_enh_norm": a_enh.detach().double().norm().item(), "t_enh_norm": t_enh.detach().double().norm().item(),
        "ce": ce.item(), "focal": focal.item(), "anchor": anchor_loss.item(), "unc_loss": unc_loss.item(),
        "proto": proto.item(), "loss": loss.item(),
        "grads": grads,
        "torch": torch.__version__,
    }
    torch.save(gold, os.path.join(OUT, f"{name}.pt"))
    print(f"{name}: loss={loss.item():.6f} ce={ce.item():.6f} focal={focal.item():.6f} proto={proto.item():.6f} "
          f"unc={unc_loss.item():.6f} anchor={anchor_loss.item()}")


class RecordedDropout:
    """Replaces torch.nn.functional.dropout while the reference runs in train() mode: every call draws its keep mask
    from a seeded generator, applies x * keep / (1 - p) (the definition of dropout) and records (shape, p, keep) in
    call order.  nn.Dropout.forward and nn.MultiheadAttention's math path both resolve `dropout` in
    torch.nn.functional at call time, so this reaches every dropout of the path."""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)
        self.calls = []

    def __call__(self, input, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return input
        keep = torch.rand(input.shape, generator=self.gen) >= p
        self.calls.append((tuple(input.shape), float(p), keep))
        return input * keep.to(input.dtype) / (1.0 - p)

    def __enter__(self):
        import torch.nn.functional as F
        self.F, self.orig = F, F.dropout
        F.dropout = self
        return self

    def __exit__(self, *exc):
        self.F.dropout = self.orig


def dropout_site_names(num_layers):
    """Call order of the dropouts in one reference training forward (src/train.py:145-168): cross attention
    (attn_a weights, dropout(a_out), attn_t weights, dropout(t_out)), fusion (proj_a[2], proj_t[2]), classifier
    (input_projection[3], per block block[3] / block[5], output_projection[3], then -- classifier.py:221-229 --
    the anchor projection's dropout, whose output is discarded, and the uncertainty head's)."""
    names = ["cross.prob_a", "cross.res_a", "cross.prob_t", "cross.res_t", "fusion.a", "fusion.t", "clf.in"]
    for i in range(num_layers):
        names += [f"clf.block{i}.hidden", f"clf.block{i}.out"]
    names += ["clf.out"]
    return names


def run_train_dropout_case(name, B, Ta, Tt, C, seed, num_layers=35):
    """Training-mode forward/backward of the reference with every dropout ACTIVE and its masks recorded: pins where
    the oracle applies which mask (tests/test_oracle_golden.py::test_oracle_matches_reference_train_dropout)."""
    weights = synth.head_weights(C, num_layers, seed=0)
    m, L = build_reference_head(C, num_layers, weights)
    for mod in m.values():
        mod.train()
    a_hid, t_hid, a_mask, t_mask, labels = synth.make_inputs(B, Ta, Tt, C, seed, True)
    with RecordedDropout(seed) as rec:
        a_seq = a_hid + m["adapter_a"](a_hid)
        t_seq = t_hid + m["adapter_t"](t_hid)
        a_enh, t_enh = m["cross"](a_seq, t_seq, a_mask, t_mask)
        a_vec = m["pool_a"](a_enh, a_mask)
        t_vec = m["pool_t"](t_enh, t_mask)
        fused = m["fusion"](a_vec, t_vec)
        logits, unc, anchor_loss = m["classifier"](fused, use_openmax=False, return_uncertainty=True)
        ce = L["ce"](logits, labels)
        focal = L["focal"](logits, labels)
        unc_loss = torch.mean(unc * (labels == logits.argmax(dim=1)).float())
        proto = m["prototypes"].prototype_loss(fused, labels)
        loss = ce + 0.3 * focal + 0.1 * anchor_loss + 0.05 * unc_loss + 0.01 * proto
        loss.backward()
    names = dropout_site_names(num_layers)
    tail = rec.calls[len(names):]
    # what follows output_projection[3] is identified by shape: [B,128] anchor projection (discarded), [B,64] uncertainty head
    tail_names = {128: "anchor (unused)", 64: "clf.unc"}
    assert len(tail) == 2 and sorted(c[0][1] for c in tail) == [64, 128], [c[0] for c in tail]
    names += [tail_names[c[0][1]] for c in tail]
    assert len(names) == len(rec.calls)
    masks = {}
    for n, (shape, p, keep) in zip(names, rec.calls):
        if n.startswith("cross.prob"):
            keep = keep.view(B, 8, shape[1], shape[2])          # MHA flattens (batch, head) -> b * H + h
        masks[n] = {"p": p, "keep": keep.to(torch.uint8)}
        print(f"  dropout site {n:24s} shape {tuple(keep.shape)} p={p}")
    grads = {}
    for k, mod in m.items():
        for pn, p in mod.named_parameters():
            grads[f"{k}/{pn}"] = None if p.grad is None else summarize(p.grad, f"{k}/{pn}")
    gold = {
        "config": dict(B=B, Ta=Ta, Tt=Tt, C=C, seed=seed, with_masks=True, num_layers=num_layers),
        "masks": masks,
        "logits": logits.detach(), "unc": unc.detach(), "fused": fused.detach(),
        "a_vec": a_vec.detach(), "t_vec": t_vec.detach(),
        "a_enh_norm": a_enh.detach().double().norm().item(), "t_enh_norm": t_enh.detach().double().norm().item(),
        "ce": ce.item(), "focal": focal.item(), "unc_loss": unc_loss.item(), "proto": proto.item(), "loss": loss.item(),
        "grads": grads, "torch": torch.__version__,
    }
    torch.save(gold, os.path.join(OUT, f"{name}.pt"))
    print(f"{name}: loss={loss.item():.6f} ({len(rec.calls)} dropout calls)")


def run_supcon_case(name, B, D, C, seed, temperature):
    """SupConLoss of the reference (src/models/losses.py:67-88) on random embeddings: loss and d loss / d features."""
    lo = load_ref("losses")
    g = torch.Generator().manual_seed(seed)
    f = (torch.randn(B, D, generator=g) * 2.0).requires_grad_(True)
    labels = torch.randint(0, C, (B,), generator=g)
    labels[-1] = C            # one sample without any positive partner
    loss = lo.SupConLoss(temperature)(f, labels)
    loss.backward()
    torch.save({"config": dict(B=B, D=D, C=C, seed=seed, temperature=temperature), "labels": labels,
                "loss": loss.item(), "grad": f.grad.clone()}, os.path.join(OUT, f"{name}.pt"))
    print(f"{name}: loss={loss.item():.6f}")


def reference_feature_fusion(attr: str, hid: int):
    """Build `self.<attr>` exactly as the reference's source does: the nn.Sequential(...) constructor expression is cut
    out of src/models/audio_encoder.py / text_encoder.py and evaluated with the given `hid` (the encoder classes
    themselves cannot be imported: quality_gates needs librosa)."""
    fname = "text_encoder.py" if attr == "asr_fusion" else "audio_encoder.py"
    src = open(os.path.join(REF, "src", "models", fname)).read()
    start = src.index(f"self.{attr} = nn.Sequential(") + len(f"self.{attr} = ")
    depth, i = 0, start
    while True:
        depth += {"(": 1, ")": -1}.get(src[i], 0)
        i += 1
        if depth == 0 and src[i - 1] == ")":
            break
    return eval(src[start:i], {"nn": nn, "hid": hid})


def run_feature_fusion_case(name, attr, F, B, T, hid, seed):
    """The per-utterance fusion loop of AudioEncoder.forward (audio_encoder.py:114-138; text side text_encoder.py:68-73):
    features expanded over the frames, concatenated, pushed through the Sequential -- eval mode and train mode (recorded
    dropout masks), with gradients of sum(y * up)."""
    mod = reference_feature_fusion(attr, hid)
    w = synth.feature_fusion_weights(attr, F, seed=0, hidden=hid)
    mod.load_state_dict(w)
    g = torch.Generator().manual_seed(seed)
    seq = torch.randn(B, T, hid, generator=g)
    feats = torch.rand(B, F, generator=g) * 2.0 - 0.5
    up = torch.randn(B, T, hid, generator=g)
    out = {"config": dict(attr=attr, F=F, B=B, T=T, hid=hid, seed=seed)}
    for mode in ("eval", "train"):
        mod.train(mode == "train")
        mod.zero_grad()
        x = seq.clone().requires_grad_(True)
        with RecordedDropout(seed + 1) as rec:
            ys = []
            for i in range(B):                      # one utterance at a time, as the reference does
                base_seq = x[i]
                all_features = feats[i].unsqueeze(0).expand(base_seq.size(0), -1)
                fused_input = torch.cat([base_seq, all_features], dim=-1)
                ys.append(mod(fused_input))
            y = torch.stack(ys)
        (y * up).sum().backward()
        out[mode] = {"y": y.detach().clone(), "dx": x.grad.clone(), "dw": mod[0].weight.grad.clone(),
                     "db": mod[0].bias.grad.clone()}
        if mode == "train":
            assert len(rec.calls) == B
            out["train"]["mask"] = torch.stack([keep.to(torch.float32) / (1.0 - p) for _, p, keep in rec.calls])
    torch.save(out, os.path.join(OUT, f"{name}.pt"))
    print(f"{name}: |y_eval|={out['eval']['y'].abs().mean():.4f} kept={float((out['train']['mask'] > 0).float().mean()):.3f}")


def run_late_ood_case(name, B, C, D, seed):
    """LateStageOODDetector of the reference (src/models/dual_gate_ood.py:331-413) with fitted prototypes
    (update_prototypes on a synthetic labelled set), a non-default temperature and mix: the detector's own components'
    outputs plus the scalars of the LateOODResult."""
    path = os.path.join(REF, "src", "models", "dual_gate_ood.py")
    spec = importlib.util.spec_from_file_location("ref_dual_gate_ood", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(seed)
    det = mod.LateStageOODDetector(C, D)
    centers = torch.randn(C, D, generator=g)
    fit_labels = torch.randint(0, C, (40 * C,), generator=g)
    fit_feats = centers[fit_labels] + 0.5 * torch.randn(40 * C, D, generator=g)
    det.prototype_detector.update_prototypes(fit_feats, fit_labels)
    with torch.no_grad():
        det.energy_detector.temperature.fill_(1.7)
        det.combination_weights.copy_(torch.tensor([0.9, 0.2]))
    labels = torch.randint(0, C, (B,), generator=g)
    feats = centers[labels] + 0.5 * torch.randn(B, D, generator=g)
    feats[::5] += 0.3                           # a few samples off their class
    logits = torch.randn(B, C, generator=g) * 3.0
    with torch.no_grad():
        energy, _ = det.energy_detector(logits)
        dist, min_d = det.prototype_detector(feats)
        res = det(logits, feats)
    torch.save({"config": dict(B=B, C=C, D=D, seed=seed), "state": {k: v.clone() for k, v in det.state_dict().items()},
                "logits": logits, "features": feats, "energy": energy, "distances": dist, "min_distance": min_d,
                "result": dict(is_ood=res.is_ood, energy_score=res.energy_score, prototype_distance=res.prototype_distance,
                               combined_score=res.combined_score, confidence_score=res.confidence_score,
                               reason=res.reason.value)}, os.path.join(OUT, f"{name}.pt"))
    print(f"{name}: combined={res.combined_score:.5f} min_dist={res.prototype_distance:.4f} reason={res.reason.value}")


def run_eval_case(name, B, C, seed, views=5):
    """cfg5 semantics at small B: classifier eval path with fitted OpenMax + TTA mean + temperature + energy."""
    weights = synth.head_weights(C, 35, seed=0)
    m, _ = build_reference_head(C, 35, weights)
    clf = m["classifier"]
    g = torch.Generator().manual_seed(seed)
    val_fused = torch.randn(64, 512, generator=g)
    val_labels = torch.randint(0, C, (64,), generator=g)
    with torch.no_grad():
        # hand-unrolled feature extraction, src/train.py:221-236
        f = val_fused
        for layer in clf.deep_classifier.input_projection:
            f = layer(f)
        for blk, ln in zip(clf.deep_classifier.residual_layers, clf.deep_classifier.layer_norms):
            f = blk(ln(f))
        for i in range(4):
            f = clf.deep_classifier.output_projection[i](f)
        clf.fit_weibull(f, val_labels)
        fused_views = torch.randn(views, B, 512, generator=g)
        labels = torch.randint(0, C, (B,), generator=g)
        logits_views = torch.stack([clf(fused_views[v]) for v in range(views)])       # OpenMax on (eval.py:198)
        logits_plain = torch.stack([clf(fused_views[v], use_openmax=False) for v in range(views)])
        mean_logits = logits_views.mean(0)                                              # eval.py:186-190
        ev = load_eval_helpers()
        T = ev["find_optimal_temperature"](logits_plain[0], labels, "cpu")
        scaled = mean_logits / T
        probs = torch.softmax(scaled, dim=-1)
        preds = probs.argmax(dim=-1)
        energy = -torch.logsumexp(scaled, dim=-1)                                       # utils.py:12-14
    gold = {
        "config": dict(B=B, C=C, seed=seed, views=views),
        "weibull": {k: getattr(clf, k).clone() for k in synth.CLASSIFIER_BUFFERS},
        "val_features": f, "logits_views": logits_views, "logits_plain": logits_plain,
        "mean_logits": mean_logits, "temperature": float(T), "probs": probs, "preds": preds, "energy": energy,
    }
    torch.save(gold, os.path.join(OUT, f"{name}.pt"))
    print(f"{name}: T={T:.4f} preds={preds[:8].tolist()}")


def run_mixed_dims_case(name, B, Ta, Tt, audio_dim, text_dim, seed):
    """CrossModalAttention(audio_dim != text_dim) and FusionLayer(2 audio_dim != 2 text_dim) of the reference
    (src/models/cross_attention.py:7-53, fusion.py:6-25), eval mode, with key-padding masks: outputs and the gradients
    of a fixed linear objective with respect to inputs and every parameter."""
    ca = load_ref("cross_attention"); fu = load_ref("fusion")
    wc = synth.cross_weights(audio_dim=audio_dim, text_dim=text_dim)
    wf = synth.fusion_weights(audio_dim=2 * audio_dim, text_dim=2 * text_dim)
    cross = ca.CrossModalAttention(audio_dim, text_dim, shared_dim=256, num_heads=8).eval()
    fusion = fu.FusionLayer(2 * audio_dim, 2 * text_dim, 512).eval()
    assert not cross.load_state_dict(wc, strict=True).missing_keys
    assert not fusion.load_state_dict(wf, strict=True).missing_keys
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, Ta, audio_dim, generator=g)
    t = torch.randn(B, Tt, text_dim, generator=g)
    la = torch.randint((Ta + 1) // 2, Ta + 1, (B,), generator=g)
    lt = torch.randint(max(1, (Tt + 3) // 4), Tt + 1, (B,), generator=g)
    am = (torch.arange(Ta)[None] < la[:, None]).float()
    tm = (torch.arange(Tt)[None] < lt[:, None]).float()
    a = (a * am[..., None]).requires_grad_(True)
    t = (t * tm[..., None]).requires_grad_(True)
    av = torch.randn(B, 2 * audio_dim, generator=g).requires_grad_(True)
    tv = torch.randn(B, 2 * text_dim, generator=g).requires_grad_(True)
    ua, ut, up = torch.randn(B, Ta, audio_dim, generator=g), torch.randn(B, Tt, text_dim, generator=g), torch.randn(B, 512, generator=g)
    ea, et = cross(a, t, am, tm)
    fused = fusion(av, tv)
    ((ea * ua).sum() + (et * ut).sum() + (fused * up).sum()).backward()
    gold = {"config": dict(B=B, Ta=Ta, Tt=Tt, audio_dim=audio_dim, text_dim=text_dim, seed=seed),
            "inputs": {"a": a.detach().clone(), "t": t.detach().clone(), "a_mask": am, "t_mask": tm, "av": av.detach().clone(),
                       "tv": tv.detach().clone(), "ua": ua, "ut": ut, "up": up},
            "audio_enh": ea.detach().clone(), "text_enh": et.detach().clone(), "fused": fused.detach().clone(),
            "din": {"a": a.grad.clone(), "t": t.grad.clone(), "av": av.grad.clone(), "tv": tv.grad.clone()},
            "grads": {}}
    for grp, mod in (("cross", cross), ("fusion", fusion)):
        for n, p in mod.named_parameters():
            gold["grads"][f"{grp}/{n}"] = summarize(p.grad, f"{grp}/{n}")
    torch.save(gold, os.path.join(OUT, f"{name}.pt"))
    print(f"{name}: |audio_enh|={ea.norm().item():.4f} |text_enh|={et.norm().item():.4f} |fused|={fused.norm().item():.4f}")


def load_eval_helpers():
    """temperature_scaling + find_optimal_temperature from src/eval.py:43-67, extracted without importing the
    script's heavy dependencies."""
    src = open(os.path.join(REF, "src", "eval.py")).read()
    start = src.index("def temperature_scaling")
    end = src.index("\ndef main", start + 10)
    ns = {"torch": torch}
    exec(compile(src[start:end], "ref_eval_calibrate", "exec"), ns)
    return ns


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    # small-B versions of BASELINE.json configs 1-4 (same T / C structure, B cut so files stay small)
    run_train_case("train_cfg1_small", B=4, Ta=50, Tt=16, C=4, seed=1235)
    run_train_case("train_cfg2_shape", B=3, Ta=250, Tt=64, C=4, seed=1236)
    run_train_case("train_cfg3_c6", B=6, Ta=40, Tt=24, C=6, seed=1237)
    run_train_case("train_nomask", B=4, Ta=33, Tt=9, C=4, seed=1238, with_masks=False)
    run_train_case("train_cfg4_long", B=2, Ta=300, Tt=96, C=4, seed=1239)
    run_eval_case("eval_cfg5_small", B=32, C=6, seed=1240)
    run_train_dropout_case("dropout_train_small", B=4, Ta=40, Tt=17, C=4, seed=1241)
    run_supcon_case("supcon_small", B=48, D=512, C=4, seed=1242, temperature=0.07)
    run_feature_fusion_case("feature_fusion_combined", "combined_fusion", F=20, B=3, T=13, hid=128, seed=1243)
    run_late_ood_case("late_ood_small", B=37, C=6, D=64, seed=1245)
    run_feature_fusion_case("feature_fusion_asr", "asr_fusion", F=8, B=2, T=7, hid=128, seed=1244)
    run_mixed_dims_case("mixed_dims_small", B=3, Ta=21, Tt=9, audio_dim=768, text_dim=1024, seed=1246)
