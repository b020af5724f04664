"""Per-module and whole-head parity report of the CUDA path against the CPU oracle (run under gpurun).

Prints one line per tensor (max relative error, normalised by the reference's max-abs) for forward outputs and
every parameter gradient; never stops at the first failure.  Usage: python tools/gpu_parity_report.py [f32|bf16|all]
"""
import sys
import traceback

import torch

sys.path.insert(0, ".")
import mmser_b200  # noqa: E402
from mmser_b200 import models as M  # noqa: E402
from oracle import fusion_head_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402

dev = torch.device("cuda:0")
WORST = {}


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    if torch.isnan(b).any():
        return 0.0 if torch.equal(torch.isnan(a), torch.isnan(b)) else float("inf")
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


def report(tag, name, got, ref, tol):
    if got is None and ref is None:
        return
    if got is None or ref is None:
        print(f"  [{tag}] {name:48s} MISSING got={got is not None} ref={ref is not None}")
        WORST[tag] = float("inf")
        return
    if ref.abs().max().item() < 1e-7 and got.detach().abs().max().item() < 1e-5:
        r = 0.0
    else:
        r = rel(got, ref)
    WORST[tag] = max(WORST.get(tag, 0.0), r)
    flag = "ok  " if r <= tol else "FAIL"
    if r > tol or "-v" in sys.argv:
        print(f"  [{tag}] {flag} {name:48s} rel={r:.3e}")


def leafify(w):
    out = {}
    for k, v in w.items():
        v = v.clone()
        if v.is_floating_point() and k not in synth.CLASSIFIER_BUFFERS:
            v.requires_grad_(True)
        out[k] = v
    return out


def compare_grads(tag, module, wref, tol):
    for n, p in module.named_parameters():
        ref = wref[n].grad
        got = p.grad
        if ref is None and (got is None or got.abs().max().item() == 0):
            continue
        if ref is None:
            ref = torch.zeros_like(wref[n])
        report(tag, "grad " + n, got, ref, tol)


def run(tag, fn):
    try:
        fn()
        print(f"[{tag}] worst rel err = {WORST.get(tag, 0.0):.3e}")
    except Exception:  # noqa: BLE001
        print(f"[{tag}] EXCEPTION")
        traceback.print_exc()
        WORST[tag] = float("inf")


def tier(dtype):
    tol_f, tol_g = (1e-4, 1e-4) if dtype == torch.float32 else (2e-2, 2e-2)
    sfx = "f32" if dtype == torch.float32 else "bf16"
    cast = lambda t: t.to(dev).to(dtype)  # noqa: E731

    def t_adapter():
        tag = f"adapter/{sfx}"
        w = synth.adapter_weights("adapter_a")
        m = M.BottleneckAdapter().to(dev); m.load_state_dict(w)
        g = torch.Generator().manual_seed(1)
        x = torch.randn(3, 37, 768, generator=g)
        xr = x.clone().requires_grad_(True)
        wr = leafify(w)
        yr = O.adapter(xr, wr)
        up = torch.randn(yr.shape, generator=g)
        (yr * up).sum().backward()
        xg = cast(x).requires_grad_(True)
        y = m.residual_forward(xg)
        (y.float() * up.to(dev)).sum().backward()
        report(tag, "y", y, yr, tol_f)
        report(tag, "dx", xg.grad, xr.grad, tol_g)
        compare_grads(tag, m, wr, tol_g)

    def t_cross(masks=True):
        tag = f"cross{'_mask' if masks else '_nomask'}/{sfx}"
        w = synth.cross_weights()
        m = M.CrossModalAttention(768, 768, dropout=0.0).to(dev); m.load_state_dict(w)
        a, t, am, tm, _ = synth.make_inputs(3, 70, 19, 4, seed=7, with_masks=masks)
        ar, tr = a.clone().requires_grad_(True), t.clone().requires_grad_(True)
        wr = leafify(w)
        ea, et = O.cross_attention(ar, tr, am, tm, wr)
        g = torch.Generator().manual_seed(2)
        ua, ut = torch.randn(ea.shape, generator=g), torch.randn(et.shape, generator=g)
        ((ea * ua).sum() + (et * ut).sum()).backward()
        ag, tg = cast(a).requires_grad_(True), cast(t).requires_grad_(True)
        oa, ot = m(ag, tg, am.to(dev) if masks else None, tm.to(dev) if masks else None)
        ((oa.float() * ua.to(dev)).sum() + (ot.float() * ut.to(dev)).sum()).backward()
        report(tag, "audio_enh", oa, ea, tol_f)
        report(tag, "text_enh", ot, et, tol_f)
        report(tag, "d audio", ag.grad, ar.grad, tol_g)
        report(tag, "d text", tg.grad, tr.grad, tol_g)
        compare_grads(tag, m, wr, tol_g)

    def t_pool():
        tag = f"pool/{sfx}"
        w = synth.pool_weights("pool_a")
        m = M.AttentiveStatsPooling(768).to(dev); m.load_state_dict(w)
        a, _, am, _, _ = synth.make_inputs(4, 53, 8, 4, seed=9)
        xr = a.clone().requires_grad_(True)
        wr = leafify(w)
        yr = O.attentive_stats_pooling(xr, am, wr)
        g = torch.Generator().manual_seed(3)
        up = torch.randn(yr.shape, generator=g)
        (yr * up).sum().backward()
        xg = cast(a).requires_grad_(True)
        y = m(xg, am.to(dev))
        (y.float() * up.to(dev)).sum().backward()
        report(tag, "pooled", y, yr, tol_f)
        report(tag, "dx", xg.grad, xr.grad, tol_g)
        compare_grads(tag, m, wr, tol_g)

    def t_fusion():
        tag = f"fusion/{sfx}"
        w = synth.fusion_weights()
        m = M.FusionLayer(1536, 1536, 512).to(dev).eval(); m.load_state_dict(w)
        g = torch.Generator().manual_seed(4)
        av, tv = torch.randn(9, 1536, generator=g), torch.randn(9, 1536, generator=g)
        avr, tvr = av.clone().requires_grad_(True), tv.clone().requires_grad_(True)
        wr = leafify(w)
        yr = O.fusion(avr, tvr, wr)
        up = torch.randn(yr.shape, generator=g)
        (yr * up).sum().backward()
        avg, tvg = cast(av).requires_grad_(True), cast(tv).requires_grad_(True)
        y = m(avg, tvg)
        (y.float() * up.to(dev)).sum().backward()
        report(tag, "fused", y, yr, tol_f)
        report(tag, "d av", avg.grad, avr.grad, tol_g)
        report(tag, "d tv", tvg.grad, tvr.grad, tol_g)
        compare_grads(tag, m, wr, tol_g)

    def t_clf():
        tag = f"classifier/{sfx}"
        C = 4
        w = synth.classifier_weights(C, 35)
        m = M.AdvancedOpenMaxClassifier(512, C, dropout=0.0).to(dev); m.load_state_dict(w)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(10, 512, generator=g)
        xr = x.clone().requires_grad_(True)
        wr = leafify(w)
        lr, ur, alr = O.classifier(xr, wr, 35, use_openmax=False, training=True, return_uncertainty=True)
        ul, uu = torch.randn(lr.shape, generator=g), torch.randn(ur.shape, generator=g)
        ((lr * ul).sum() + (ur * uu).sum()).backward()
        xg = cast(x).requires_grad_(True)
        m.train()
        lg, un, al = m(xg, use_openmax=False, return_uncertainty=True)
        ((lg * ul.to(dev)).sum() + (un * uu.to(dev)).sum()).backward()
        report(tag, "logits", lg, lr, tol_f)
        report(tag, "unc", un, ur, tol_f)
        report(tag, "features", m.last_features, O.classifier_features(x, w, 35), tol_f)
        report(tag, "dx", xg.grad, xr.grad, tol_g)
        compare_grads(tag, m, wr, tol_g)

    def t_loss():
        tag = f"loss/{sfx}"
        C = 6
        g = torch.Generator().manual_seed(6)
        B = 37
        logits = torch.randn(B, C, generator=g) * 5.0          # some beyond the +-10 clamp
        unc = torch.rand(B, 1, generator=g)
        emb = torch.randn(B, 512, generator=g) * 4.0            # some beyond the +-10 clamp
        protos = torch.randn(C, 512, generator=g) * 0.5
        labels = torch.randint(0, C, (B,), generator=g)
        lr, ur, er, pr = [t.clone().requires_grad_(True) for t in (logits, unc, emb, protos)]
        ref = O.train_loss(lr, ur, torch.zeros(()), er, labels, pr, C)
        ref["loss"].backward()
        lgv, ung, emg, prg = logits.to(dev).requires_grad_(True), unc.to(dev).requires_grad_(True), \
            cast(emb).requires_grad_(True), protos.to(dev).requires_grad_(True)
        terms = mmser_b200.functional.HeadLossFn.apply(lgv, ung, emg, prg, labels.to(dev),
                                                       dict(w_ce=1.0, w_focal=0.3, w_unc=0.05, w_proto=0.01))
        terms[4].backward()
        tl = tol_f if dtype == torch.float32 else 2e-2
        for i, k in enumerate(("ce", "focal", "unc_loss", "proto", "loss")):
            report(tag, k, terms[i], ref[k], tl)
        report(tag, "dlogits", lgv.grad, lr.grad, tol_g)
        report(tag, "dunc", ung.grad, ur.grad, tol_g)
        report(tag, "demb", emg.grad, er.grad, tol_g)
        report(tag, "dprotos", prg.grad, pr.grad, tol_g)
        # individual loss modules
        ce_m = M.LabelSmoothingCrossEntropy(0.1)(logits.to(dev), labels.to(dev))
        report(tag, "module ce", ce_m, O.label_smoothing_ce(logits, labels, 0.1), tl)
        fo_m = M.ClassBalancedFocalLoss(num_classes=C).to(dev)(logits.to(dev), labels.to(dev))
        report(tag, "module focal", fo_m, O.class_balanced_focal(logits, labels, num_classes=C), tl)
        pm = M.PrototypeMemory(C, 512).to(dev)
        pm.load_state_dict({"prototypes": protos})
        report(tag, "module proto", pm.prototype_loss(cast(emb), labels.to(dev)), O.prototype_loss(emb, labels, protos), tl)

    def t_head(case):
        B, Ta, Tt, C, masks = case
        tag = f"head B{B} Ta{Ta} Tt{Tt} C{C}{'' if masks else ' nomask'}/{sfx}"
        w = synth.head_weights(C, 35)
        head = mmser_b200.FusionHead(C).to(dev); head.load_group_state(w)
        a, t, am, tm, labels = synth.make_inputs(B, Ta, Tt, C, seed=1234, with_masks=masks)
        wr = {k: leafify(v) for k, v in w.items()}
        ref = O.head_forward(a, t, am, tm, labels, wr, C, 35)
        ref["loss"].backward()
        out = head(cast(a), cast(t), am.to(dev) if masks else None, tm.to(dev) if masks else None, labels.to(dev))
        out["loss"].backward()
        for k in ("a_enh", "t_enh", "a_vec", "t_vec", "fused", "logits", "unc", "ce", "focal", "unc_loss", "proto", "loss"):
            report(tag, k, out[k], ref[k], tol_f)
        same = torch.equal(out["logits"].argmax(1).cpu(), ref["logits"].argmax(1))
        print(f"  [{tag}] argmax identical: {same}")
        if not same:
            WORST[tag] = float("inf")
        for grp in head.GROUPS:
            compare_grads(tag, getattr(head, grp), {n: wr[grp][n] for n in wr[grp]}, tol_g)

    run(f"adapter/{sfx}", t_adapter)
    run(f"cross_mask/{sfx}", lambda: t_cross(True))
    run(f"cross_nomask/{sfx}", lambda: t_cross(False))
    run(f"pool/{sfx}", t_pool)
    run(f"fusion/{sfx}", t_fusion)
    run(f"classifier/{sfx}", t_clf)
    run(f"loss/{sfx}", t_loss)
    for case in ((4, 50, 16, 4, True), (5, 33, 9, 6, False)):
        B, Ta, Tt, C, masks = case
        run(f"head B{B} Ta{Ta} Tt{Tt} C{C}{'' if masks else ' nomask'}/{sfx}", lambda c=case: t_head(c))


which = [a for a in sys.argv[1:] if not a.startswith("-")]
which = which[0] if which else "all"
torch.set_num_threads(16)
if which in ("f32", "all"):
    tier(torch.float32)
if which in ("bf16", "all"):
    tier(torch.bfloat16)
print("==== SUMMARY ====")
for k, v in WORST.items():
    print(f"{k:50s} {v:.3e}")
