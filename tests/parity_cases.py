"""Parity cases shared by tests/test_gpu_parity.py and tools/gpu_parity_report.py.

Methodology (DESIGN.md "Parity"):
  E    = oracle in float64                       -- the exact answer
  R32  = oracle in float32                       -- the reference's own arithmetic (what 1e-4 is measured against)
  Rbf  = oracle in float32 with every Linear's operands rounded to bf16 (what torch.autocast(bf16) does)
  G    = the CUDA path (fp32 tier or bf16 tier)
A tensor passes when  err(G, E) <= max(tol, NOISE_K * err(R, E))  with R = R32 (fp32 tier) or Rbf (bf16 tier):
tol is north_star's 1e-4 / 2e-2, and the second term is the noise floor of the reference arithmetic itself --
ReLU-mask / clamp / argmax flips of near-zero pre-activations perturb gradients by O(1/B) in ANY implementation,
fp32 CPU-vs-GPU included, so a fixed tolerance alone would be a coin toss on small batches.
Metric: max-abs error / max-abs reference (fp32 tier); Frobenius error / Frobenius reference (bf16 tier).
Tensors that are mathematically zero (attention key biases, softmax-shift biases) are floored at 1% of the
module's largest gradient.  Bias gradients with fewer than 16 elements (the [1] gate biases) take their noise floor
from the weight of the same layer: one element is one draw of the noise, not an estimate of its scale.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from oracle import fusion_head_oracle as O
from oracle import synth

NOISE_K = 3.0
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


OUTLIER_FRAC = 0.002      # gradients only: a ReLU-mask flip perturbs a handful of entries (one row of dW, one of db)


def _err(got: torch.Tensor, ref: torch.Tensor, frob: bool, floor: float = 0.0, outliers: bool = False) -> float:
    """fp32 tier: max-abs error / max-abs reference; with `outliers` the largest max(2, 0.2%) entries are set aside
    (they must still be sane: the Frobenius error of the whole tensor is bounded separately by the caller).
    bf16 tier: Frobenius error / Frobenius reference."""
    a, b = got.detach().double().cpu().reshape(-1), ref.detach().double().cpu().reshape(-1)
    if torch.isnan(b).any() or torch.isnan(a).any():
        return 0.0 if torch.equal(torch.isnan(a), torch.isnan(b)) and torch.allclose(a[~torch.isnan(a)], b[~torch.isnan(b)], rtol=1e-2, atol=1e-3) else float("inf")
    if frob:
        return (a - b).norm().item() / max(b.norm().item(), floor * (b.numel() ** 0.5), 1e-30)
    d = (a - b).abs()
    if outliers and d.numel() > 8:
        k = max(2, int(OUTLIER_FRAC * d.numel()))
        d = torch.topk(d, k + 1, largest=True).values[-1:]      # (k+1)-th largest
    return d.max().item() / max(b.abs().max().item(), floor, 1e-30)


def _leaf(w: Dict[str, torch.Tensor], dtype) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in w.items():
        v = v.detach().clone()
        if v.is_floating_point():
            v = v.to(dtype)
            if k not in synth.CLASSIFIER_BUFFERS:
                v.requires_grad_(True)
        out[k] = v
    return out


class Case:
    """One parity problem.  `oracle(inputs, weights)` -> (dict of output tensors, scalar objective);
    `cuda(inputs, dtype, device)` -> (dict of outputs, dict 'group/param' -> grad, dict input-name -> grad)."""

    def __init__(self, name: str, inputs: Dict[str, torch.Tensor], weights: Dict[str, Dict[str, torch.Tensor]],
                 oracle: Callable, cuda: Callable, grad_inputs=(), prepare: Optional[Callable] = None):
        self.name, self.inputs, self.weights, self.oracle, self.cuda, self.grad_inputs = \
            name, inputs, weights, oracle, cuda, tuple(grad_inputs)
        # prepare(device) -> extra inputs that need the GPU library (the dropout masks the kernels will apply)
        self.prepare = prepare

    def run_oracle(self, dtype, rounding=None):
        ins = {}
        for k, v in self.inputs.items():
            if torch.is_tensor(v) and v.is_floating_point():
                v = v.detach().clone().to(dtype)
                if k in self.grad_inputs:
                    v.requires_grad_(True)
            ins[k] = v
        ws = {g: _leaf(w, dtype) for g, w in self.weights.items()}
        ctx = O.operand_rounding(rounding) if rounding is not None else _Null()
        with ctx:
            outs, obj = self.oracle(ins, ws)
            obj.backward()
        grads = {f"{g}/{n}": (t.grad if t.grad is not None else torch.zeros_like(t))
                 for g, w in ws.items() for n, t in w.items() if t.requires_grad}
        igrads = {k: ins[k].grad for k in self.grad_inputs}
        return {k: v.detach() for k, v in outs.items()}, grads, igrads

    def check(self, dtype, device, verbose=False):
        """Returns (failures: list[str], worst: float)."""
        frob = dtype != torch.float32
        tol = TOL[dtype]
        # the bf16 tier feeds bf16-rounded activations; hand the oracle the same rounded inputs
        saved = self.inputs
        if self.prepare is not None:
            self.inputs = dict(saved, **self.prepare(device))
        if dtype != torch.float32:
            self.inputs = {k: (v.to(dtype).float() if (torch.is_tensor(v) and v.is_floating_point() and k not in ("a_mask", "t_mask", "mask"))
                               else v) for k, v in self.inputs.items()}
        try:
            e_out, e_g, e_ig = self.run_oracle(torch.float64)
            r_out, r_g, r_ig = self.run_oracle(torch.float32, None if dtype == torch.float32 else dtype)
            g_out, g_g, g_ig = self.cuda(self.inputs, dtype, device)
        finally:
            self.inputs = saved
        fails, worst = [], 0.0
        # every tensor's raw numbers, no noise term and no outlier set-aside: (kind, name, err(G,E), err(R,E), err(G,R),
        # elements) in the tier's metric -- what tools/gpu_parity_report.py prints against the PLAIN tolerance
        self.records = []

        def one(kind, name, got, ref_exact, ref_noise, floor=0.0, companion=None):
            nonlocal worst
            if got is None:
                fails.append(f"{kind} {name}: missing")
                return
            is_grad = kind != "out"
            self.records.append((kind, name, _err(got, ref_exact, frob, floor), _err(ref_noise, ref_exact, frob, floor),
                                 _err(got, ref_noise, frob, floor), int(ref_exact.numel())))
            e = _err(got, ref_exact, frob, floor, outliers=is_grad)
            n = _err(ref_noise, ref_exact, frob, floor, outliers=is_grad)
            if companion is not None:
                # a tensor with a handful of elements (a [1] gate bias) yields ONE draw of the reference-arithmetic
                # noise, not its scale; take the scale from the weight of the same layer (thousands of draws)
                n = max(n, _err(companion[1], companion[0], frob, floor, outliers=is_grad))
            lim = max(tol, NOISE_K * n)
            if is_grad and not frob:
                # the set-aside outliers must be flip-sized, not garbage: bound the whole tensor in Frobenius norm
                ef = _err(got, ref_exact, True, floor)
                if ef > 50 * lim:
                    e = max(e, ef / 50)
            worst = max(worst, e / lim)
            if verbose or e > lim:
                print(f"    {'FAIL' if e > lim else 'ok  '} {kind:5s} {name:52s} err={e:.2e} noise={n:.2e} limit={lim:.2e}")
            if e > lim:
                fails.append(f"{kind} {name}: err {e:.3e} > limit {lim:.3e} (reference-arithmetic noise {n:.3e})")

        for k in e_out:
            if k in g_out:
                one("out", k, g_out[k], e_out[k], r_out[k])
        for k in e_ig:
            one("din", k, g_ig.get(k), e_ig[k], r_ig[k])
        groups = sorted({k.split("/")[0] for k in e_g})
        for grp in groups:
            keys = [k for k in e_g if k.startswith(grp + "/")]
            scale = max(e_g[k].abs().max().item() for k in keys)
            for k in keys:
                if k.endswith("anchor_clustering.temperature"):
                    if g_g.get(k) is not None and float(g_g[k].abs().max()) != 0.0:
                        fails.append(f"grad {k}: must be None / zero")
                    continue
                comp = None
                if e_g[k].numel() < 16 and k.endswith(".bias"):
                    kw = k[:-len(".bias")] + ".weight"
                    if kw in e_g:
                        comp = (e_g[kw], r_g[kw])
                one("grad", k, g_g.get(k), e_g[k], r_g[k], floor=1e-2 * scale, companion=comp)
        return fails, worst

    def group_errors(self, dtype, device):
        """Raw parity numbers per parameter group at any size, NO noise term, NO outlier set-aside, NO floors:
        {group: dict(ge, re, gr, n)} with  ge = ||G - E|| / ||E||  over all gradient tensors of the group concatenated
        (E = fp64 oracle, R = the oracle in the tier's reference arithmetic, G = CUDA), plus 'out/<name>' entries for the
        forward quantities.  Used by the BASELINE-size parity tests."""
        saved = self.inputs
        if self.prepare is not None:
            self.inputs = dict(saved, **self.prepare(device))
        if dtype != torch.float32:
            self.inputs = {k: (v.to(dtype).float() if (torch.is_tensor(v) and v.is_floating_point() and k not in ("a_mask", "t_mask", "mask"))
                               else v) for k, v in self.inputs.items()}
        try:
            e_out, e_g, e_ig = self.run_oracle(torch.float64)
            r_out, r_g, r_ig = self.run_oracle(torch.float32, None if dtype == torch.float32 else dtype)
            g_out, g_g, g_ig = self.cuda(self.inputs, dtype, device)
        finally:
            self.inputs = saved

        def rel(x, y):
            return (x - y).norm().item() / max(y.norm().item(), 1e-30)

        def cat(d, keys):
            return torch.cat([d[k].detach().double().cpu().reshape(-1) for k in keys])

        res = {}
        for k in e_out:
            if k in g_out:
                E, R, G = (t[k].detach().double().cpu().reshape(-1) for t in (e_out, r_out, g_out))
                res[f"out/{k}"] = dict(ge=rel(G, E), re=rel(R, E), gr=rel(G, R), n=int(E.numel()))
        for k in e_ig:
            E, R, G = (t[k].detach().double().cpu().reshape(-1) for t in (e_ig, r_ig, g_ig))
            res[f"din/{k}"] = dict(ge=rel(G, E), re=rel(R, E), gr=rel(G, R), n=int(E.numel()))
        for grp in sorted({k.split("/")[0] for k in e_g}):
            keys = [k for k in e_g if k.startswith(grp + "/") and not k.endswith("anchor_clustering.temperature")
                    and g_g.get(k) is not None]
            if not keys:
                continue
            E, R, G = cat(e_g, keys), cat(r_g, keys), cat(g_g, keys)
            res[f"grad/{grp}"] = dict(ge=rel(G, E), re=rel(R, E), gr=rel(G, R), n=int(E.numel()))
        return res


class _Null:
    def __enter__(self): return None
    def __exit__(self, *a): return False


# ---------------------------------------------------------------------------------------------------
def _param_grads(group: str, module) -> Dict[str, torch.Tensor]:
    return {f"{group}/{n}": p.grad for n, p in module.named_parameters()}


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


# ---------------------------------------------------------------------------------------------------
# dropout: the kernels' masks are a pure function of (seed, site, row, col) (csrc/dropout.cuh).  The tests pin the seed
# of each module, export the masks through the C-ABI (ser_dropout_mask) and run the oracle with exactly those masks.
# (Seed choice: a mask realisation can put one of the ~10^5 ReLU inputs of a case within fp32 rounding of zero; the CUDA
#  fp32 tier and the fp64 oracle then disagree on that unit's gate and every gradient upstream of it moves by ~3e-3 --
#  a discrete event the noise floor cannot see.  Observed once while experimenting with another hash: 1 of 6
#  consecutive seeds did that for head_dropout/fp32, the others passed with >= 5x margin.)
DROP_SEEDS = {"feat": 0x5DEECE66D1234567, "cross": 0x1234567887654321, "fusion": 0x0BADC0FFEE123457, "classifier": 0x7EDCBA9876543210}


def _seed_tensor(name, dev):
    return torch.tensor([DROP_SEEDS[name]], dtype=torch.int64, device=dev)


def _pin_seed(module, name, dev):
    module._drop_seed.counter = _seed_tensor(name, dev)      # the next forward uses exactly this value


def cross_masks(dev, p, B, H, Ta, Tt, D, Dt=None):
    from mmser_b200.functional import dropout_mask as dm
    s = _seed_tensor("cross", dev)
    return {"cross.prob_a": dm(s, 1, p, B * H * Ta, Tt).view(B, H, Ta, Tt).cpu(),
            "cross.prob_t": dm(s, 2, p, B * H * Tt, Ta).view(B, H, Tt, Ta).cpu(),
            "cross.res_a": dm(s, 3, p, B * Ta, D).view(B, Ta, D).cpu(),
            "cross.res_t": dm(s, 4, p, B * Tt, Dt or D).view(B, Tt, Dt or D).cpu()}


def fusion_masks(dev, p, B, P=512):
    from mmser_b200.functional import dropout_mask as dm
    s = _seed_tensor("fusion", dev)
    return {"fusion.a": dm(s, 5, p, B, P).cpu(), "fusion.t": dm(s, 6, p, B, P).cpu()}


def classifier_masks(dev, p, B, L, P=512, F=256, U=64):
    from mmser_b200.functional import dropout_mask as dm
    s = _seed_tensor("classifier", dev)
    out = {"clf.in": dm(s, 7, p, B, P).cpu(), "clf.out": dm(s, 8, p, B, F).cpu(), "clf.unc": dm(s, 9, p, B, U).cpu()}
    for i in range(L):
        out[f"clf.block{i}.hidden"] = dm(s, 16 + 2 * i, p, B, P).cpu()
        out[f"clf.block{i}.out"] = dm(s, 16 + 2 * i + 1, p, B, P).cpu()
    return out


def adapter_case(B=3, T=37):
    import mmser_b200
    from mmser_b200 import models as M
    w = {"adapter": synth.adapter_weights("adapter_a")}
    ins = {"x": _rand((B, T, 768), 1), "up": _rand((B, T, 768), 2)}

    def oracle(i, ws):
        y = O.adapter(i["x"], ws["adapter"])
        return {"y": y}, (y * i["up"]).sum()

    def cuda(i, dtype, dev):
        m = M.BottleneckAdapter().to(dev); m.load_state_dict(w["adapter"])
        x = i["x"].to(dev).to(dtype).requires_grad_(True)
        y = m.residual_forward(x)
        (y.float() * i["up"].to(dev)).sum().backward()
        return {"y": y}, _param_grads("adapter", m), {"x": x.grad}

    return Case("adapter", ins, w, oracle, cuda, grad_inputs=("x",))


def feature_fusion_case(B=3, T=37, F=20, p_drop=0.0):
    """Linear(768 + F -> 768) . ReLU . Dropout on [frames ; utterance features]."""
    from mmser_b200 import models as M
    w = {"featfuse": synth.feature_fusion_weights("combined_fusion", F)}
    g = torch.Generator().manual_seed(31)
    ins = {"x": _rand((B, T, 768), 21), "feats": torch.rand(B, F, generator=g) * 2.0 - 0.5, "up": _rand((B, T, 768), 22)}

    def oracle(i, ws):
        with O.dropout_masks(i.get("_masks")):
            y = O.utterance_feature_fusion(i["x"], i["feats"], ws["featfuse"])
        return {"y": y}, (y * i["up"]).sum()

    def cuda(i, dtype, dev):
        m = M.UtteranceFeatureFusion(768, F, dropout=p_drop).to(dev); m.load_state_dict(w["featfuse"])
        m.train()
        _pin_seed(m, "feat", dev)
        x = i["x"].to(dev).to(dtype).requires_grad_(True)
        y = m(x, i["feats"].to(dev))
        (y.float() * i["up"].to(dev)).sum().backward()
        return {"y": y}, _param_grads("featfuse", m), {"x": x.grad}

    def prep(dev):
        from mmser_b200.functional import dropout_mask as dm
        return {"_masks": {"feat.out": dm(_seed_tensor("feat", dev), 10, p_drop, B * T, 768).view(B, T, 768).cpu()}}

    return Case("feature_fusion_dropout" if p_drop > 0 else "feature_fusion", ins, w, oracle, cuda, grad_inputs=("x",),
                prepare=prep if p_drop > 0 else None)


def cross_case(B=3, Ta=70, Tt=19, masks=True, seed=7, p_drop=0.0, text_dim=768):
    """text_dim != 768: CrossModalAttention(audio_dim != text_dim) (cross_attention.py:7-30), the library's unfolded path."""
    from mmser_b200 import models as M
    w = {"cross": synth.cross_weights(text_dim=text_dim)}
    a, t, am, tm, _ = synth.make_inputs(B, Ta, Tt, 4, seed=seed, with_masks=masks)
    if text_dim != 768:
        t = _rand((B, Tt, text_dim), 5) * (t.abs().sum(-1, keepdim=True) > 0)      # same padding pattern (zero-filled)
    ins = {"a": a, "t": t, "a_mask": am, "t_mask": tm, "ua": _rand((B, Ta, 768), 3), "ut": _rand((B, Tt, text_dim), 4)}

    def oracle(i, ws):
        am_ = None if i["a_mask"] is None else i["a_mask"].to(i["a"].dtype)
        tm_ = None if i["t_mask"] is None else i["t_mask"].to(i["a"].dtype)
        with O.dropout_masks(i.get("_masks")):
            ea, et = O.cross_attention(i["a"], i["t"], am_, tm_, ws["cross"])
        return {"audio_enh": ea, "text_enh": et}, (ea * i["ua"]).sum() + (et * i["ut"]).sum()

    def cuda(i, dtype, dev):
        m = M.CrossModalAttention(768, text_dim, dropout=p_drop).to(dev); m.load_state_dict(w["cross"])
        m.train()
        _pin_seed(m, "cross", dev)
        ag = i["a"].to(dev).to(dtype).requires_grad_(True)
        tg = i["t"].to(dev).to(dtype).requires_grad_(True)
        ea, et = m(ag, tg, None if i["a_mask"] is None else i["a_mask"].to(dev),
                   None if i["t_mask"] is None else i["t_mask"].to(dev))
        ((ea.float() * i["ua"].to(dev)).sum() + (et.float() * i["ut"].to(dev)).sum()).backward()
        return {"audio_enh": ea, "text_enh": et}, _param_grads("cross", m), {"a": ag.grad, "t": tg.grad}

    prep = (lambda dev: {"_masks": cross_masks(dev, p_drop, B, 8, Ta, Tt, 768, text_dim)}) if p_drop > 0 else None
    return Case(f"cross{'_masked' if masks else '_nomask'}{'_dropout' if p_drop > 0 else ''}", ins, w, oracle, cuda,
                grad_inputs=("a", "t"), prepare=prep)


def pool_case(B=4, T=53, masks=True):
    from mmser_b200 import models as M
    w = {"pool": synth.pool_weights("pool_a")}
    a, _, am, _, _ = synth.make_inputs(B, T, 8, 4, seed=9, with_masks=masks)
    ins = {"x": a, "mask": am, "up": _rand((B, 1536), 5)}

    def oracle(i, ws):
        m_ = None if i["mask"] is None else i["mask"].to(i["x"].dtype)
        y = O.attentive_stats_pooling(i["x"], m_, ws["pool"])
        return {"pooled": y}, (y * i["up"]).sum()

    def cuda(i, dtype, dev):
        m = M.AttentiveStatsPooling(768).to(dev); m.load_state_dict(w["pool"])
        x = i["x"].to(dev).to(dtype).requires_grad_(True)
        y = m(x, None if i["mask"] is None else i["mask"].to(dev))
        (y.float() * i["up"].to(dev)).sum().backward()
        return {"pooled": y}, _param_grads("pool", m), {"x": x.grad}

    return Case("pool", ins, w, oracle, cuda, grad_inputs=("x",))


def fusion_case(B=9, p_drop=0.0, text_dim=1536):
    """text_dim != 1536: FusionLayer(audio_dim != text_dim) (fusion.py:6-16)."""
    from mmser_b200 import models as M
    w = {"fusion": synth.fusion_weights(text_dim=text_dim)}
    # (bf16 tier: a gate-MLP ReLU input within bf16 rounding of zero is one discrete flip = 1/sqrt(B * 128) ~ 3e-2 on that
    #  layer's gradient for whichever implementation draws it -- DESIGN.md section 4; the seeds are fixed, not tuned per run)
    ins = {"av": _rand((B, 1536), 6), "tv": _rand((B, text_dim), 7 if text_dim == 1536 else 27), "up": _rand((B, 512), 8)}

    def oracle(i, ws):
        with O.dropout_masks(i.get("_masks")):
            y = O.fusion(i["av"], i["tv"], ws["fusion"])
        return {"fused": y}, (y * i["up"]).sum()

    def cuda(i, dtype, dev):
        m = M.FusionLayer(1536, text_dim, 512).to(dev); m.load_state_dict(w["fusion"])
        m.proj_a[2].p = m.proj_t[2].p = p_drop
        m.train()
        _pin_seed(m, "fusion", dev)
        av = i["av"].to(dev).to(dtype).requires_grad_(True)
        tv = i["tv"].to(dev).to(dtype).requires_grad_(True)
        y = m(av, tv)
        (y.float() * i["up"].to(dev)).sum().backward()
        return {"fused": y}, _param_grads("fusion", m), {"av": av.grad, "tv": tv.grad}

    prep = (lambda dev: {"_masks": fusion_masks(dev, p_drop, B)}) if p_drop > 0 else None
    return Case("fusion_dropout" if p_drop > 0 else "fusion", ins, w, oracle, cuda, grad_inputs=("av", "tv"), prepare=prep)


def classifier_case(B=64, C=4, L=35, p_drop=0.0, input_dim=512):
    from mmser_b200 import models as M
    w = {"classifier": synth.classifier_weights(C, L, input_dim=input_dim)}
    ins = {"x": _rand((B, input_dim), 10), "ul": _rand((B, C), 11), "uu": _rand((B, 1), 12)}

    def oracle(i, ws):
        with O.dropout_masks(i.get("_masks")):
            lg, un, _ = O.classifier(i["x"], ws["classifier"], L, use_openmax=False, training=True, return_uncertainty=True)
            f = O.classifier_features(i["x"], ws["classifier"], L)
        return {"logits": lg, "unc": un, "features": f}, (lg * i["ul"]).sum() + (un * i["uu"]).sum()

    def cuda(i, dtype, dev):
        m = M.AdvancedOpenMaxClassifier(input_dim, C, num_layers=L, dropout=p_drop).to(dev); m.load_state_dict(w["classifier"])
        m.train()
        _pin_seed(m, "classifier", dev)
        x = i["x"].to(dev).to(dtype).requires_grad_(True)
        lg, un, al = m(x, use_openmax=False, return_uncertainty=True)
        assert float(al) == 0.0
        ((lg * i["ul"].to(dev)).sum() + (un * i["uu"].to(dev)).sum()).backward()
        return {"logits": lg, "unc": un, "features": m.last_features}, _param_grads("classifier", m), {"x": x.grad}

    prep = (lambda dev: {"_masks": classifier_masks(dev, p_drop, B, L)}) if p_drop > 0 else None
    return Case(("classifier" if B == 64 else f"classifier_b{B}") + ("_dropout" if p_drop > 0 else "") +
                ("" if input_dim == 512 else f"_in{input_dim}"), ins, w, oracle,
                cuda, grad_inputs=("x",), prepare=prep)


def loss_case(B=37, C=6):
    import mmser_b200
    g = torch.Generator().manual_seed(6)
    ins = {"logits": torch.randn(B, C, generator=g) * 5.0,            # some beyond the +-10 clamp
           "unc": torch.rand(B, 1, generator=g),
           "emb": torch.randn(B, 512, generator=g) * 4.0,             # some beyond the +-10 clamp
           "labels": torch.randint(0, C, (B,), generator=g)}
    w = {"prototypes": {"prototypes": torch.randn(C, 512, generator=g) * 0.5}}

    def oracle(i, ws):
        out = O.train_loss(i["logits"], i["unc"], torch.zeros((), dtype=i["logits"].dtype), i["emb"], i["labels"],
                           ws["prototypes"]["prototypes"], C)
        return {k: out[k] for k in ("ce", "focal", "unc_loss", "proto", "loss")}, out["loss"]

    def cuda(i, dtype, dev):
        lg = i["logits"].to(dev).requires_grad_(True)
        un = i["unc"].to(dev).requires_grad_(True)
        em = i["emb"].to(dev).to(dtype).requires_grad_(True)
        pr = w["prototypes"]["prototypes"].to(dev).requires_grad_(True)
        t = mmser_b200.functional.HeadLossFn.apply(lg, un, em, pr, i["labels"].to(dev),
                                                   dict(w_ce=1.0, w_focal=0.3, w_unc=0.05, w_proto=0.01))
        t[4].backward()
        outs = {k: t[j] for j, k in enumerate(("ce", "focal", "unc_loss", "proto", "loss"))}
        return outs, {"prototypes/prototypes": pr.grad}, {"logits": lg.grad, "unc": un.grad, "emb": em.grad}

    return Case("loss", ins, w, oracle, cuda, grad_inputs=("logits", "unc", "emb"))


def supcon_case(B=96, D=512, C=4, temperature=0.07):
    from mmser_b200 import models as M
    g = torch.Generator().manual_seed(21)
    labels = torch.randint(0, C, (B,), generator=g)
    labels[-1] = C                                   # a sample without positives
    ins = {"f": torch.randn(B, D, generator=g) * 2.0, "labels": labels}

    def oracle(i, ws):
        loss = O.supcon_loss(i["f"], i["labels"], temperature)
        return {"supcon": loss}, loss

    def cuda(i, dtype, dev):
        f = i["f"].to(dev).to(dtype).requires_grad_(True)
        loss = M.SupConLoss(temperature)(f, i["labels"].to(dev))
        loss.backward()
        return {"supcon": loss}, {}, {"f": f.grad}

    return Case("supcon", ins, {}, oracle, cuda, grad_inputs=("f",))


def head_case(B=4, Ta=50, Tt=16, C=4, masks=True, seed=1234, L=35, p_drop=0.0):
    import mmser_b200
    w = synth.head_weights(C, L)
    a, t, am, tm, labels = synth.make_inputs(B, Ta, Tt, C, seed=seed, with_masks=masks)
    ins = {"a": a, "t": t, "a_mask": am, "t_mask": tm, "labels": labels}
    keys = ("a_enh", "t_enh", "a_vec", "t_vec", "fused", "logits", "unc", "ce", "focal", "unc_loss", "proto", "loss")

    def oracle(i, ws):
        am_ = None if i["a_mask"] is None else i["a_mask"].to(i["a"].dtype)
        tm_ = None if i["t_mask"] is None else i["t_mask"].to(i["a"].dtype)
        with O.dropout_masks(i.get("_masks")):
            out = O.head_forward(i["a"], i["t"], am_, tm_, i["labels"], ws, C, L)
        return {k: out[k] for k in keys}, out["loss"]

    def cuda(i, dtype, dev):
        head = mmser_b200.FusionHead(C, num_layers=L, dropout=p_drop).to(dev); head.load_group_state(w)
        head.train()
        for grp in ("cross", "fusion", "classifier"):
            _pin_seed(getattr(head, grp), grp, dev)
        out = head(i["a"].to(dev).to(dtype), i["t"].to(dev).to(dtype),
                   None if i["a_mask"] is None else i["a_mask"].to(dev),
                   None if i["t_mask"] is None else i["t_mask"].to(dev), i["labels"].to(dev))
        out["loss"].backward()
        grads = {}
        for grp in head.GROUPS:
            grads.update(_param_grads(grp, getattr(head, grp)))
        return {k: out[k] for k in keys}, grads, {}

    def prep(dev):
        m = cross_masks(dev, p_drop, B, 8, Ta, Tt, 768)
        m.update(fusion_masks(dev, p_drop, B))
        m.update(classifier_masks(dev, p_drop, B, L))
        return {"_masks": m}

    return Case(f"head_B{B}_Ta{Ta}_Tt{Tt}_C{C}{'' if masks else '_nomask'}{'_dropout' if p_drop > 0 else ''}", ins, w, oracle,
                cuda, prepare=prep if p_drop > 0 else None)


ALL_CASES = {
    "adapter": adapter_case,
    "feature_fusion": feature_fusion_case,
    "feature_fusion_asr": lambda: feature_fusion_case(B=5, T=19, F=8),
    "feature_fusion_dropout": lambda: feature_fusion_case(B=4, T=130, F=20, p_drop=0.1),
    "cross_masked": lambda: cross_case(masks=True),
    "cross_nomask": lambda: cross_case(masks=False),
    "cross_long": lambda: cross_case(B=2, Ta=300, Tt=130, masks=True, seed=11),
    # long enough for the tcgen05 attention kernels in BOTH directions, forward and backward (csrc/attention.cu use_tc5:
    # Tq * Tk >= 131072), ragged tails on both sides, masks, dropout on the attention weights
    "cross_xlong": lambda: cross_case(B=1, Ta=610, Tt=250, masks=True, seed=13),
    "cross_xlong_dropout": lambda: cross_case(B=2, Ta=530, Tt=260, masks=True, seed=17, p_drop=0.1),
    "pool": pool_case,
    "pool_nomask": lambda: pool_case(masks=False),
    "fusion": fusion_case,
    # unequal audio / text widths (wav2vec2-base 768 with a 1024-wide text encoder; pooled 1536 / 2048)
    "cross_mixed_dims": lambda: cross_case(B=3, Ta=70, Tt=19, masks=True, text_dim=1024),
    "cross_mixed_dims_dropout": lambda: cross_case(B=2, Ta=66, Tt=35, masks=True, seed=9, p_drop=0.1, text_dim=1024),
    "fusion_mixed_dims": lambda: fusion_case(B=9, text_dim=2048),
    "fusion_mixed_dims_dropout": lambda: fusion_case(B=9, p_drop=0.1, text_dim=2048),
    "classifier": classifier_case,
    # two 128-row clusters of the fused stack kernel, the second one partially filled (rows >= B must stay inert)
    "classifier_b200": lambda: classifier_case(B=200, C=6),
    # the M = 128 row-group path of the fused stack kernel (B > 1152; BASELINE cfg5 runs the classifier at B = 4096)
    # AdvancedOpenMaxClassifier(input_dim != base_dim) (classifier.py:96-105): only input_projection[0] sees input_dim
    "classifier_in384": lambda: classifier_case(B=48, L=4, input_dim=384),
    "classifier_b1280": lambda: classifier_case(B=1280),
    "classifier_b4096": lambda: classifier_case(B=4096, C=6),
    "loss": loss_case,
    "supcon": supcon_case,
    "head_cfg1": lambda: head_case(4, 50, 16, 4, True),
    # (B = 24: with a handful of samples the reference arithmetic itself is 20-25 % away from the exact gradients --
    #  ReLU / argmax flips are O(1/B) -- and the comparison degenerates into noise against noise)
    "head_c6_nomask": lambda: head_case(24, 33, 9, 6, False),
    "head_b48": lambda: head_case(48, 60, 20, 4, True),
    # training mode with dropout active: the oracle is driven by the masks the kernels export (odd Tt: ragged mask pairs)
    "cross_dropout": lambda: cross_case(masks=True, p_drop=0.1),
    "cross_long_dropout": lambda: cross_case(B=2, Ta=300, Tt=130, masks=True, seed=11, p_drop=0.25),
    "fusion_dropout": lambda: fusion_case(p_drop=0.1),
    # (fp32 tier: B = 64 keeps the expected number of ReLU inputs within fp32 rounding of zero well below one per case;
    #  each such unit flips its gate between any two implementations and moves every upstream gradient by ~3e-3)
    "classifier_dropout": lambda: classifier_case(p_drop=0.15),
    "head_dropout": lambda: head_case(24, 50, 16, 4, True, p_drop=0.1),
}
