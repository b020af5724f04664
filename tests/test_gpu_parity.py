"""GPU parity tests: the CUDA path (through the drop-in modules -> C-ABI) against the CPU oracle.
Run on the B200 box:  python -m pytest tests -m gpu -q"""
import os

import pytest
import torch

from tests import parity_cases as PC

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("case", list(PC.ALL_CASES))
def test_parity(case, dtype):
    dev = _dev()
    if dtype == torch.float32 and case in ("classifier_b1280", "classifier_b4096"):
        # these two exist for the M = 128 row-group path of the fused bf16 stack kernel.  In the fp32 tier a batch this
        # large contains several ReLU inputs within fp32 rounding of zero (expected ~2e-7 per hidden unit x 23-73 M units);
        # each flips its gate between any two fp32 implementations and moves all upstream gradients by ~1e-3 -- the
        # comparison is then noise against noise (measured: forward 0.03x of 1e-4, gradients 2.7e-3 vs a CPU-fp32 own
        # error of 0.8e-3).  The fp32 tier at B = 256 is covered by test_baseline_size_fp32_classifier_against_fp64_oracle.
        pytest.skip("bf16-only case (M = 128 path of the fused stack kernel)")
    torch.set_num_threads(max(8, torch.get_num_threads()))
    fails, worst = PC.ALL_CASES[case]().check(dtype, dev)
    assert not fails, f"{case}/{dtype}: {len(fails)} tensors out of tolerance, e.g. {fails[:5]}"


def test_dropout_mask_statistics_and_determinism():
    """The counter-based masks: keep rate = 1 - p, multiplier 1/(1-p), neighbouring draws uncorrelated, different
    sites / seeds give unrelated masks, the same (seed, site) always gives the same mask."""
    from mmser_b200.functional import dropout_mask as dm
    dev = _dev()
    s1 = torch.tensor([0x1234567887654321], dtype=torch.int64, device=dev)
    s2 = s1 + 1
    for p in (0.1, 0.15, 0.5):
        m = dm(s1, 3, p, 4096, 768)
        keep = (m > 0).float()
        n = keep.numel()
        assert abs(keep.mean().item() - (1 - p)) < 4 * (p * (1 - p) / n) ** 0.5 + 2e-5      # 4 sigma + 16-bit rounding of p
        assert torch.allclose(m[m > 0], torch.tensor(1 / (1 - p), device=dev))
        # the two halves of a 32-bit draw, neighbouring pairs, neighbouring rows: correlations within 5 sigma of 0
        k = keep - keep.mean()
        var = (k * k).mean()
        for a, b in ((k[:, 0::2], k[:, 1::2]), (k[:, :-2], k[:, 2:]), (k[:-1], k[1:])):
            assert abs(((a * b).mean() / var).item()) < 5 / (a.numel() ** 0.5)
        # no structure along rows or columns: per-row / per-column keep rates scatter like binomial samples
        for means, cnt in ((keep.mean(1), 768), (keep.mean(0), 4096)):
            z = (means - (1 - p)) / (p * (1 - p) / cnt) ** 0.5
            assert z.abs().max().item() < 5.5 and abs(z.std().item() - 1.0) < 0.1
        # longer lags inside a row and between rows
        for lag in (4, 32, 257):
            assert abs(((k[:, :-lag] * k[:, lag:]).mean() / var).item()) < 5 / ((k.numel()) ** 0.5) * 1.1
        for lag in (2, 16):
            assert abs(((k[:-lag] * k[lag:]).mean() / var).item()) < 5 / ((k.numel()) ** 0.5) * 1.1
        assert torch.equal(m, dm(s1, 3, p, 4096, 768))
        for other in (dm(s1, 4, p, 4096, 768), dm(s2, 3, p, 4096, 768)):
            k2 = (other > 0).float() - keep.mean()
            assert abs(((k * k2).mean() / var).item()) < 5 / (n ** 0.5)
    # odd column count (attention weights with odd Tk): rows do not share pairs
    m = dm(s1, 1, 0.25, 1000, 33)
    assert abs((m > 0).float().mean().item() - 0.75) < 0.02
    assert dm(s1, 1, 0.0, 4, 6).eq(1).all()


def test_dropout_train_eval_and_reseeding():
    """eval() switches dropout off (bit-identical to p = 0); in train() two consecutive steps draw different masks,
    torch.manual_seed makes a run reproducible, and forward / backward of one step agree on the mask (checked by the
    parity cases) -- here: gradient flows only through kept units of fusion.proj_*[0]."""
    import mmser_b200
    from oracle import synth
    dev = _dev()
    C = 4
    w = synth.head_weights(C)
    a, t, am, tm, labels = synth.make_inputs(8, 40, 12, C, seed=5)
    args = (a.to(dev), t.to(dev), am.to(dev), tm.to(dev), labels.to(dev))

    def build(p):
        torch.manual_seed(1234)
        h = mmser_b200.FusionHead(C, dropout=p).to(dev); h.load_group_state(w)
        return h

    ref = build(0.0)(*args)["logits"]
    hd = build(0.2)
    assert torch.equal(hd.eval()(*args)["logits"], ref)           # eval: dropout off, bit-identical
    hd.train()
    l1 = hd(*args)["logits"]; l2 = hd(*args)["logits"]
    assert not torch.equal(l1, ref) and not torch.equal(l1, l2)   # active, and re-drawn every step
    hd2 = build(0.2).train()
    assert torch.equal(hd2(*args)["logits"], l1)                  # same torch seed + same module order -> same masks
    assert torch.equal(hd2(*args)["logits"], l2)


@pytest.mark.parametrize("shape", [
    # (M, N, K, a_trans, b_trans, rowsum)   -- the first three are large enough for the cta_group::2 pair path
    (8192, 768, 1024, False, False, False),      # forward, K-major operands
    (8192, 768, 1024, False, True, False),       # dgrad, B MN-major
    (768, 768, 24000, True, True, True),         # dW with split-K and the fused bias gradient
    (1000, 256, 200, False, False, False),       # ragged M and K tails, single-CTA path
    (200, 256, 3000, True, True, True),          # row tail inside the fused row sum
    (512, 512, 256, True, True, True),           # tiny dW, no split
])
def test_tcgen05_gemm_variants_through_the_cabi(shape):
    """The dominant kernel directly: ser_gemm (bf16 operands, fp32 accumulate) against an fp64 matmul, including the
    CTA-pair path, MN-major operands, split-K and the row-sum by-product that replaces the bias-gradient launches."""
    from mmser_b200 import _lib as L
    dev = _dev()
    M, N, K, ta, tb, rs = shape
    g = torch.Generator(device="cpu").manual_seed(M * 31 + N * 7 + K)
    a = torch.randn((K, M) if ta else (M, K), generator=g).to(dev).bfloat16()
    b = torch.randn((K, N) if tb else (N, K), generator=g).to(dev).bfloat16()
    rowsum = torch.full((M,), 3.0, device=dev) if rs else None
    out = L.gemm(a, b, a_trans=ta, b_trans=tb, out_dtype=torch.float32, rowsum=rowsum)
    A = a.double().t() if ta else a.double()
    Bm = b.double() if tb else b.double().t()
    ref = A @ Bm
    assert ((out.double() - ref).abs().max() / ref.abs().max()).item() < 2e-3
    if rs:
        ref_rs = A.sum(1)
        assert ((rowsum.double() - ref_rs).abs().max() / ref_rs.abs().max()).item() < 1e-4
    # bf16 output with bias + ReLU epilogue (forward shapes only)
    if not ta:
        bias = torch.randn(N, generator=g).to(dev)
        o2 = L.gemm(a, b, b_trans=tb, bias=bias, act=L.ACT_RELU)
        ref2 = torch.relu(ref + bias.double())
        assert ((o2.double() - ref2).abs().max() / ref2.abs().max()).item() < 1e-2


def test_fused_adamw_matches_torch_adamw():
    """FusedAdamW + fused clip_grad_norm_ against the reference's optimizer (torch.optim.AdamW and
    torch.nn.utils.clip_grad_norm_, run on CPU in float64) over several steps with a LambdaLR schedule, two parameter
    groups with their own lr / weight decay (src/train.py:72-83), odd tensor sizes and unaligned views."""
    import mmser_b200
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    shapes = [(768, 256), (256,), (3, 5, 7), (1,), (1027,), (35, 512)]
    flat = torch.randn(sum(int(torch.tensor(s).prod()) for s in shapes) + 3, generator=g)
    ref_params, cu_params, off = [], [], 1                      # offset 1: 4-byte aligned only -> scalar path
    flat_cu = flat.to(dev)
    for s_ in shapes:
        n = int(torch.tensor(s_).prod())
        ref_params.append(flat[off:off + n].view(s_).double().clone().requires_grad_(True))
        cu_params.append(torch.nn.Parameter(flat_cu[off:off + n].view(s_)))
        off += n
    def groups(ps):
        return [dict(params=ps[:3], lr=2e-3, weight_decay=0.05), dict(params=ps[3:], lr=3e-3, weight_decay=0.0)]
    ref_opt = torch.optim.AdamW(groups(ref_params), betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    cu_opt = mmser_b200.optim.FusedAdamW(groups(cu_params), betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    lam = lambda step: 0.5 + 0.1 * step                                          # noqa: E731
    ref_sched = torch.optim.lr_scheduler.LambdaLR(ref_opt, lam)
    cu_sched = torch.optim.lr_scheduler.LambdaLR(cu_opt, lam)
    for it in range(4):
        for rp, cp in zip(ref_params, cu_params):
            gr = torch.randn(rp.shape, generator=g) * (10.0 if it == 2 else 1.0)
            rp.grad = gr.double()
            cp.grad = gr.to(dev)
        if it >= 1:                                                              # clipping active from the second step on
            n_ref = torch.nn.utils.clip_grad_norm_(ref_params, 5.0)
            n_cu = cu_opt.clip_grad_norm_(5.0)
            assert abs(float(n_cu) - float(n_ref)) <= 1e-5 * float(n_ref)
        ref_opt.step(); cu_opt.step()
        ref_sched.step(); cu_sched.step()
        for rp, cp in zip(ref_params, cu_params):
            err = (cp.detach().cpu().double() - rp.detach()).abs().max().item() / (rp.detach().abs().max().item() + 1e-12)
            assert err < 2e-6, (it, tuple(rp.shape), err)
    st = cu_opt.state_dict()
    assert len(st["state"]) == len(shapes) and st["param_groups"][1]["weight_decay"] == 0.0
    # gradients are NOT modified by the fused clipping (the coefficient is applied inside the update)
    assert torch.equal(cu_params[0].grad, cu_params[0].grad.clone())


def test_training_steps_fused_adamw_vs_torch_adamw():
    """Three optimisation steps of the whole head (fp32 tier) with FusedAdamW against the same steps with
    torch.optim.AdamW on an identically initialised head: losses and parameters stay together; the loss goes down."""
    import mmser_b200
    from oracle import synth
    dev = _dev()
    C = 4
    w = synth.head_weights(C)
    a, t, am, tm, labels = synth.make_inputs(12, 40, 12, C, seed=31)
    args = (a.to(dev), t.to(dev), am.to(dev), tm.to(dev), labels.to(dev))

    def groups(h):
        return [dict(params=list(h.classifier.parameters()), lr=1.5e-4, weight_decay=0.06),
                dict(params=[p for n, p in h.named_parameters() if not n.startswith("classifier.")], lr=1e-4, weight_decay=0.05)]

    heads, opts, losses = [], [], [[], []]
    for k in range(2):
        h = mmser_b200.FusionHead(C).to(dev); h.load_group_state(w); h.train()
        heads.append(h)
        opts.append(mmser_b200.optim.FusedAdamW(groups(h)) if k == 0 else torch.optim.AdamW(groups(h)))
    signal = {}                     # per parameter: entries whose gradient was a signal (not rounding noise) at EVERY step
    for _ in range(3):
        for k in range(2):
            opts[k].zero_grad(set_to_none=True)
            out = heads[k](*args)
            out["loss"].backward()
            if k == 1:
                for n, p in heads[1].named_parameters():
                    # (tensors of a few elements cannot be screened against their own scale -- the pooling scorer's
                    #  [1] bias has a mathematically zero gradient and is its own maximum)
                    if p.grad is not None and p.numel() >= 16 and float(p.grad.abs().max()) > 0.0:
                        m = p.grad.abs() > 1e-2 * p.grad.abs().max()
                        signal[n] = m if n not in signal else (signal[n] & m)
            opts[k].step()
            losses[k].append(float(out["loss"].detach()))
    assert losses[0][-1] < losses[0][0]
    for l0, l1 in zip(*losses):
        assert abs(l0 - l1) <= 2e-4 * abs(l1), losses
    # Adam normalises every element's step to ~lr, so entries whose gradient is mathematically zero (softmax-shift
    # biases: pooling attention.2.bias, the key part of the MHA in_proj_bias) turn rounding noise into +-lr steps in
    # ANY implementation; compare where the gradient is a signal
    # (and, within a tensor, entries whose gradient is far below the tensor's scale at some step are dominated by the
    # run-to-run noise of the atomically accumulated gradients -- two runs of the SAME implementation differ there)
    checked = 0
    for (n, p0), (_, p1) in zip(heads[0].named_parameters(), heads[1].named_parameters()):
        if n not in signal or not bool(signal[n].any()):
            continue
        d = (p0 - p1).abs()[signal[n]].max().item()
        assert d <= 2e-5, (n, d)                   # a step is ~1e-4 per element: agreement to a fraction of one step
        checked += int(signal[n].sum())
    assert checked > 500_000


def test_library_side_zeroing_path_matches():
    """grads_zeroed = 0 (what a C-ABI caller without zero-filled gradient buffers passes): the library zeroes its
    split-K / accumulated outputs itself and must produce the same gradients."""
    import mmser_b200
    from mmser_b200 import functional as SF
    from oracle import synth
    dev = _dev()
    C = 4
    w = synth.head_weights(C)
    a, t, am, tm, labels = synth.make_inputs(6, 40, 12, C, seed=41)
    args = (a.to(dev).bfloat16(), t.to(dev).bfloat16(), am.to(dev), tm.to(dev), labels.to(dev))
    grads = []
    try:
        for flag in (1, 0):
            SF.GRADS_ZEROED = flag
            h = mmser_b200.FusionHead(C).to(dev); h.load_group_state(w); h.train()
            h(*args)["loss"].backward()
            grads.append({n: p.grad.detach().clone() for n, p in h.named_parameters() if p.grad is not None})
    finally:
        SF.GRADS_ZEROED = 1
    gmax = max(g.abs().max().item() for g in grads[0].values())
    for n, g1 in grads[0].items():
        assert (g1 - grads[1][n]).abs().max().item() <= 1e-5 * gmax + 1e-4 * g1.abs().max().item(), n


def test_precast_tracks_parameter_updates():
    """FusionHead casts the bf16 operand copies of all modules in one launch per forward (FlatParams.precast); an
    in-place parameter update between two forwards (what an optimizer step is) must be picked up."""
    import mmser_b200
    from oracle import synth
    dev = _dev()
    C = 4
    w = synth.head_weights(C)
    a, t, am, tm, labels = synth.make_inputs(8, 40, 12, C, seed=9)
    args = (a.to(dev).bfloat16(), t.to(dev).bfloat16(), am.to(dev), tm.to(dev))
    head = mmser_b200.FusionHead(C).to(dev).eval(); head.load_group_state(w)
    with torch.no_grad():
        l1 = head(*args)["logits"].clone()
        head.fusion.proj_a[0].weight.mul_(1.5)                      # in-place, like optimizer.step()
        head.adapter_t[0].bias.add_(0.25)
        l2 = head(*args)["logits"].clone()
        w2 = {g: {k: v.clone() for k, v in head_state.items()} for g, head_state in
              ((g, getattr(head, g).state_dict()) for g in head.GROUPS)}
        fresh = mmser_b200.FusionHead(C).to(dev).eval(); fresh.load_group_state(w2)
        l3 = fresh(*args)["logits"]
    assert not torch.equal(l1, l2)
    assert torch.equal(l2, l3)
    # calling a module on its own (no head-level precast) after another update also sees the new weights
    with torch.no_grad():
        f1 = head.fusion(torch.ones(2, 1536, device=dev).bfloat16(), torch.ones(2, 1536, device=dev).bfloat16()).clone()
        head.fusion.proj_t[3].weight.mul_(0.5)
        f2 = head.fusion(torch.ones(2, 1536, device=dev).bfloat16(), torch.ones(2, 1536, device=dev).bfloat16())
    assert not torch.equal(f1, f2)


def test_argmax_bit_exact_and_anchor_zero():
    """Predictions (argmax of logits) are bit-identical to the oracle; anchor loss is exactly 0 with zero grads."""
    import mmser_b200
    from oracle import fusion_head_oracle as O, synth
    dev = _dev()
    C = 4
    w = synth.head_weights(C)
    a, t, am, tm, labels = synth.make_inputs(16, 40, 12, C, seed=77)
    ref = O.head_forward(a, t, am, tm, labels, w, C)
    for dtype in (torch.float32, torch.bfloat16):
        head = mmser_b200.FusionHead(C).to(dev); head.load_group_state(w)
        out = head(a.to(dev).to(dtype), t.to(dev).to(dtype), am.to(dev), tm.to(dev), labels.to(dev))
        assert torch.equal(out["logits"].argmax(1).cpu(), ref["logits"].argmax(1)), dtype
        assert float(out["anchor"]) == 0.0
        out["loss"].backward()
        ac = head.classifier.anchor_clustering
        assert ac.temperature.grad is None
        assert float(ac.class_anchors.grad.abs().max()) == 0.0
        assert all(float(p.grad.abs().max()) == 0.0 for p in ac.anchor_projection.parameters())


def test_fully_padded_rows_give_nan():
    """Padding-mask handling: a sample whose keys are all padded yields NaN exactly where the reference does."""
    from mmser_b200 import models as M
    from oracle import fusion_head_oracle as O, synth
    dev = _dev()
    cw, pw = synth.cross_weights(), synth.pool_weights("pool_a")
    g = torch.Generator().manual_seed(3)
    a, t = torch.randn(3, 9, 768, generator=g), torch.randn(3, 5, 768, generator=g)
    tm = torch.tensor([[1., 1., 1., 0., 0.], [0., 0., 0., 0., 0.], [1., 0., 0., 0., 0.]])
    am = torch.ones(3, 9); am[2, 4:] = 0
    ea, et = O.cross_attention(a, t, am, tm, cw)
    pooled_ref = O.attentive_stats_pooling(t, tm, pw)
    for dtype in (torch.float32, torch.bfloat16):
        m = M.CrossModalAttention(768, 768, dropout=0.0).to(dev); m.load_state_dict(cw)
        oa, ot = m(a.to(dev).to(dtype), t.to(dev).to(dtype), am.to(dev), tm.to(dev))
        assert torch.equal(torch.isnan(oa).cpu(), torch.isnan(ea)), dtype
        assert torch.equal(torch.isnan(ot).cpu(), torch.isnan(et)), dtype
        p = M.AttentiveStatsPooling(768).to(dev); p.load_state_dict(pw)
        pooled = p(t.to(dev).to(dtype), tm.to(dev))
        assert torch.equal(torch.isnan(pooled).cpu(), torch.isnan(pooled_ref)), dtype
        tol = 1e-4 if dtype == torch.float32 else 3e-2
        ok = ~torch.isnan(pooled_ref)
        assert (pooled.float().cpu()[ok] - pooled_ref[ok]).abs().max() <= tol * pooled_ref[ok].abs().max()


@pytest.mark.parametrize("name", ["train_cfg1_small", "train_cfg2_shape", "train_cfg3_c6", "train_nomask", "train_cfg4_long"])
def test_against_reference_golden(name, golden_dir):
    """fp32 tier against the fixtures produced by the reference's own modules (oracle/make_golden.py)."""
    import mmser_b200
    from oracle import synth
    dev = _dev()
    gold = torch.load(os.path.join(golden_dir, f"{name}.pt"), weights_only=False)
    cfg = gold["config"]
    head = mmser_b200.FusionHead(cfg["C"], num_layers=cfg["num_layers"]).to(dev)
    head.load_group_state(synth.head_weights(cfg["C"], cfg["num_layers"]))
    a, t, am, tm, labels = synth.make_inputs(cfg["B"], cfg["Ta"], cfg["Tt"], cfg["C"], cfg["seed"], cfg["with_masks"])
    mv = lambda x: None if x is None else x.to(dev)   # noqa: E731
    out = head(a.to(dev), t.to(dev), mv(am), mv(tm), labels.to(dev))
    out["loss"].backward()
    rel = lambda x, y: (x.detach().double().cpu() - y.double()).abs().max().item() / (y.double().abs().max().item() + 1e-12)  # noqa: E731
    for k in ("logits", "unc", "fused", "a_vec", "t_vec"):
        assert rel(out[k], gold[k]) < 1e-4, k
    assert rel(out["a_enh"][:, :4], gold["a_enh_head"]) < 1e-4
    assert rel(out["t_enh"][:, :4], gold["t_enh_head"]) < 1e-4
    for k in ("ce", "focal", "unc_loss", "proto", "loss"):
        assert abs(float(out[k]) - gold[k]) <= 1e-4 * max(1.0, abs(gold[k])), k
    assert torch.equal(out["logits"].argmax(1).cpu(), gold["logits"].argmax(1))
    # gradients: norms within 1e-3 (ReLU-flip noise of fp32 arithmetic, see tests/parity_cases.py), probes alike
    checked = 0
    for key, summ in gold["grads"].items():
        grp, pname = key.split("/", 1)
        p = dict(getattr(head, grp).named_parameters())[pname]
        if summ is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        if summ["norm"] < 1e-6:
            assert float(p.grad.double().norm()) < 1e-4, key
            continue
        assert abs(float(p.grad.double().norm()) - summ["norm"]) <= 2e-3 * summ["norm"], key
        checked += 1
    assert checked > 300


def test_mixed_dims_against_reference_golden(golden_dir):
    """CrossModalAttention(768, 1024) and FusionLayer(1536, 2048, 512) -- unequal audio / text widths -- against the
    fixture produced by the reference's own modules (oracle/make_golden.py run_mixed_dims_case), fp32 tier."""
    from mmser_b200 import models as M
    from oracle import synth
    dev = _dev()
    gold = torch.load(os.path.join(golden_dir, "mixed_dims_small.pt"), weights_only=False)
    cfg, ins = gold["config"], gold["inputs"]
    cross = M.CrossModalAttention(cfg["audio_dim"], cfg["text_dim"], dropout=0.0).to(dev)
    cross.load_state_dict(synth.cross_weights(audio_dim=cfg["audio_dim"], text_dim=cfg["text_dim"]))
    fusion = M.FusionLayer(2 * cfg["audio_dim"], 2 * cfg["text_dim"], 512).to(dev)
    fusion.load_state_dict(synth.fusion_weights(audio_dim=2 * cfg["audio_dim"], text_dim=2 * cfg["text_dim"]))
    cross.eval(); fusion.eval()
    x = {k: ins[k].to(dev).requires_grad_(True) for k in ("a", "t", "av", "tv")}
    ea, et = cross(x["a"], x["t"], ins["a_mask"].to(dev), ins["t_mask"].to(dev))
    fused = fusion(x["av"], x["tv"])
    ((ea * ins["ua"].to(dev)).sum() + (et * ins["ut"].to(dev)).sum() + (fused * ins["up"].to(dev)).sum()).backward()
    rel = lambda g, r: (g.detach().double().cpu() - r.double()).abs().max().item() / (r.double().abs().max().item() + 1e-12)  # noqa: E731
    assert rel(ea, gold["audio_enh"]) < 1e-4 and rel(et, gold["text_enh"]) < 1e-4 and rel(fused, gold["fused"]) < 1e-4
    for k in ("a", "t", "av", "tv"):
        assert rel(x[k].grad, gold["din"][k]) < 1e-4, k
    checked = 0
    for grp, mod in (("cross", cross), ("fusion", fusion)):
        for n, p in mod.named_parameters():
            summ = gold["grads"][f"{grp}/{n}"]
            if summ["norm"] < 1e-6:           # mathematically zero (MHA key bias: softmax shift invariance)
                assert p.grad.double().norm().item() < 1e-4, n
                continue
            assert abs(p.grad.double().norm().item() - summ["norm"]) <= 2e-3 * summ["norm"], n
            if "full" in summ:
                assert rel(p.grad, summ["full"]) < 1e-3, n
            checked += 1
    assert checked >= 40


@pytest.mark.parametrize("name", ["feature_fusion_combined", "feature_fusion_asr"])
def test_feature_fusion_against_reference_golden(name, golden_dir):
    """SURVEY 8(f) rank 1 against the fixture produced by the reference's own nn.Sequential, called the way the
    reference calls it (one utterance at a time on the pre-concatenated [frames, hid + F] input, audio_encoder.py:131)
    and batched; eval mode pins the values, train mode the dropout semantics (output is 0 or eval / (1 - p))."""
    from mmser_b200 import models as M
    from oracle import synth
    dev = _dev()
    gold = torch.load(os.path.join(golden_dir, f"{name}.pt"), weights_only=False)
    cfg = gold["config"]
    B, T, hid, F = cfg["B"], cfg["T"], cfg["hid"], cfg["F"]
    g = torch.Generator().manual_seed(cfg["seed"])
    seq = torch.randn(B, T, hid, generator=g)
    feats = torch.rand(B, F, generator=g) * 2.0 - 0.5
    up = torch.randn(B, T, hid, generator=g)
    m = M.UtteranceFeatureFusion(hid, F).to(dev)
    m.load_state_dict(synth.feature_fusion_weights(cfg["attr"], F, seed=0, hidden=hid))
    rel = lambda x, y: (x.detach().double().cpu() - y.double()).abs().max().item() / (y.double().abs().max().item() + 1e-12)  # noqa: E731
    m.eval()
    x = seq.to(dev).requires_grad_(True)
    ys = []
    for i in range(B):
        fused_input = torch.cat([x[i], feats[i].to(dev).unsqueeze(0).expand(T, -1)], dim=-1)
        ys.append(m(fused_input))
    y = torch.stack(ys)
    (y * up.to(dev)).sum().backward()
    ref = gold["eval"]
    assert rel(y, ref["y"]) < 1e-4 and rel(x.grad, ref["dx"]) < 1e-4
    assert rel(m[0].weight.grad, ref["dw"]) < 1e-4 and rel(m[0].bias.grad, ref["db"]) < 1e-4
    # batched call = the per-utterance calls
    m.zero_grad()
    xb = seq.to(dev).requires_grad_(True)
    yb = m(xb, feats.to(dev))
    (yb * up.to(dev)).sum().backward()
    assert rel(yb, ref["y"]) < 1e-4 and rel(xb.grad, ref["dx"]) < 1e-4 and rel(m[0].weight.grad, ref["dw"]) < 1e-4
    # train mode: every element is dropped or the eval value / (1 - p); about p of them dropped
    m.train()
    yt = m(seq.to(dev), feats.to(dev)).cpu()
    ye = ref["y"]
    kept = yt != 0
    assert (yt[kept] - ye[kept] / 0.9).abs().max() <= 1e-4 * ye.abs().max()
    frac = 1.0 - kept[ye > 0].float().mean().item()
    assert 0.05 < frac < 0.15, frac
    # bf16 tier
    m.eval()
    y16 = m(seq.to(dev).bfloat16(), feats.to(dev))
    assert y16.dtype == torch.bfloat16
    assert (y16.float().cpu() - ye).norm() <= 2e-2 * ye.norm()


def test_eval_path_against_reference_golden(golden_dir):
    """cfg5 semantics: fitted OpenMax, 5-view TTA mean, temperature sweep, softmax/argmax/energy."""
    import mmser_b200
    from mmser_b200 import functional as SF
    from oracle import synth
    dev = _dev()
    gold = torch.load(os.path.join(golden_dir, "eval_cfg5_small.pt"), weights_only=False)
    cfg = gold["config"]
    C = cfg["C"]
    clf = mmser_b200.models.AdvancedOpenMaxClassifier(512, C, dropout=0.15).to(dev)
    clf.load_state_dict(synth.classifier_weights(C))
    clf.eval()
    g = torch.Generator().manual_seed(cfg["seed"])
    val_fused = torch.randn(64, 512, generator=g)
    val_labels = torch.randint(0, C, (64,), generator=g)
    rel = lambda x, y: (x.detach().double().cpu() - y.double()).abs().max().item() / (y.double().abs().max().item() + 1e-12)  # noqa: E731
    with torch.no_grad():
        clf(val_fused.to(dev), use_openmax=False)
        feats = clf.last_features
        assert rel(feats, gold["val_features"]) < 1e-4
        clf.fit_weibull(feats, val_labels.to(dev))
        for k in synth.CLASSIFIER_BUFFERS:
            assert rel(getattr(clf, k), gold["weibull"][k]) < 1e-3, k
        fused_views = torch.randn(cfg["views"], cfg["B"], 512, generator=g)
        labels = torch.randint(0, C, (cfg["B"],), generator=g)
        lv = torch.stack([clf(fused_views[v].to(dev)) for v in range(cfg["views"])])
        lp = torch.stack([clf(fused_views[v].to(dev), use_openmax=False) for v in range(cfg["views"])])
        assert rel(lp, gold["logits_plain"]) < 1e-4
        assert rel(lv, gold["logits_views"]) < 1e-3
        T = SF.find_optimal_temperature(lp[0], labels.to(dev))
        assert abs(T - gold["temperature"]) < 1e-3 * gold["temperature"]
        post = SF.eval_post(lv, T)
        assert rel(post["mean_logits"], gold["mean_logits"]) < 1e-3
        assert rel(post["probs"], gold["probs"]) < 1e-3
        assert torch.equal(post["preds"].cpu(), gold["preds"])
        assert rel(post["energy"], gold["energy"]) < 1e-3


def test_weibull_fitting_pass_matches_hand_unrolled_walk():
    """FusionHead.fit_weibull_on = the reference's after-last-epoch pass (src/train.py:204-245): same buffers as walking
    the classifier's children one by one over the same validation batches."""
    import mmser_b200
    from oracle import synth
    dev = _dev()
    C = 4
    head = mmser_b200.FusionHead(C, num_layers=35, dropout="reference").to(dev)
    head.load_group_state(synth.head_weights(C, 35))
    batches = []
    for i in range(3):
        a, t, am, tm, y = synth.make_inputs(6, 40, 12, C, 2000 + i, True)
        batches.append((a.to(dev), t.to(dev), am.to(dev), tm.to(dev), y.to(dev)))
    head.train()
    n = head.fit_weibull_on(batches)
    assert n == 18 and head.training
    got = {k: getattr(head.classifier, k).clone() for k in synth.CLASSIFIER_BUFFERS}
    head.eval()
    feats = []
    with torch.no_grad():
        for a, t, am, tm, y in batches:
            f = head.features(a, t, am, tm)["fused"]
            dc = head.classifier.deep_classifier
            for layer in dc.input_projection:
                f = layer(f)
            for blk, ln in zip(dc.residual_layers, dc.layer_norms):
                f = blk(ln(f))
            for i in range(4):
                f = dc.output_projection[i](f)
            feats.append(f)
        head.classifier.fit_weibull(torch.cat(feats), torch.cat([b[4] for b in batches]))
    for k in synth.CLASSIFIER_BUFFERS:
        ref = getattr(head.classifier, k)
        assert (got[k] - ref).abs().max() <= 1e-3 * ref.abs().max().clamp_min(1e-6), k
    assert float(got["weibull_beta"].min()) > 0 and float(got["activation_vectors"].abs().max()) > 0


def test_late_ood_against_reference_golden(golden_dir):
    """LateStageOODDetector (one launch) against the fixture produced by the reference's detector."""
    from mmser_b200 import models as M
    from oracle import fusion_head_oracle as O
    dev = _dev()
    gold = torch.load(os.path.join(golden_dir, "late_ood_small.pt"), weights_only=False)
    cfg = gold["config"]
    det = M.LateStageOODDetector(cfg["C"], cfg["D"]).to(dev)
    det.load_state_dict(gold["state"])
    logits, feats = gold["logits"].to(dev), gold["features"].to(dev)
    rel = lambda x, y: (x.detach().double().cpu() - y.double()).abs().max().item() / (y.double().abs().max().item() + 1e-12)  # noqa: E731
    s = det.scores(logits, feats)
    ref = O.late_ood_scores(gold["logits"].double(), gold["features"].double(), {k: v.double() for k, v in gold["state"].items()})
    assert rel(s["energy"], gold["energy"]) < 1e-4 and rel(s["distances"], gold["distances"]) < 1e-4
    assert rel(s["min_distance"], gold["min_distance"]) < 1e-4
    for k in ("energy_norm", "distance_norm", "combined"):
        assert rel(s[k], ref[k]) < 1e-4, k
    assert torch.equal(s["is_ood"].cpu(), ref["combined"] < 0.5)
    e, _ = det.energy_detector(logits)
    d, md = det.prototype_detector(feats)
    assert rel(e, gold["energy"]) < 1e-4 and rel(d, gold["distances"]) < 1e-4 and rel(md, gold["min_distance"]) < 1e-4
    res, r = det(logits, feats), gold["result"]
    assert res.is_ood == r["is_ood"] and res.reason.value == r["reason"]
    for k in ("energy_score", "prototype_distance", "combined_score", "confidence_score"):
        assert abs(getattr(res, k) - r[k]) <= 1e-4 * max(1e-3, abs(r[k])), k
    # bf16 features (what the bf16 tier's classifier hands out): distances within the tier's tolerance
    s16 = det.scores(logits, feats.bfloat16())
    assert rel(s16["distances"], gold["distances"]) < 2e-2


def test_children_callable_like_reference():
    """src/train.py:221-236 walks the classifier's children one by one; they must stay callable and agree with the
    fused forward."""
    import mmser_b200
    from oracle import synth
    dev = _dev()
    clf = mmser_b200.models.AdvancedOpenMaxClassifier(512, 4, dropout=0.15).to(dev).eval()
    clf.load_state_dict(synth.classifier_weights(4))
    x = torch.randn(6, 512, device=dev)
    with torch.no_grad():
        f = x
        for layer in clf.deep_classifier.input_projection:
            f = layer(f)
        for blk, ln in zip(clf.deep_classifier.residual_layers, clf.deep_classifier.layer_norms):
            f = blk(ln(f))
        for i in range(4):
            f = clf.deep_classifier.output_projection[i](f)
        clf(x, use_openmax=False)
    assert (f - clf.last_features).abs().max() <= 1e-4 * clf.last_features.abs().max()


def test_cpu_tensors_fail_loudly():
    import mmser_b200
    from mmser_b200 import _lib
    pool = mmser_b200.models.AttentiveStatsPooling(768)
    with pytest.raises(_lib.SerError):
        pool(torch.randn(2, 3, 768))


def _record_parity(tag, res):
    """Keep the measured numbers: gpurun_out/ travels back from the GPU box (tools/gpu_parity_report.py formats them)."""
    import json
    out = os.environ.get("SER_PARITY_OUT") or (os.path.join("gpurun_out") if os.path.isdir("gpurun_out") else None)
    if out:
        with open(os.path.join(out, f"parity_{tag}.json"), "w") as f:
            json.dump(res, f, indent=1)
    for k, v in res.items():
        print(f"    {tag:28s} {k:22s} |G-E|/|E| = {v['ge']:.3e}   |R-E|/|E| = {v['re']:.3e}   |G-R|/|R| = {v['gr']:.3e}   n = {v['n']}")


@pytest.mark.parametrize("which", ["head_cfg2", "classifier_b256"])
def test_baseline_size_bf16_against_fp64_oracle(which):
    """BASELINE.json cfg2 size (B = 256, Ta = 250, Tt = 64, C = 4), bf16 tier, against the fp64 oracle with NO noise term
    and NO outlier set-aside (Frobenius over all tensors of a parameter group).

    What holds and what does not (DESIGN.md section 4 has the measured table):
      * every forward quantity meets north_star's plain 2e-2;
      * gradients do NOT: the head ends in 35 LayerNorm'd ReLU blocks, a perturbation of relative size eps in a
        pre-activation flips a fraction ~eps of the gates and moves the gradient by ~sqrt(eps) -- bf16 operands
        (eps = 2^-9) give 20-30 %, at any batch size, for ANY implementation that rounds operands to bf16, the
        reference under torch.autocast(bfloat16) included (column R: the oracle with bf16-rounded Linear operands).
        The assertion therefore is that the CUDA path is no further from the exact gradient than that autocast
        arithmetic (x1.5 margin: the kernels also round dY and the stored activations), and the numbers are recorded."""
    dev = _dev()
    torch.set_num_threads(max(8, torch.get_num_threads()))
    case = PC.head_case(256, 250, 64, 4, True, seed=1235) if which == "head_cfg2" else PC.classifier_case(B=256)
    res = case.group_errors(torch.bfloat16, dev)
    _record_parity(f"{which}_bf16", res)
    for k, v in res.items():
        if k.startswith("out/"):
            assert v["ge"] <= 2e-2, (k, v)
        else:
            assert v["ge"] <= 1.5 * v["re"] + 2e-2, (k, v)


def test_baseline_size_fp32_classifier_against_fp64_oracle():
    """The same B = 256 classifier problem in the fp32 tier: forward at 1e-4; gradients within 1e-2 Frobenius -- the
    fp32 CPU reference itself sits at ~3e-3 there (one gate flip in 4.6 M hidden units moves every upstream gradient)."""
    dev = _dev()
    res = PC.classifier_case(B=256).group_errors(torch.float32, dev)
    _record_parity("classifier_b256_f32", res)
    for k, v in res.items():
        assert v["ge"] <= (1e-4 if k.startswith("out/") else 1e-2), (k, v)


def test_data_parallel_gradients_two_gpus():
    """DataParallelHead.train_step on 2 real GPUs over NCCL (different shards per rank, overlapped and non-overlapped
    bucket all-reduces, both tiers) against one process on the concatenated batch: tools/dp_check.py under torchrun."""
    import subprocess
    import sys
    _dev()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(root, "tools", "dp_check.py")],
                       cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.count("dp_check") >= 3, (r.stdout[-1500:], r.stderr[-3000:])


def test_large_shape_properties():
    """BASELINE.json cfg2 size (B=256, Ta=250, Tt=64), bf16: finite outputs, softmax weights sum to one, loss terms
    consistent, gradient of every parameter finite; fp32-vs-bf16 logits agree within the bf16 tolerance."""
    import mmser_b200
    from oracle import synth
    dev = _dev()
    C = 4
    head = mmser_b200.FusionHead(C).to(dev); head.load_group_state(synth.head_weights(C))
    a, t, am, tm, labels = synth.make_inputs(256, 250, 64, C, seed=1235)
    out = head(a.to(dev).bfloat16(), t.to(dev).bfloat16(), am.to(dev), tm.to(dev), labels.to(dev))
    out["loss"].backward()
    assert all(torch.isfinite(out[k].float()).all() for k in ("a_enh", "t_enh", "fused", "logits", "unc", "loss"))
    assert all(torch.isfinite(p.grad).all() for n, p in head.named_parameters() if p.grad is not None)
    total = float(out["ce"]) + 0.3 * float(out["focal"]) + 0.05 * float(out["unc_loss"]) + 0.01 * float(out["proto"])
    assert abs(total - float(out["loss"])) < 1e-4 * max(1.0, abs(total))
    head.zero_grad(set_to_none=True)
    out32 = head(a.to(dev)[:32], t.to(dev)[:32], am.to(dev)[:32], tm.to(dev)[:32], labels.to(dev)[:32])
    d = (out["logits"][:32] - out32["logits"]).abs().max() / out32["logits"].abs().max()
    assert float(d) < 2e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_unpack_frames_bit_exact(dtype):
    """Packed valid frames -> zero-padded batch + mask (ser_unpack_frames) against the padding the encoders do
    (src/models/audio_encoder.py:140-163): bit-exact, including an empty utterance, a full-length one and T = 1."""
    from mmser_b200.functional import pack_frames, unpack_frames
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    for B, T, D in ((7, 53, 768), (3, 1, 768), (5, 130, 256)):
        lens = torch.randint(0, T + 1, (B,), generator=g)
        lens[0], lens[-1] = T, 0
        mask = (torch.arange(T)[None, :] < lens[:, None]).float()
        x = (torch.randn(B, T, D, generator=g) * mask[:, :, None]).to(dtype)
        packed, offsets = pack_frames(x, mask)                       # host side
        assert packed.shape[0] == int(lens.sum()) and offsets[-1] == lens.sum()
        out, m = unpack_frames(packed.to(dev), offsets.to(dev), T)
        assert torch.equal(out.cpu(), x) and torch.equal(m.cpu(), mask)
        # into preallocated (dirty) buffers, as the benchmark's copy stream does
        out2 = torch.full((B, T, D), 7.0, device=dev, dtype=dtype)
        m2 = torch.full((B, T), 7.0, device=dev)
        unpack_frames(packed.to(dev), offsets.to(dev), T, out=out2, mask_out=m2)
        assert torch.equal(out2.cpu(), x) and torch.equal(m2.cpu(), mask)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_forward_views_equals_per_view_forward(dtype):
    """FusionHead.forward_views (test-time augmentation, src/eval.py:174-190: V audio views, one text batch, the text-side
    adapter / projections computed once) against one full head call per view: bit-identical logits, with fitted
    (non-default) OpenMax buffers, and a second text batch through a fresh cache."""
    import mmser_b200
    from oracle import synth
    dev = _dev()
    C = 6
    head = mmser_b200.FusionHead(C, num_layers=3, dropout="reference").to(dev)
    head.load_group_state(synth.head_weights(C, 3))
    head.eval()
    a, t, am, tm, labels = synth.make_inputs(6, 70, 19, C, seed=21)
    a, t, am, tm, labels = a.to(dev).to(dtype), t.to(dev).to(dtype), am.to(dev), tm.to(dev), labels.to(dev)
    with torch.no_grad():
        head.fit_weibull_on([(a, t, am, tm, labels)])
        g = torch.Generator(device="cpu").manual_seed(3)
        views = [a] + [(a.float() + 0.05 * torch.randn(a.shape, generator=g).to(dev) * am[..., None]).to(dtype) for _ in range(3)]
        ref = torch.stack([head(v, t, am, tm)["logits"] for v in views])
        got = head.forward_views(views, t, am, tm)
        assert torch.equal(got, ref)
        t2 = torch.flip(t, dims=[0]).contiguous()
        assert torch.equal(head.forward_views(views[:2], t2, am, torch.flip(tm, dims=[0]).contiguous()),
                           torch.stack([head(v, t2, am, torch.flip(tm, dims=[0]).contiguous())["logits"] for v in views[:2]]))
