"""The same-box PyTorch-eager baseline (baseline/eager_head.py, stock torch modules) must be the arithmetic of the
reference: it is held to the CPU oracle (itself pinned to the reference's own modules by tests/golden/)."""
import pytest
import torch

from baseline.eager_head import EagerHead
from oracle import fusion_head_oracle as O
from oracle import synth


@pytest.mark.parametrize("graph_safe", [False, True])
@pytest.mark.parametrize("cfg", [(6, 40, 12, 4, True), (5, 33, 9, 6, False)])
def test_eager_baseline_equals_oracle(cfg, graph_safe):
    B, Ta, Tt, C, masks = cfg
    torch.manual_seed(0)
    w = synth.head_weights(C, 35)
    a, t, am, tm, labels = synth.make_inputs(B, Ta, Tt, C, seed=77, with_masks=masks)
    ws = {g: {k: v.clone().requires_grad_(v.is_floating_point() and k not in synth.CLASSIFIER_BUFFERS)
              for k, v in grp.items()} for g, grp in w.items()}
    ref = O.head_forward(a, t, am, tm, labels, ws, C)
    ref["loss"].backward()
    head = EagerHead(C, graph_safe=graph_safe)
    head.load_group_state(w)            # strict: the module tree carries the reference's state_dict keys
    head.eval()                         # dropout off (rates are 0 anyway)
    out = head(a, t, am, tm, labels)
    out["loss"].backward()
    rel = lambda x, y: (x.detach() - y.detach()).abs().max().item() / (y.detach().abs().max().item() + 1e-12)   # noqa: E731
    for k in ("a_enh", "t_enh", "a_vec", "t_vec", "fused", "logits", "unc", "ce", "focal", "unc_loss", "proto", "loss"):
        assert rel(out[k], ref[k]) < 2e-5, k
    checked = 0
    for g in EagerHead.GROUPS:
        params = dict(getattr(head, g).named_parameters())
        for k, v in ws[g].items():
            if not v.requires_grad or k.endswith("temperature"):
                continue
            gr = v.grad if v.grad is not None else torch.zeros_like(v)
            scale = max(x.grad.abs().max().item() for x in ws[g].values() if x.grad is not None)
            pg = params[k].grad if params[k].grad is not None else torch.zeros_like(v)
            assert (pg - gr).abs().max().item() <= 2e-4 * scale + 1e-9, (g, k)
            checked += 1
    assert checked > 300
