"""CPU-side tests: the C-ABI library loads and exports every symbol include/ser_head.h declares, the ctypes
layouts match, the drop-in modules expose the reference's parameter names, and the host logic around the kernels
(flat parameter packing, data-parallel reduction order, global-batch loss algebra) is correct.  No GPU needed."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as G
    G.build()
    from mmser_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "ser_head.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(ser_\w+)\s*\(", header)))
    assert len(declared) >= 30
    dll = ctypes.CDLL(lib.LIB_PATH)
    missing = [n for n in declared if not hasattr(dll, n)]
    assert not missing, missing
    assert dll.ser_version() >= 100


def test_struct_layouts_match_the_compiled_library(lib):
    l = lib.load()
    for i, S in enumerate((lib.GemmDesc, lib.AdapterDesc, lib.XattnDesc, lib.AspDesc, lib.FusionDesc, lib.ClfDesc,
                           lib.LossDesc, lib.FeatFuseDesc, lib.AttnDesc)):
        assert l.ser_desc_size(i) == ctypes.sizeof(S), S.__name__


def test_blackwell_instructions_present(lib):
    """The GEMM path, the fused classifier stack and the attention core (attention_tc5.cu: forward, dQ, dK/dV) must be
    tcgen05 + TMA + TMEM (SASS: UTCHMMA / UTMALDG / LDTM ...).  The legacy mma.sync path (SASS HMMA) is allowed in
    exactly one place: the small-shape attention kernels of attention_tc.cu, and nowhere else."""
    sass = subprocess.run(["cuobjdump", "-sass", lib.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump unavailable")
    funcs, cur = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    body = {k: "\n".join(v) for k, v in funcs.items()}
    gemm = [k for k in body if "gemm_tc_kernel" in k]
    stack = [k for k in body if "clf_stack_fwd_kernel" in k or "clf_stack_bwd_kernel" in k]
    assert gemm and len(stack) == 4          # forward / backward x dropout off / on
    for k in gemm:
        assert "UTCHMMA" in body[k] and "UTMALDG" in body[k] and "LDTM" in body[k] and "UTMASTG" in body[k], k
    for k in stack:
        assert "UTCHMMA" in body[k] and "LDTM" in body[k] and "UTMASTG" in body[k] and "UBLKCP" in body[k], k
    attn5 = [k for k in body if "attn5_fwd_kernel" in k or "attn5_bwd_dq_kernel" in k or "attn5_bwd_dkv_kernel" in k]
    assert len(attn5) == 8                   # forward, dQ, dK/dV (one / two threads per key row) x dropout off / on
    for k in attn5:
        assert "UTCHMMA" in body[k] and "UTMALDG" in body[k] and "LDTM" in body[k], k
        assert "HMMA." not in body[k].replace("UTCHMMA", ""), k
    legacy = [k for k, b in body.items() if "HMMA." in b.replace("UTCHMMA", "")]
    assert legacy and all("attn_tc_" in k for k in legacy), legacy


def test_state_dict_keys_match_reference_goldens(lib, golden_dir):
    """Checkpoint compatibility: every parameter the reference's modules own exists under the same name/shape."""
    import mmser_b200
    gold = torch.load(os.path.join(golden_dir, "train_cfg1_small.pt"), weights_only=False)
    head = mmser_b200.FusionHead(4)
    ours = {f"{g}/{n}": tuple(p.shape) for g in head.GROUPS for n, p in getattr(head, g).named_parameters()}
    ref = {k: (tuple(v["shape"]) if v is not None else None) for k, v in gold["grads"].items()}
    assert set(ours) == set(ref)
    for k, shp in ref.items():
        if shp is not None:
            assert ours[k] == shp, k
    assert sum(p.numel() for p in head.parameters()) == 24359690
    clf_keys = set(head.classifier.state_dict().keys())
    assert {"weibull_alpha", "weibull_beta", "weibull_tau", "activation_vectors"} <= clf_keys and len(clf_keys) == 304


def test_reference_checkpoint_wire_format(lib):
    """src/train.py:247-262 / src/eval.py:109-123: dict of state_dicts; the adapters travel inside the encoders' entries."""
    import mmser_b200
    from oracle import synth
    C = 4
    w = synth.head_weights(C, 2)
    enc_a = {"encoder.layers.0.weight": torch.randn(3, 3), "pool.attention.0.weight": torch.randn(128, 768)}
    enc_a.update({f"adapter.{k}": v for k, v in w["adapter_a"].items()})
    enc_t = {f"adapter.{k}": v for k, v in w["adapter_t"].items()}
    ckpt = {"audio_encoder": enc_a, "text_encoder": enc_t, "epoch": 3, "f1": 0.5}
    ckpt.update({g: w[g] for g in ("cross", "pool_a", "pool_t", "fusion", "classifier", "prototypes")})
    head = mmser_b200.FusionHead(C, num_layers=2)
    head.load_checkpoint_state(ckpt)
    out = head.checkpoint_state(encoders={"audio_encoder": enc_a}, epoch=4, f1=0.75)
    assert set(out) == {"audio_encoder", "text_encoder", "cross", "pool_a", "pool_t", "fusion", "classifier",
                        "prototypes", "epoch", "f1"}
    assert out["epoch"] == 4 and set(out["audio_encoder"]) == set(enc_a) and set(out["text_encoder"]) == set(enc_t)
    for grp in ("cross", "pool_a", "pool_t", "fusion", "classifier", "prototypes"):
        assert set(out[grp]) == set(w[grp]), grp
        for k in w[grp]:
            assert torch.equal(out[grp][k].cpu(), w[grp][k]), (grp, k)
    for k, v in w["adapter_a"].items():
        assert torch.equal(out["audio_encoder"][f"adapter.{k}"], v)
    bad = dict(ckpt, fusion={k: v for k, v in w["fusion"].items() if k != "gate_a.0.bias"})
    with pytest.raises(RuntimeError):
        head.load_checkpoint_state(bad)


def test_flat_param_packing(lib):
    from mmser_b200._params import ALIGN, FlatParams
    import mmser_b200
    m = mmser_b200.models.CrossModalAttention(768, 768)
    fp = m._flat
    assert fp.names[:3] == ["q_a.weight", "k_a.weight", "v_a.weight"]
    assert all(o % ALIGN == 0 for o in fp.offsets)
    i = fp.index["q_a.weight"]
    assert fp.offsets[i + 1] - fp.offsets[i] == 256 * 768 and fp.offsets[i + 2] - fp.offsets[i] == 2 * 256 * 768
    flat = torch.arange(fp.total, dtype=torch.float32)
    v = fp.view(flat, "q_a.weight", 3)
    assert v.shape == (768, 768) and v[256, 0] == fp.offsets[i + 1]
    with pytest.raises(lib.SerError):
        fp.ensure()                      # CPU parameters: no CPU path, fail loudly


def test_unequal_audio_text_widths_keep_the_reference_layout(lib):
    """CrossModalAttention / FusionLayer with audio_dim != text_dim (reference constructors cross_attention.py:7-30,
    fusion.py:6-16): reference parameter names and shapes, packed q|k|v views per modality, descriptor fields present."""
    import mmser_b200
    from mmser_b200 import synth
    m = mmser_b200.models.CrossModalAttention(768, 1024)
    shapes = {n: tuple(p.shape) for n, p in m.named_parameters()}
    assert shapes["q_a.weight"] == (256, 768) and shapes["k_a.weight"] == (256, 768) and shapes["v_a.weight"] == (256, 768)
    assert shapes["q_t.weight"] == (256, 1024) and shapes["k_t.weight"] == (256, 1024) and shapes["v_t.weight"] == (256, 1024)
    assert shapes["out_a.weight"] == (768, 256) and shapes["out_t.weight"] == (1024, 256)
    assert shapes["norm_a.weight"] == (768,) and shapes["norm_t.weight"] == (1024,)
    m.load_state_dict(synth.cross_weights(text_dim=1024))
    fp = m._flat
    flat = torch.arange(fp.total, dtype=torch.float32)
    assert fp.view(flat, "q_t.weight", 3).shape == (768, 1024) and fp.view(flat, "q_a.weight", 3).shape == (768, 768)
    f = mmser_b200.models.FusionLayer(1536, 2048, 512)
    f.load_state_dict(synth.fusion_weights(text_dim=2048))
    assert tuple(f.proj_a[0].weight.shape) == (512, 1536) and tuple(f.proj_t[0].weight.shape) == (512, 2048)
    assert "Dt" in dict(lib.XattnDesc._fields_) and "Din_t" in dict(lib.FusionDesc._fields_)


def test_no_cpu_fallback(lib):
    import mmser_b200
    with pytest.raises(lib.SerError):
        mmser_b200.models.FusionLayer(1536, 1536, 512).eval()(torch.randn(2, 1536), torch.randn(2, 1536))
    with pytest.raises(lib.SerError):
        lib.gemm(torch.randn(4, 8), torch.randn(4, 8))
    fuse = mmser_b200.models.UtteranceFeatureFusion(768, 20)
    assert list(fuse.state_dict()) == ["0.weight", "0.bias"] and fuse[0].weight.shape == (768, 788)   # audio_encoder.py:47-52
    with pytest.raises(lib.SerError):
        fuse.eval()(torch.randn(2, 5, 768), torch.randn(2, 20))
    with pytest.raises(lib.SerError):            # shape errors are reported before anything reaches the device
        fuse(torch.randn(2, 5, 768), torch.randn(3, 20))


def test_dropout_is_active_only_in_training_mode(lib):
    """nn.Dropout semantics of the drop-in modules: p > 0 takes effect in train() only, and never silently on CPU."""
    import mmser_b200
    from mmser_b200.models._common import DropoutSeed, active_dropout
    m = mmser_b200.models.CrossModalAttention(768, 768, dropout=0.1)
    assert active_dropout(m.train(), m.dropout.p) == pytest.approx(0.1)
    assert active_dropout(m.eval(), m.dropout.p) == 0.0
    with pytest.raises(lib.SerError):            # CPU tensors: no fallback, dropout or not
        m.train()(torch.randn(1, 2, 768), torch.randn(1, 2, 768))
    assert DropoutSeed().counter is None         # the seed is drawn lazily, on the first training forward
    f = mmser_b200.models.FusionLayer(1536, 1536, 512).train()
    f.proj_a[2].p = 0.3                          # the fused kernel has one rate for both branches
    with pytest.raises(lib.SerError):
        f(torch.randn(2, 1536), torch.randn(2, 1536))


def test_fused_adamw_host_side(lib):
    """Constructor contract of the torch.optim.AdamW replacement; CPU tensors fail loudly (no fallback)."""
    import mmser_b200
    w = torch.nn.Parameter(torch.zeros(4, 4))
    with pytest.raises(ValueError):
        mmser_b200.optim.FusedAdamW([w], lr=-1.0)
    opt = mmser_b200.optim.FusedAdamW([dict(params=[w], lr=1e-3, weight_decay=0.1)], weight_decay=0.05)
    assert opt.param_groups[0]["weight_decay"] == 0.1 and opt.param_groups[0]["betas"] == (0.9, 0.999)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 0.5)
    assert opt.param_groups[0]["lr"] == pytest.approx(5e-4) and sched is not None
    w.grad = torch.ones(4, 4)
    with pytest.raises(lib.SerError):
        opt.step()


def test_global_batch_loss_algebra():
    """per-shard raw sums + global class counts + B_global reproduce the single-process loss."""
    from oracle import fusion_head_oracle as O
    g = torch.Generator().manual_seed(0)
    B, C, W = 24, 6, 3
    logits = torch.randn(B, C, generator=g) * 3
    unc = torch.rand(B, 1, generator=g)
    labels = torch.randint(0, C, (B,), generator=g)
    full_focal = O.class_balanced_focal(logits, labels, num_classes=C)
    full_ce = O.label_smoothing_ce(logits, labels)
    correct = (labels == logits.argmax(1)).float()
    full_unc = (unc * correct).mean()
    counts = torch.bincount(labels, minlength=C)
    ce_sum = focal_sum = unc_sum = corr_sum = 0.0
    for r in range(W):
        sl = slice(r * B // W, (r + 1) * B // W)
        n = sl.stop - sl.start
        ce_sum += float(O.label_smoothing_ce(logits[sl], labels[sl])) * n
        focal_sum += float(O.class_balanced_focal(logits[sl], labels[sl], num_classes=C, counts=counts)) * n
        unc_sum += float(unc[sl].sum()); corr_sum += float(correct[sl].sum())
    assert abs(ce_sum / B - float(full_ce)) < 1e-5
    assert abs(focal_sum / B - float(full_focal)) < 1e-5
    assert abs((unc_sum / B) * (corr_sum / B) - float(full_unc)) < 1e-6


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from mmser_b200.parallel import GradBucketReducer, global_loss_cfg
    red = GradBucketReducer()
    bufs = {n: torch.full((1000,), float(rank + 1) * (i + 1)) for i, n in enumerate(["classifier", "fusion", "cross"])}
    for n, b in bufs.items():                 # backward order: classifier first
        red.reduce_async(n, b)
    red.finish()
    # the "without overlap" arm: nothing is launched before finish(), the result is the same
    red2 = GradBucketReducer()
    red2.overlap = False
    late = {n: torch.full((2_000_000,), float(rank + 1)) for n in ("big",)}
    for n, b in late.items():
        red2.reduce_async(n, b)
    assert not red2.pending and len(red2.deferred) == 1
    red2.finish()
    assert float(late["big"][0]) == 3.0 and float(late["big"][-1]) == 3.0
    labels = torch.tensor([0, 1, 1, 3]) if rank == 0 else torch.tensor([2, 2, 1, 0])
    cfg = global_loss_cfg(labels, 4)
    cfg["before_loss"]()                      # the count all-reduce is asynchronous: the loss awaits it, so do we
    sums = torch.tensor([1.0, 2.0]) * (rank + 1)
    cfg["all_reduce"](sums)
    q.put((rank, red.order, {n: float(b[0]) for n, b in bufs.items()}, cfg["counts"].tolist(), cfg["B_global"], sums.tolist()))
    dist.destroy_process_group()


def test_data_parallel_reduction_world2_gloo(lib):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, order, vals, counts, bg, sums in res:
        assert order == ["classifier", "fusion", "cross"]
        assert vals == {"classifier": 3.0, "fusion": 6.0, "cross": 9.0}      # SUM over ranks
        assert counts == [2.0, 3.0, 2.0, 1.0] and bg == 8
        assert sums == [3.0, 6.0]


def test_pack_frames_host_side():
    """functional.pack_frames (host half of ser_unpack_frames): valid frames in batch order, B + 1 offsets, and a clear
    error for masks that are not right-padded."""
    sys.path.insert(0, ROOT)
    from mmser_b200.functional import pack_frames
    x = torch.arange(2 * 4 * 3, dtype=torch.float32).reshape(2, 4, 3)
    mask = torch.tensor([[1., 1., 0., 0.], [1., 1., 1., 0.]])
    packed, off = pack_frames(x * mask[:, :, None], mask)
    assert off.tolist() == [0, 2, 5] and packed.shape == (5, 3)
    assert torch.equal(packed, torch.cat([x[0, :2], x[1, :3]]))
    with pytest.raises(ValueError):
        pack_frames(x, torch.tensor([[1., 0., 1., 0.], [1., 1., 1., 0.]]))
