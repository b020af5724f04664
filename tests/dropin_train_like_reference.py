"""A training run written the way the reference's src/train.py is, executed against the drop-in head.

Run with the import-path shim in front (tests/test_dropin_train_loop.py does):
    PYTHONPATH=<repo>/dropin:<repo> python tests/dropin_train_like_reference.py [--use_amp] [--fused_optimizer]

What is kept from the reference, statement by statement: the `from models...` imports (src/train.py:4-9), the module
construction with its keyword arguments (:54-69), the ten AdamW parameter groups that reach into
classifier.deep_classifier / .anchor_clustering / .uncertainty_head (:72-83), the loss modules (:84-86), GradScaler (:88),
the warm-up-cosine LambdaLR (:114-121), the six-call forward and the loss composition under autocast (:145-168), the
AMP / non-AMP backward + step (:169-177), the validation forward with OpenMax on (:181-201) and the after-last-epoch
Weibull pass that walks the classifier's children by hand (:204-245).  What is replaced: the dataset and the two frozen
HuggingFace encoders (not available offline, out of scope) by stub encoders that emit synthetic hidden states and own the
head's bottleneck adapters, exactly where the reference's encoders own theirs (audio_encoder.py:19-21,112).
"""
import argparse
import math
import sys

import torch
import torch.nn as nn
import torch.optim as optim
from torch.amp import GradScaler, autocast

from models import FusionLayer
from models.classifier import AdvancedOpenMaxClassifier
from models.cross_attention import CrossModalAttention
from models.pooling import AttentiveStatsPooling
from models.losses import LabelSmoothingCrossEntropy, ClassBalancedFocalLoss, SupConLoss
from models.prototypes import PrototypeMemory
from models.adapter import BottleneckAdapter

NUM_LABELS = 4


class _Cfg:
    hidden_size = 768


class _Frozen(nn.Module):
    config = _Cfg()


class StubEncoder(nn.Module):
    """Stands in for AudioEncoder / TextEncoder: `encoder.config.hidden_size`, a trainable `adapter`, and a forward that
    returns (sequence + adapter(sequence), float mask) for a batch of pre-computed hidden states."""

    def __init__(self):
        super().__init__()
        self.encoder = _Frozen()
        self.adapter = BottleneckAdapter(768, 256)

    def forward(self, batch):
        seq, mask = batch
        return seq + self.adapter(seq), mask


def make_loader(n_batches, B, Ta, Tt, device, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        la = torch.randint(Ta // 2, Ta + 1, (B,), generator=g)
        lt = torch.randint(max(1, Tt // 4), Tt + 1, (B,), generator=g)
        am = (torch.arange(Ta)[None] < la[:, None]).float()
        tm = (torch.arange(Tt)[None] < lt[:, None]).float()
        labels = torch.randint(0, NUM_LABELS, (B,), generator=g)
        # class-dependent mean so that there is something to learn
        a = (torch.randn(B, Ta, 768, generator=g) + 0.5 * labels[:, None, None].float() / NUM_LABELS) * am[..., None]
        t = torch.randn(B, Tt, 768, generator=g) * tm[..., None]
        out.append(((a.to(device), am.to(device)), (t.to(device), tm.to(device)), labels))
    return out


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--epochs", type=int, default=2)
    p.add_argument("--lr", type=float, default=2e-4)
    p.add_argument("--warmup_ratio", type=float, default=0.1)
    p.add_argument("--proto_weight", type=float, default=1.0)
    p.add_argument("--use_amp", action="store_true")
    p.add_argument("--fused_optimizer", action="store_true")
    args = p.parse_args()
    device = torch.device("cuda")
    torch.manual_seed(0)

    audio_encoder = StubEncoder().to(device)
    text_encoder = StubEncoder().to(device)
    audio_hid = audio_encoder.encoder.config.hidden_size
    text_hid = text_encoder.encoder.config.hidden_size
    cross = CrossModalAttention(audio_hid, text_hid, shared_dim=256, num_heads=8).to(device)
    pool_a = AttentiveStatsPooling(audio_hid).to(device)
    pool_t = AttentiveStatsPooling(text_hid).to(device)
    fusion = FusionLayer(audio_hid * 2, text_hid * 2, 512).to(device)
    classifier = AdvancedOpenMaxClassifier(input_dim=512, num_labels=NUM_LABELS, num_layers=35, base_dim=512,
                                           dropout=0.15).to(device)
    prototypes = PrototypeMemory(NUM_LABELS, 512).to(device)

    groups = [
        {"params": audio_encoder.parameters(), "lr": args.lr * 0.1, "weight_decay": 0.025},
        {"params": text_encoder.parameters(), "lr": args.lr * 0.1, "weight_decay": 0.025},
        {"params": cross.parameters(), "lr": args.lr, "weight_decay": 0.05},
        {"params": pool_a.parameters(), "lr": args.lr, "weight_decay": 0.05},
        {"params": pool_t.parameters(), "lr": args.lr, "weight_decay": 0.05},
        {"params": fusion.parameters(), "lr": args.lr, "weight_decay": 0.05},
        {"params": classifier.deep_classifier.parameters(), "lr": args.lr * 1.5, "weight_decay": 0.06},
        {"params": classifier.anchor_clustering.parameters(), "lr": args.lr * 2.0, "weight_decay": 0.04},
        {"params": classifier.uncertainty_head.parameters(), "lr": args.lr * 1.0, "weight_decay": 0.05},
        {"params": prototypes.parameters(), "lr": args.lr, "weight_decay": 0.05},
    ]
    if args.fused_optimizer:
        from mmser_b200.optim import FusedAdamW          # same constructor as optim.AdamW
        optimizer = FusedAdamW(groups, weight_decay=0.05)
    else:
        optimizer = optim.AdamW(groups, weight_decay=0.05)
    ce_smooth = LabelSmoothingCrossEntropy(0.1)
    cb_focal = ClassBalancedFocalLoss(beta=0.9999, gamma=2.0, num_classes=NUM_LABELS)
    supcon = SupConLoss(temperature=0.07)   # noqa: F841  (constructed, never called -- as in the reference)
    scaler = GradScaler(enabled=args.use_amp)

    train_loader = make_loader(6, 16, 40, 12, device, 1)
    val_loader = make_loader(2, 16, 40, 12, device, 2)
    total_steps = len(train_loader) * args.epochs
    warmup_steps = int(total_steps * args.warmup_ratio)

    def lr_lambda(step):
        if step < warmup_steps:
            return float(step) / max(1, warmup_steps)
        progress = (step - warmup_steps) / max(1, total_steps - warmup_steps)
        return 0.5 * (1.0 + math.cos(progress * math.pi))

    scheduler = optim.lr_scheduler.LambdaLR(optimizer, lr_lambda)
    losses = []
    for epoch in range(args.epochs):
        audio_encoder.train(); text_encoder.train(); fusion.train(); classifier.train()
        for audio_batch, text_batch, labels in train_loader:
            labels = labels.to(device)
            a_seq, a_mask = audio_encoder(audio_batch)
            t_seq, t_mask = text_encoder(text_batch)
            a_enh, t_enh = cross(a_seq, t_seq, a_mask, t_mask)
            a_vec = pool_a(a_enh, a_mask)
            t_vec = pool_t(t_enh, t_mask)
            fused = fusion(a_vec, t_vec)
            with autocast("cuda", enabled=args.use_amp):
                logits, uncertainty, anchor_loss = classifier(fused, use_openmax=False, return_uncertainty=True)
                ce_loss = ce_smooth(logits, labels)
                focal_loss = cb_focal(logits, labels)
                loss = ce_loss + 0.3 * focal_loss
                loss = loss + 0.1 * anchor_loss
                uncertainty_loss = torch.mean(uncertainty * (labels == logits.argmax(dim=1)).float())
                loss = loss + 0.05 * uncertainty_loss
                if args.proto_weight > 0:
                    proto_loss = prototypes.prototype_loss(fused, labels)
                    loss = loss + 0.01 * proto_loss
            optimizer.zero_grad(set_to_none=True)
            if args.use_amp:
                scaler.scale(loss).backward()
                scaler.step(optimizer)
                scaler.update()
            else:
                loss.backward()
                optimizer.step()
            scheduler.step()
            losses.append(float(loss))

        # validation: OpenMax on (classifier(fused) with the default use_openmax=True in eval mode)
        audio_encoder.eval(); text_encoder.eval(); fusion.eval(); classifier.eval()
        correct = total = 0
        with torch.no_grad():
            for audio_batch, text_batch, labels in val_loader:
                labels = labels.to(device)
                a_seq, a_mask = audio_encoder(audio_batch)
                t_seq, t_mask = text_encoder(text_batch)
                a_enh, t_enh = cross(a_seq, t_seq, a_mask, t_mask)
                fused = fusion(pool_a(a_enh, a_mask), pool_t(t_enh, t_mask))
                preds = classifier(fused).argmax(dim=1)
                correct += int((preds == labels).sum()); total += int(labels.numel())
        print(f"epoch {epoch}: mean loss {sum(losses[-len(train_loader):]) / len(train_loader):.4f}  val acc {correct / total:.3f}")

        if epoch == args.epochs - 1:
            # Weibull fitting pass: the classifier's children are called one by one, as src/train.py:204-245 does
            classifier.eval()
            all_features, all_val_labels = [], []
            with torch.no_grad():
                for audio_batch, text_batch, labels in val_loader:
                    labels = labels.to(device)
                    a_seq, a_mask = audio_encoder(audio_batch)
                    t_seq, t_mask = text_encoder(text_batch)
                    a_enh, t_enh = cross(a_seq, t_seq, a_mask, t_mask)
                    fused = fusion(pool_a(a_enh, a_mask), pool_t(t_enh, t_mask))
                    features = fused
                    for layer in classifier.deep_classifier.input_projection:
                        features = layer(features)
                    for residual_block, layer_norm in zip(classifier.deep_classifier.residual_layers,
                                                          classifier.deep_classifier.layer_norms):
                        features = layer_norm(features)
                        features = residual_block(features)
                    for i in range(4):
                        features = classifier.deep_classifier.output_projection[i](features)
                    all_features.append(features)
                    all_val_labels.append(labels)
            classifier.fit_weibull(torch.cat(all_features, dim=0), torch.cat(all_val_labels, dim=0))

    ckpt = {"cross": cross.state_dict(), "pool_a": pool_a.state_dict(), "pool_t": pool_t.state_dict(),
            "fusion": fusion.state_dict(), "classifier": classifier.state_dict(), "prototypes": prototypes.state_dict(),
            "optimizer": optimizer.state_dict(), "scheduler": scheduler.state_dict()}
    n_keys = sum(len(v) for k, v in ckpt.items() if k not in ("optimizer", "scheduler"))
    first, last = sum(losses[:3]) / 3, sum(losses[-3:]) / 3
    ok = all(math.isfinite(x) for x in losses) and last < first and float(classifier.weibull_beta.min()) > 0 \
        and float(classifier.activation_vectors.abs().max()) > 0
    print(f"RESULT ok={ok} first={first:.4f} last={last:.4f} ckpt_keys={n_keys} weibull_beta_min={float(classifier.weibull_beta.min()):.4f}")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
