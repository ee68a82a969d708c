"""oracle/train_oracle.py (CPU restatement of the reference's loss + BertAdam step) against the fixtures produced by
the reference's own utils/losses.py / optimization.py (oracle/gen_golden.py train)."""
import torch

from oracle import train_oracle as T
from tests.helpers import load_golden


def test_seg_loss_matches_reference():
    fx = load_golden("train_loss_tiny")
    x = fx["logits"].clone().requires_grad_(True)
    loss, ce, dice = T.seg_loss(x, fx["mask"], fx["pos_weight"], fx["class_weights"], fx["dice_w"])
    assert abs(float(loss) - float(fx["loss"])) < 1e-6
    assert abs(float(ce) - float(fx["ce"])) < 1e-6 and abs(float(dice) - float(fx["dice"])) < 1e-6
    (g,) = torch.autograd.grad(loss, [x])
    assert float((g - fx["dlogits"]).abs().max()) < 1e-7 + 1e-5 * float(fx["dlogits"].abs().max())


def test_bert_adam_matches_reference():
    fx = load_golden("train_bertadam_tiny")
    params = [p.clone() for p in fx["init"]]
    state = {}
    for step, gs in enumerate(fx["grads"]):
        gs = [None if g is None else g.clone() for g in gs]
        T.clip_grad_norm([g for g in gs if g is not None], fx["grad_clip"])
        T.bert_adam_step(params, gs, state, lr=fx["lr"], weight_decay=fx["weight_decay"], warmup=fx["warmup"],
                         t_total=fx["t_total"], max_grad_norm=fx["max_grad_norm"])
        for p, want in zip(params, fx["after"][step]):
            assert float((p - want).abs().max()) < 1e-6 * max(1.0, float(want.abs().max())), step
    assert torch.equal(params[4], fx["init"][4])          # the never-used parameter is untouched (grad None)
