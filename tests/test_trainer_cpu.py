"""Host logic of the network-level training step (uni_renderer_b200/trainer.py) on the CPU emulation of the ops: the tape
(op order, gradient accumulation at fan-outs, the exchange between the three networks, parameter naming, loss scaling,
clipping, optimizer bookkeeping) against torch autograd of the oracle.  The kernels themselves are checked on the GPU by
tests/test_trainer_gpu.py; here every op is fp32 torch math on fp16-stored tensors, so the gate is tight."""
import torch

from tests import cpu_ops_emulator as emu
from tests.test_trainer_gpu import _oracle_grads, _rel, _setup


def test_tape_gradients_match_autograd_of_the_oracle_on_the_emulator(monkeypatch):
    from uni_renderer_b200.trainer import DualStreamTrainer
    emu.install_training(monkeypatch)
    nets, cfgs, batch = _setup(S=16, Lc=77)
    loss_ref, img_ref, msk_ref, gref = _oracle_grads(nets, cfgs, batch, torch.device("cpu"))
    tr = DualStreamTrainer(nets, cfgs, loss_scale=256.0, max_grad_norm=None, device="cpu")
    loss, img, msk = tr.forward_backward(batch["x_img"], batch["t_img"], batch["x_attr"], batch["t_attr"], batch["ehs"],
                                         batch["img_target"], batch["attr_target"])
    assert _rel(img, img_ref) <= 3e-3 and _rel(msk, msk_ref) <= 3e-3
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
    errs = sorted(((_rel(tr.P.g[n] / tr.loss_scale, g), n) for n, g in gref.items()), reverse=True)
    assert set(gref) == set(tr.P.g)
    assert errs[0][0] <= 1.5e-2, errs[:8]
    flat_ref = torch.cat([gref[n].reshape(-1) for n in tr.P.g])
    assert _rel(tr.P.grad / tr.loss_scale, flat_ref) <= 4e-3


def test_consistency_pass_on_the_emulator(monkeypatch):
    from uni_renderer_b200.trainer import DualStreamTrainer
    emu.install_training(monkeypatch)
    nets, cfgs, batch = _setup(seed=21, S=16, Lc=16)
    g = torch.Generator().manual_seed(77)
    cycle = (torch.randn(2, 4, 16, 16, generator=g).half().float(), torch.tensor([801.0, 121.0]))
    loss_ref, _, _, gref = _oracle_grads(nets, cfgs, batch, torch.device("cpu"), cycle=cycle)
    tr = DualStreamTrainer(nets, cfgs, loss_scale=256.0, max_grad_norm=None, device="cpu")
    loss, _, _ = tr.forward_backward(batch["x_img"], batch["t_img"], batch["x_attr"], batch["t_attr"], batch["ehs"],
                                     batch["img_target"], batch["attr_target"], cycle=cycle)
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
    errs = sorted(((_rel(tr.P.g[n] / tr.loss_scale, gr), n) for n, gr in gref.items()), reverse=True)
    assert errs[0][0] <= 2e-2, errs[:8]


def test_gradient_checkpointing_gives_the_same_gradients_on_the_emulator(monkeypatch):
    from uni_renderer_b200.trainer import DualStreamTrainer
    emu.install_training(monkeypatch)
    nets, cfgs, batch = _setup(seed=4, S=16, Lc=16)
    args = (batch["x_img"], batch["t_img"], batch["x_attr"], batch["t_attr"], batch["ehs"], batch["img_target"],
            batch["attr_target"])
    grads = []
    for ck in (False, True):
        tr = DualStreamTrainer(nets, cfgs, loss_scale=256.0, gradient_checkpointing=ck, device="cpu")
        tr.forward_backward(*args)
        grads.append(tr.P.grad.clone())
    assert _rel(grads[1], grads[0]) <= 1e-3, _rel(grads[1], grads[0])       # regrouped fp16 gradient sums only


def test_optimizer_step_bookkeeping_on_the_emulator(monkeypatch):
    from uni_renderer_b200.trainer import DualStreamTrainer
    emu.install_training(monkeypatch)
    nets, cfgs, batch = _setup(seed=9, S=16, Lc=16)
    tr = DualStreamTrainer(nets, cfgs, lr=2e-4, loss_scale=256.0, max_grad_norm=1.0, device="cpu")
    args = (batch["x_img"], batch["t_img"], batch["x_attr"], batch["t_attr"], batch["ehs"], batch["img_target"],
            batch["attr_target"])
    l0, _, _ = tr.forward_backward(*args)
    p_ref = tr.P.flat.clone().requires_grad_(True)
    g = tr.P.grad.clone() / tr.loss_scale
    gn = g.norm().item()
    if gn > 1.0:
        g = g * (1.0 / (gn + 1e-6))
    opt = torch.optim.AdamW([p_ref], lr=2e-4, weight_decay=1e-2)
    p_ref.grad = g
    opt.step()
    info = tr.optimizer_step()
    assert info["skipped"] == 0.0 and (tr.P.flat - p_ref.detach()).abs().max().item() <= 1e-6
    assert float(tr.P.grad.abs().max()) == 0.0 and not tr.P._cache
    # an overflowing gradient skips the update and leaves the parameters alone
    tr.P.grad[0] = float("inf")
    before = tr.P.flat.clone()
    assert tr.optimizer_step()["skipped"] == 1.0 and torch.equal(before, tr.P.flat)
    losses = [l0.item()] + [tr.step(*args)["loss"] for _ in range(2)]
    assert losses[-1] < losses[0], losses
