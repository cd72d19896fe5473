"""CPU tests of the learning-rate schedules the reference's training scripts import from timm (restated in sunb200/optim.py;
timm is not installed here, so these pin the published formulas, not timm itself) and of make_optimizer's contract."""
import math

import torch

from sunb200.optim import CosineLRScheduler, MultiStepLRScheduler


def _opt(lr):
    return torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)


def test_cosine_schedule_offline_config():
    """offline.py:231: CosineLRScheduler(warmup_lr_init=1e-6, t_initial=800, cycle_decay=0.1, warmup_t=5)."""
    base = 5e-4
    opt = _opt(base)
    sch = CosineLRScheduler(opt, t_initial=800, cycle_decay=0.1, warmup_t=5, warmup_lr_init=1e-6)
    assert abs(opt.param_groups[0]["lr"] - 1e-6) < 1e-15                 # construction sets the warm-up start
    for t in range(5):
        sch.step(t)
        assert abs(opt.param_groups[0]["lr"] - (1e-6 + t * (base - 1e-6) / 5)) < 1e-12
    for t in (5, 100, 400, 799):
        sch.step(t)
        assert abs(opt.param_groups[0]["lr"] - 0.5 * base * (1 + math.cos(math.pi * t / 800))) < 1e-12
    sch.step(800)                                                        # cycle_limit 1: lr_min afterwards
    assert opt.param_groups[0]["lr"] == 0.0


def test_multistep_warmup_schedule():
    """train_meta_warmup.py:141: MultiStepLRScheduler(milestones, decay_rate=0.5, warmup_lr_init=1e-5, warmup_t=3)."""
    opt = _opt(1e-3)
    sch = MultiStepLRScheduler(opt, [10, 20], decay_rate=0.5, warmup_lr_init=1e-5, warmup_t=3)
    lrs = []
    for t in range(25):
        sch.step(t)
        lrs.append(opt.param_groups[0]["lr"])
    assert abs(lrs[0] - 1e-5) < 1e-12 and abs(lrs[2] - (1e-5 + 2 * (1e-3 - 1e-5) / 3)) < 1e-12
    assert lrs[3] == 1e-3 and lrs[8] == 1e-3
    assert lrs[9] == 5e-4 and lrs[18] == 5e-4 and lrs[19] == 2.5e-4      # decay counted at t + 1, as timm does


def test_make_optimizer_contract():
    import utils
    p = [torch.nn.Parameter(torch.zeros(3))]
    opt, sch = utils.make_optimizer(p, "sgd", lr=1e-3, weight_decay=5e-4, milestones=[2], gamma=0.5)
    g = opt.param_groups[0]
    assert g["lr"] == 1e-3 and g["momentum"] == 0.9 and g["weight_decay"] == 5e-4
    assert isinstance(opt, torch.optim.Optimizer) and sch is not None
    assert "param_groups" in opt.state_dict()                            # checkpoints store optimizer.state_dict()
    opt2, sch2 = utils.make_optimizer(p, "adamw", lr=1e-3, weight_decay=0.05)
    assert sch2 is None and opt2.param_groups[0]["betas"] == (0.9, 0.999)
