"""Host-side training loop pieces next to the hot path: the learning-rate and
regulariser-weight schedules of experiments/src/utils.jl:1-68 / experiments/src/construct.jl:78-152, the CSV
metric columns of experiments/src/logging.jl:102-128 and one MNIST neural-ODE training iteration
(experiments/src/utils.jl:106-115) built from the C-ABI calls (layer forward, head + cross-entropy, adjoint,
Adam).  Scalars and bookkeeping only: every array operation runs in libLRNDE.so."""
from __future__ import annotations

import csv
import ctypes as C
import math
import time
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check, lib
from .layers import (AugmenterLayer, Chain, Context, Conv, ConvChain, Dense, NeuralODE, TDChain, TDConvChain, _whcn,
                     batchnorm_backward, batchnorm_forward, conv2d_backward, conv2d_forward, glorot_uniform, nparams)


# ------------------------------------------------------------------ schedules (experiments/src/utils.jl:1-68)
class ExponentialDecay:
    def __init__(self, l0: float, l1: float, nsteps: int):
        self.l0, self.l1, self.nsteps = float(l0), float(l1), int(nsteps)
        self.k = math.log(self.l0 / self.l1) / self.nsteps

    def __call__(self, t):
        return self.l0 * math.exp(-self.k * t)


class InverseDecay:
    def __init__(self, l0: float, gamma: float):
        self.l0, self.gamma = float(l0), float(gamma)

    def __call__(self, t):
        return self.l0 / (1 + self.gamma * t)


class Step:
    def __init__(self, l0: float, gamma: float, step_sizes):
        self.l0, self.gamma = float(l0), float(gamma)
        self.step_sizes = [step_sizes] if isinstance(step_sizes, int) else list(step_sizes)

    def __call__(self, t):
        # searchsortedfirst(step_sizes, t - 1) - 1 (1-based index of the first entry >= t - 1, minus one)
        idx = sum(1 for s in self.step_sizes if s < t - 1)
        return self.l0 * self.gamma ** idx


class Constant:
    def __init__(self, lam: float):
        self.lam = float(lam)

    def __call__(self, t):
        return self.lam


class CosineAnneal:
    def __init__(self, l0: float, l1: float, period: int, restart: bool = False, dampen: float = 1.0):
        self.range, self.offset = abs(l0 - l1), min(l0, l1)
        self.dampen, self.period, self.restart = float(dampen), int(period), bool(restart)

    def __call__(self, t):
        if self.restart:
            d = self.dampen ** ((t - 1) // self.period)
            return (self.range * (1 + math.cos(math.pi * ((t - 1) % self.period) / self.period)) / 2 + self.offset) / d
        return self.range * (1 + math.cos(math.pi * (t - 1) / self.period)) / 2 + self.offset


def w_kl_schedule(t):
    """construct.jl:97: t -> max(0, 1 - 0.99f0^(t - 100))."""
    return max(0.0, 1.0 - 0.99 ** (t - 100))


def lr_scheduler(name: str, lr: float, *, total_steps: int = 1, cosine_lr_div_factor=100.0, cosine_cycle_length=50000,
                 cosine_dampen=1.0, step_lr_step_decay=0.1, step_lr_steps=(1,), inverse_decay_factor=1e-4,
                 exponential_lr_div_factor=100.0):
    """construct.jl:127-148."""
    if name == "cosine":
        return CosineAnneal(lr, lr / cosine_lr_div_factor, cosine_cycle_length, restart=True, dampen=cosine_dampen)
    if name == "constant":
        return Constant(lr)
    if name == "step":
        return Step(lr, step_lr_step_decay, step_lr_steps)
    if name == "inverse":
        return InverseDecay(lr, inverse_decay_factor)
    if name == "exponential":
        return ExponentialDecay(lr, lr / exponential_lr_div_factor, total_steps)
    raise ValueError("unknown value for `scheduler` = %s. Supported options are: `constant`, `step`, "
                     "`exponential`, `inverse` and `cosine`." % name)


# ------------------------------------------------------------------ optimisers (experiments/src/construct.jl:103-126)
class Optimiser:
    """``construct(expt, cfg::OptimizerConfig)``: cfg.optimizer in ("adam", "adamw", "adamax", "sgd"); "sgd" is
    Nesterov(lr, momentum) when cfg.nesterov, Descent(lr) when momentum == 0, else Momentum(lr, momentum);
    weight_decay != 0 chains WeightDecay(weight_decay) after the rule (Optimisers.jl OptimiserChain).  AdamW(lr) is
    Adam chained with WeightDecay(0) in Optimisers.jl, i.e. the Adam update.  The state lives in device buffers
    (torch tensors) and every update is one ``lrnde_opt_step`` launch per parameter block."""

    def __init__(self, optimizer: str = "adam", learning_rate: float = 1e-3, momentum: float = 0.0,
                 nesterov: bool = False, weight_decay: float = 0.0, beta=(0.9, 0.999), eps: float = 1e-8):
        if optimizer in ("adam", "adamw"):
            self.kind = "adam"
        elif optimizer == "adamax":
            self.kind = "adamax"
        elif optimizer == "sgd":
            self.kind = "nesterov" if nesterov else ("descent" if momentum == 0 else "momentum")
        else:
            raise ValueError(f"unknown value for `optimizer` = {optimizer}. Supported options are: `adam`, `adamax` "
                             "and `sgd`.")
        self.lr, self.momentum, self.weight_decay, self.beta, self.eps = learning_rate, momentum, weight_decay, beta, eps
        self.state = {}

    def update(self, ctx: Context, key, p, g, step: int, lr: Optional[float] = None):
        """p <- p - rule(g) in place (p, g: CUDA tensors); ``step`` is the 1-based iteration count."""
        import torch
        if key not in self.state:
            n1 = self.kind != "descent"
            n2 = self.kind in ("adam", "adamax")
            self.state[key] = (torch.zeros_like(p) if n1 else None, torch.zeros_like(p) if n2 else None)
        s1, s2 = self.state[key]
        a, b = (self.beta if self.kind in ("adam", "adamax") else (self.momentum, 0.0))
        g = g.contiguous()
        check(lib().lrnde_opt_step(ctx._h, _lib.OPT[self.kind], p.data_ptr(), g.data_ptr(),
                                   s1.data_ptr() if s1 is not None else None, s2.data_ptr() if s2 is not None else None,
                                   p.numel(), float(self.lr if lr is None else lr), float(a), float(b), float(self.eps),
                                   float(self.weight_decay), int(step)))


# ------------------------------------------------------------------ CSV schema (experiments/src/logging.jl:102-128)
def train_csv_header(latent_ode: bool = False, sde: bool = False) -> List[str]:
    return (["Step", "Batch Time", "Data Time", "Forward Pass Time", "Backward Pass Time", "Optimizer Time"]
            + (["Neg Log Likelihood", "KL Divergence"] if latent_ode else ["Cross Entropy Loss"])
            + ["Regularize Value", "Net Loss"] + (["NFE Drift", "NFE Diffusion"] if sde else ["NFE"])
            + ([] if latent_ode else ["Accuracy (Top 1)", "Accuracy (Top 5)"]))


def eval_csv_header(latent_ode: bool = False, sde: bool = False) -> List[str]:
    return (["Step", "Batch Time", "Data Time", "Forward Pass Time"] + ([] if latent_ode else ["Cross Entropy Loss"])
            + ["Regularize Value", "Net Loss"] + (["NFE Drift", "NFE Diffusion"] if sde else ["NFE"])
            + ([] if latent_ode else ["Accuracy (Top 1)", "Accuracy (Top 5)"]))


class CSVLogger:
    def __init__(self, path: str, header: Sequence[str]):
        self.path, self.header = path, list(header)
        with open(path, "w", newline="") as f:
            csv.writer(f).writerow(self.header)

    def __call__(self, *row):
        if len(row) != len(self.header):
            raise ValueError(f"expected {len(self.header)} columns, got {len(row)}")
        with open(self.path, "a", newline="") as f:
            csv.writer(f).writerow(row)


# ------------------------------------------------------------------ MNIST neural-ODE trainer (torch tensors = device memory)
class MnistODETrainer:
    """FlattenLayer -> NeuralODE(TDChain(Dense(785 => H, tanh), Dense(H+1 => 784))) -> sol.u[end] ->
    Dense(784 => 10) with logitcrossentropy + w_reg * reg_val (experiments/src/construct.jl:180-200, :19-33),
    Adam (construct.jl:104), scheduled learning rate and w_reg."""

    def __init__(self, *, hidden: int = 100, lr: float = 1e-3, scheduler: str = "inverse", w_reg_start: float = 100.0,
                 w_reg_end: float = 10.0, w_reg_decay: str = "exponential", total_steps: int = 1000, abstol: float = 1.4e-8,
                 reltol: float = 1.4e-8, regularize: str = "unbiased", seed: int = 0, ctx: Optional[Context] = None,
                 device: int = 0, optimizer: str = "adam", momentum: float = 0.0, nesterov: bool = False,
                 weight_decay: float = 0.0):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.ctx = ctx or Context(device, torch.cuda.current_stream(self.dev).cuda_stream)
        self.optim = Optimiser(optimizer, lr, momentum, nesterov, weight_decay)
        self.D, self.C = 784, 10
        self.chain = TDChain(Chain(Dense(self.D, hidden, "tanh"), Dense(hidden, self.D)))
        self.node = NeuralODE(self.chain, regularize=regularize, save_start=False, abstol=abstol, reltol=reltol,
                              maxiters=10_000, ctx=self.ctx)
        rng = np.random.default_rng(seed)
        self.ps = torch.from_numpy(self.node.initialparameters(rng)).to(self.dev)
        a = math.sqrt(6.0 / (self.D + self.C))
        wc = np.concatenate([rng.uniform(-a, a, (self.C, self.D)).astype(np.float32).ravel(order="F"),
                             np.zeros(self.C, np.float32)])
        self.wc = torch.from_numpy(wc).to(self.dev)
        self.st = self.node.initialstates(rng)
        self.lr_sched = lr_scheduler(scheduler, lr, total_steps=total_steps)
        self.w_reg_sched = ExponentialDecay(w_reg_start, w_reg_end, total_steps) if w_reg_decay == "exponential" \
            else Constant(w_reg_start)
        self.step = 0

    def train_step(self, x, y, data_time: float = 0.0):
        """x: (B, 784) float32 CUDA tensor, y: (B,) int32 labels.  Returns the CSV row of this step."""
        torch, L = self.torch, lib()
        t_batch = time.perf_counter()
        self.step += 1
        w_reg, lr = self.w_reg_sched(self.step), self.lr_sched(self.step)
        B = x.shape[0]
        t0 = time.perf_counter()
        sol, st2 = self.node(x.t(), self.ps, self.st)
        u_last = sol.u[-1].t()
        loss = C.c_float()
        d_u = torch.empty_like(u_last)
        d_wc = torch.empty_like(self.wc)
        check(L.lrnde_head_ce(self.ctx._h, self.wc.data_ptr(), u_last.data_ptr(), y.data_ptr(), B, self.D, self.C, 0,
                              C.byref(loss), d_u.data_ptr(), d_wc.data_ptr()))
        torch.cuda.synchronize(self.dev)
        t_fwd = time.perf_counter() - t0
        t0 = time.perf_counter()
        d_x, d_ps = self.node.backward(sol, [None] * (len(sol.u) - 1) + [d_u.t()], w_reg)
        torch.cuda.synchronize(self.dev)
        t_bwd = time.perf_counter() - t0
        t0 = time.perf_counter()
        for key, p, g in (("ps", self.ps, d_ps), ("wc", self.wc, d_wc)):
            self.optim.update(self.ctx, key, p, g, self.step, lr)
        torch.cuda.synchronize(self.dev)
        t_opt = time.perf_counter() - t0
        # metrics (not on the timed path): logits = W u + b
        W = self.wc[: self.C * self.D].view(self.D, self.C)       # column-major [C x D]
        logits = u_last @ W + self.wc[self.C * self.D:]
        top5 = logits.topk(5, dim=1).indices
        yl = y.long()
        acc1 = 100.0 * float((top5[:, 0] == yl).float().mean())        # percent, as accuracy() of experiments/src/utils.jl
        acc5 = 100.0 * float((top5 == yl[:, None]).any(dim=1).float().mean())
        reg = float(st2["reg_val"])
        nfe = int(st2["nfe"])
        sol.free()
        self.st = st2
        return [self.step, time.perf_counter() - t_batch, data_time, t_fwd, t_bwd, t_opt, float(loss.value), reg,
                float(loss.value) + w_reg * reg, nfe, acc1, acc5]


# ------------------------------------------------------------------ CIFAR10 conv neural-ODE trainer
class Cifar10ODETrainer:
    """AugmenterLayer(Conv 3=>5) -> BatchNorm(8) -> NeuralODE(TDChain(Conv 9=>h + BN gelu, Conv h+1=>h + BN gelu,
    Conv h+1=>8)) -> sol.u[end] -> Conv(8=>1, gelu) -> Flatten -> Dense(W*H => 10) with logitcrossentropy +
    w_reg * reg_val (experiments/src/construct.jl:212-227, :19-33), Adam, scheduled learning rate and w_reg.
    Images are (B, 3, H, W) CUDA tensors: that memory IS the reference's column-major (W, H, 3, B) array."""

    def __init__(self, *, hidden: int = 64, image: int = 32, lr: float = 1e-3, scheduler: str = "inverse",
                 w_reg_start: float = 100.0, w_reg_end: float = 10.0, total_steps: int = 1000, abstol: float = 1e-4,
                 reltol: float = 1e-4, regularize: str = "unbiased", seed: int = 0, ctx: Optional[Context] = None,
                 device: int = 0, optimizer: str = "adam", momentum: float = 0.0, nesterov: bool = False,
                 weight_decay: float = 0.0):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.ctx = ctx or Context(device, torch.cuda.current_stream(self.dev).cuda_stream)
        self.optim = Optimiser(optimizer, lr, momentum, nesterov, weight_decay)
        self.W = self.H = int(image)
        self.C = 10
        self.aug = AugmenterLayer(3, 5)
        self.chain = TDConvChain(ConvChain(Conv(8, hidden, True, "gelu"), Conv(hidden, hidden, True, "gelu"),
                                           Conv(hidden, 8), width=self.W, height=self.H))
        self.node = NeuralODE(self.chain, regularize=regularize, save_start=False, abstol=abstol, reltol=reltol,
                              maxiters=10_000, ctx=self.ctx)
        rng = np.random.default_rng(seed)

        def conv_init(cin, cout):          # Lux Conv: Glorot-uniform weight, zero bias
            a = math.sqrt(6.0 / (9 * cin + 9 * cout))
            return np.concatenate([rng.uniform(-a, a, 9 * cin * cout), np.zeros(cout)]).astype(np.float32)

        D = self.W * self.H
        a = math.sqrt(6.0 / (D + self.C))
        params = dict(augment=conv_init(3, 5), bn=np.concatenate([np.ones(8), np.zeros(8)]).astype(np.float32),
                      node=glorot_uniform(self.chain, rng), cls_conv=conv_init(8, 1),
                      head=np.concatenate([rng.uniform(-a, a, self.C * D), np.zeros(self.C)]).astype(np.float32))
        self.ps = {k: torch.from_numpy(v).to(self.dev) for k, v in params.items()}
        self.st = self.node.initialstates(rng)
        self.bn_state = dict(running=torch.cat([torch.zeros(8), torch.ones(8)]).to(self.dev), training=True)
        self.lr_sched = lr_scheduler(scheduler, lr, total_steps=total_steps)
        self.w_reg_sched = ExponentialDecay(w_reg_start, w_reg_end, total_steps)
        self.step = 0

    def train_step(self, x, y, data_time: float = 0.0):
        """x: (B, 3, H, W) float32 CUDA tensor, y: (B,) int32 labels.  Returns the CSV row of this step."""
        torch, L = self.torch, lib()
        t_batch = time.perf_counter()
        self.step += 1
        w_reg, lr = self.w_reg_sched(self.step), self.lr_sched(self.step)
        B, W, H, ps = x.shape[0], self.W, self.H, self.ps
        t0 = time.perf_counter()
        xl = x.permute(3, 2, 1, 0)                                       # logical (W, H, 3, B)
        a0 = self.aug(xl, ps["augment"], self.ctx)
        a1, bn_state = batchnorm_forward(a0, ps["bn"], self.bn_state, "identity", self.ctx)
        sol, st2 = self.node(_whcn(a1).reshape(B, -1).t(), ps["node"], self.st)
        u2 = sol.u[-1].t().reshape(B, 8, H, W).permute(3, 2, 1, 0)
        c = conv2d_forward(u2, ps["cls_conv"], 1, "gelu", True, self.ctx)
        flat = _whcn(c).reshape(B, W * H)                                # FlattenLayer: column-major [W*H, B]
        loss = C.c_float()
        d_flat = torch.empty_like(flat)
        grads = {"head": torch.empty_like(ps["head"])}
        check(L.lrnde_head_ce(self.ctx._h, ps["head"].data_ptr(), flat.data_ptr(), y.data_ptr(), B, W * H, self.C, 0,
                              C.byref(loss), d_flat.data_ptr(), grads["head"].data_ptr()))
        torch.cuda.synchronize(self.dev)
        t_fwd = time.perf_counter() - t0
        t0 = time.perf_counter()
        d_u2, grads["cls_conv"] = conv2d_backward(u2, ps["cls_conv"], d_flat.reshape(B, 1, H, W).permute(3, 2, 1, 0),
                                                  "gelu", True, True, self.ctx)
        cot = _whcn(d_u2).reshape(B, -1).t()
        d_a1, grads["node"] = self.node.backward(sol, [None] * (len(sol.u) - 1) + [cot], w_reg)
        d_a0, grads["bn"] = batchnorm_backward(a0, ps["bn"], d_a1.t().reshape(B, 8, H, W).permute(3, 2, 1, 0),
                                               dict(running=None, training=True), "identity", self.ctx)
        _, grads["augment"] = self.aug.backward(xl, ps["augment"], d_a0, self.ctx)
        torch.cuda.synchronize(self.dev)
        t_bwd = time.perf_counter() - t0
        t0 = time.perf_counter()
        for k, p in ps.items():
            self.optim.update(self.ctx, k, p, grads[k], self.step, lr)
        torch.cuda.synchronize(self.dev)
        t_opt = time.perf_counter() - t0
        Wh = ps["head"][: self.C * W * H].view(W * H, self.C)             # column-major [C x D]
        logits = flat @ Wh + ps["head"][self.C * W * H:]
        top5 = logits.topk(5, dim=1).indices
        yl = y.long()
        acc1 = 100.0 * float((top5[:, 0] == yl).float().mean())        # percent, as accuracy() of experiments/src/utils.jl
        acc5 = 100.0 * float((top5 == yl[:, None]).any(dim=1).float().mean())
        reg, nfe = float(st2["reg_val"]), int(st2["nfe"])
        sol.free()
        self.st, self.bn_state = st2, bn_state
        return [self.step, time.perf_counter() - t_batch, data_time, t_fwd, t_bwd, t_opt, float(loss.value), reg,
                float(loss.value) + w_reg * reg, nfe, acc1, acc5]
