"""Rollout post-processing ops used by the MAPPO trainer, on the GPU.

Reference: `/root/reference/pacman_mappo_resnet.py`
  * ``compute_gae``            `:277-290`  -> :func:`compute_gae` (kernel pm_gae)
  * minibatch adv. normalise   `:577`      -> :func:`normalize_advantages` (kernel pm_adv_norm)
  * ``compute_heuristic_shaping`` `:251-264` is fused into the step kernel
    (``step_batched()['shaping']`` / ``['shaped_reward']``).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

GAMMA = 0.99        # pacman_mappo_resnet.py:19
GAE_LAMBDA = 0.95   # pacman_mappo_resnet.py:20


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def compute_gae(rewards, values, dones, last_value, gamma=GAMMA, gae_lambda=GAE_LAMBDA):
    """Reverse-scan GAE over M independent columns.

    rewards / values / dones: fp32 CUDA tensors [T] or [T, M] (any trailing shape is
    flattened to columns); last_value: Python float, or a tensor broadcastable to [M].
    Returns (advantages, returns) shaped like ``rewards``.  ``dones[t]`` gates the
    bootstrap at step t exactly as in the reference (quirk: not dones[t+1]).
    """
    L = _lib.load()
    if not rewards.is_cuda:
        raise RuntimeError("compute_gae expects CUDA tensors; there is no CPU fallback")
    dev = rewards.device
    shape = rewards.shape
    T = shape[0]
    r = rewards.reshape(T, -1).contiguous().float()
    v = values.reshape(T, -1).contiguous().float()
    d = dones.reshape(T, -1).contiguous().float()
    M = r.shape[1]
    lv = torch.as_tensor(last_value, dtype=torch.float32, device=dev).reshape(-1)
    lv = lv.expand(M).contiguous() if lv.numel() == 1 else lv.reshape(M).contiguous()
    adv = torch.empty_like(r)
    ret = torch.empty_like(r)
    with torch.cuda.device(dev):
        _lib.check(L.pm_gae(r.data_ptr(), v.data_ptr(), d.data_ptr(), lv.data_ptr(), float(gamma), float(gae_lambda),
                            adv.data_ptr(), ret.data_ptr(), int(T), int(M), _stream(dev)), "pm_gae")
    return adv.reshape(shape), ret.reshape(shape)


def normalize_advantages(adv_flat, inds, minibatch):
    """out[s:s+mb] = (a - a.mean()) / (a.std() + 1e-8) with a = adv_flat[inds[s:s+mb]]
    for every consecutive minibatch of ``inds`` (unbiased std, like torch.std)."""
    L = _lib.load()
    if not adv_flat.is_cuda:
        raise RuntimeError("normalize_advantages expects CUDA tensors; there is no CPU fallback")
    dev = adv_flat.device
    a = adv_flat.contiguous().float()
    idx = inds.to(device=dev, dtype=torch.int64).contiguous()
    out = torch.empty(idx.numel(), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.pm_adv_norm(a.data_ptr(), idx.data_ptr(), out.data_ptr(), int(idx.numel()), int(minibatch),
                                 _stream(dev)), "pm_adv_norm")
    return out


# ---- the PPO minibatch pipeline around the networks (pacman_mappo_resnet.py:555-600) -----------
def gather_rows(src_u8: torch.Tensor, idx: torch.Tensor, out_dtype=torch.float32, out: torch.Tensor | None = None):
    """``b_obs[mb].to(device)`` (`:569-573`) on the uint8 rollout buffer: ``src_u8`` [S, ...] uint8 (contiguous
    rows), ``idx`` int64 [n] -> [n, ...] in ``out_dtype`` (fp32 or bf16); the widening is fused into the gather."""
    L = _lib.load()
    if not src_u8.is_cuda or src_u8.dtype != torch.uint8 or not src_u8.is_contiguous():
        raise RuntimeError("gather_rows expects a contiguous uint8 CUDA tensor; there is no CPU fallback")
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("out_dtype must be float32 or bfloat16")
    dev = src_u8.device
    row = int(src_u8[0].numel())
    idx = idx.to(device=dev, dtype=torch.int64).contiguous()
    n = idx.numel()
    if out is None:
        out = torch.empty((n,) + tuple(src_u8.shape[1:]), dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.pm_gather_rows(src_u8.data_ptr(), idx.data_ptr(), n, row, out.data_ptr(),
                                    int(out_dtype == torch.bfloat16), _stream(dev)), "pm_gather_rows")
    return out


_scratch = {}


def _zeroed_scratch(kind: str, nbytes: int, dev):
    key = (kind, dev.index)
    if key not in _scratch:
        _scratch[key] = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    return _scratch[key]


class _PPOLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, values, actions, old_logp, adv, ret, clip_eps, vf_coef, ent_coef):
        L = _lib.load()
        dev = logits.device
        lg, v = logits.contiguous().float(), values.contiguous().float()
        B, A = lg.shape
        dl = torch.empty_like(lg)
        dv = torch.empty(B, dtype=torch.float32, device=dev)
        stats = torch.empty(8, dtype=torch.float32, device=dev)
        scratch = _zeroed_scratch("ppo", int(L.pm_ppo_scratch_bytes()), dev)
        with torch.cuda.device(dev):
            _lib.check(L.pm_ppo_loss(lg.data_ptr(), v.data_ptr(), actions.contiguous().data_ptr(),
                                     old_logp.contiguous().float().data_ptr(), adv.contiguous().float().data_ptr(),
                                     ret.contiguous().float().data_ptr(), B, v.numel(), A, float(clip_eps), float(vf_coef),
                                     float(ent_coef), dl.data_ptr(), dv.data_ptr(), stats.data_ptr(), scratch.data_ptr(),
                                     _stream(dev)), "pm_ppo_loss")
        ctx.save_for_backward(dl, dv)
        ctx.nv = v.numel()
        ctx.mark_non_differentiable(stats)
        return stats[0], stats

    @staticmethod
    def backward(ctx, g_loss, _g_stats):
        dl, dv = ctx.saved_tensors
        if ctx.nv != dv.numel():                        # both learners of an env-step share one critic value
            dv = dv.view(-1, ctx.nv).sum(0)
        return dl * g_loss, dv * g_loss, None, None, None, None, None, None, None


def ppo_loss(logits, values, actions, old_logp, norm_adv, returns, clip_eps, vf_coef, ent_coef):
    """The minibatch loss of `pacman_mappo_resnet.py:575-585` and its gradient in ONE kernel launch.
    ``logits`` [B, 5] (actor head output), ``values`` [B] or [B/2] (sample i reads ``values[i % len]``).  Returns
    (loss, stats) with stats[1:5] = pg loss, value loss, entropy mean, clip fraction (`:597-600`); ``loss`` is
    differentiable w.r.t. logits and values."""
    if not logits.is_cuda:
        raise RuntimeError("ppo_loss expects CUDA tensors; there is no CPU fallback")
    return _PPOLoss.apply(logits, values, actions, old_logp, norm_adv, returns, clip_eps, vf_coef, ent_coef)


class _Attention8(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, n_heads):
        L = _lib.load()
        dev = qkv.device
        x = qkv.contiguous().float()
        S, B, E3 = x.shape
        out = torch.empty((S, B, E3 // 3), dtype=torch.float32, device=dev)
        need_grad = qkv.requires_grad
        lse = torch.empty((B, n_heads, S), dtype=torch.float32, device=dev) if need_grad else None
        with torch.cuda.device(dev):
            _lib.check(L.pm_attn_fwd(x.data_ptr(), out.data_ptr(), lse.data_ptr() if need_grad else None, S, B, n_heads,
                                     _stream(dev)), "pm_attn_fwd")
        if need_grad:
            ctx.save_for_backward(x, out, lse)
        ctx.n_heads, ctx.in_dtype = n_heads, qkv.dtype
        return out.to(qkv.dtype)

    @staticmethod
    def backward(ctx, g):
        x, out, lse = ctx.saved_tensors
        L = _lib.load()
        dev = x.device
        S, B, _ = x.shape
        g = g.contiguous().float()
        dqkv = torch.empty_like(x)
        with torch.cuda.device(dev):
            _lib.check(L.pm_attn_bwd(x.data_ptr(), out.data_ptr(), g.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), S, B,
                                     ctx.n_heads, _stream(dev)), "pm_attn_bwd")
        return dqkv.to(ctx.in_dtype), None


def attention_hd8(qkv: torch.Tensor, n_heads: int) -> torch.Tensor:
    """The attention core of ``nn.MultiheadAttention`` for heads of dimension 8 (the reference critic's
    ``nn.TransformerEncoder(d_model 32, nhead 4)``, `pacman_mappo_resnet.py:131-136`): ``qkv`` [S, B, 3E] =
    ``in_proj(x)`` -> concatenated heads [S, B, E] = softmax(q k^T / sqrt(8)) v, differentiable (kernels pm_attn_fwd /
    pm_attn_bwd, plain fp32).  No [B*heads, S, S] tensor is ever materialised."""
    if not qkv.is_cuda:
        raise RuntimeError("attention_hd8 expects CUDA tensors; there is no CPU fallback")
    if qkv.dim() != 3 or qkv.shape[2] != 24 * n_heads or qkv.shape[0] > 1024:
        raise ValueError("attention_hd8: qkv must be [S <= 1024, B, 3 * 8 * n_heads]")
    return _Attention8.apply(qkv, n_heads)


class _AddLayerNorm32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, weight, bias, eps):
        L = _lib.load()
        dev = a.device
        a_, b_ = a.contiguous(), (b.contiguous() if b is not None else None)
        rows = a_.numel() // 32
        y = torch.empty_like(a_)
        need = a.requires_grad or (b is not None and b.requires_grad) or weight.requires_grad
        stats = torch.empty((rows, 2), dtype=torch.float32, device=dev) if need else None
        with torch.cuda.device(dev):
            _lib.check(L.pm_add_layernorm32_fwd(a_.data_ptr(), b_.data_ptr() if b_ is not None else None, weight.data_ptr(),
                                                bias.data_ptr(), y.data_ptr(), stats.data_ptr() if need else None, rows,
                                                float(eps), _stream(dev)), "pm_add_layernorm32_fwd")
        if need:
            ctx.save_for_backward(a_, b_ if b_ is not None else a_.new_empty(0), weight, stats)
        ctx.has_b = b is not None
        return y

    @staticmethod
    def backward(ctx, g):
        a_, b_, weight, stats = ctx.saved_tensors
        L = _lib.load()
        dev = a_.device
        rows = a_.numel() // 32
        g = g.contiguous()
        dx = torch.empty_like(a_)
        partial = torch.empty((int(L.pm_layernorm32_blocks(rows)), 64), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.pm_add_layernorm32_bwd(a_.data_ptr(), b_.data_ptr() if ctx.has_b else None, weight.data_ptr(),
                                                stats.data_ptr(), g.data_ptr(), dx.data_ptr(), partial.data_ptr(), rows,
                                                _stream(dev)), "pm_add_layernorm32_bwd")
        s = partial.sum(0)
        return dx, (dx if ctx.has_b else None), s[:32], s[32:], None


def _kernel_ok(*tensors) -> bool:
    return all(t is None or (t.is_cuda and t.dtype == torch.float32 and t.data_ptr() % 16 == 0) for t in tensors)


def add_layer_norm_d32(a: torch.Tensor, b, norm: "torch.nn.LayerNorm") -> torch.Tensor:
    """``norm(a + b)`` for ``nn.LayerNorm(32)`` (the post-norm residuals of the critic's encoder layers,
    `pacman_mappo_resnet.py:131-136`) as one kernel forward and one backward; ``b`` may be None.  Falls through to
    ``norm`` itself for anything the kernel does not cover (other widths, bf16 autocast, CPU tensors)."""
    if (a.shape[-1] != 32 or tuple(norm.normalized_shape) != (32,) or norm.weight is None or norm.bias is None
            or torch.is_autocast_enabled("cuda") or not _kernel_ok(a, b, norm.weight, norm.bias)
            or (b is not None and b.shape != a.shape)):
        return norm(a if b is None else a + b)
    return _AddLayerNorm32.apply(a, b, norm.weight, norm.bias, norm.eps)


class _GroupNorm4x8(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, res, weight, bias, eps, gelu):
        L = _lib.load()
        dev = x.device
        x_, r_ = x.contiguous(), (res.contiguous() if res is not None else None)
        B, HW = x_.shape[0], x_.shape[2] * x_.shape[3]
        y = torch.empty_like(x_)
        need = x.requires_grad or (res is not None and res.requires_grad) or weight.requires_grad
        stats = torch.empty((B, 4, 2), dtype=torch.float32, device=dev) if need else None
        with torch.cuda.device(dev):
            _lib.check(L.pm_groupnorm_4x8_fwd(x_.data_ptr(), r_.data_ptr() if r_ is not None else None, weight.data_ptr(),
                                              bias.data_ptr(), y.data_ptr(), stats.data_ptr() if need else None, B, HW,
                                              float(eps), int(gelu), _stream(dev)), "pm_groupnorm_4x8_fwd")
        if need:
            ctx.save_for_backward(x_, r_ if r_ is not None else x_.new_empty(0), weight, bias, stats)
        ctx.has_res, ctx.gelu = res is not None, int(gelu)
        return y

    @staticmethod
    def backward(ctx, g):
        x_, r_, weight, bias, stats = ctx.saved_tensors
        L = _lib.load()
        dev = x_.device
        B, HW = x_.shape[0], x_.shape[2] * x_.shape[3]
        g = g.contiguous()
        dx = torch.empty_like(x_)
        dres = torch.empty_like(x_) if ctx.has_res else None
        partial = torch.empty((int(L.pm_groupnorm_blocks(B)), 64), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.pm_groupnorm_4x8_bwd(x_.data_ptr(), r_.data_ptr() if ctx.has_res else None, weight.data_ptr(), bias.data_ptr(),
                                              stats.data_ptr(), g.data_ptr(), dx.data_ptr(), dres.data_ptr() if ctx.has_res else None,
                                              partial.data_ptr(), B, HW, ctx.gelu, _stream(dev)), "pm_groupnorm_4x8_bwd")
        s = partial.sum(0)
        return dx, dres, s[:32], s[32:], None, None


def group_norm_4x8(x: torch.Tensor, norm: "torch.nn.GroupNorm", res=None, gelu: bool = False) -> torch.Tensor:
    """``[gelu](norm(x) [+ res])`` for ``nn.GroupNorm(4, 32)`` on NCHW ``x`` - the two halves of the actor's ResidualBlock
    (`pacman_mappo_resnet.py:49-68`) as one kernel each, forward and backward.  Falls through to the PyTorch modules for
    anything the kernels do not cover (odd H*W, bf16 autocast, channels-last or CPU tensors)."""
    ok = (x.dim() == 4 and x.shape[1] == 32 and norm.num_groups == 4 and norm.num_channels == 32 and norm.weight is not None
          and not torch.is_autocast_enabled("cuda") and _kernel_ok(x, res, norm.weight, norm.bias) and x.is_contiguous()
          and (res is None or (res.shape == x.shape and res.is_contiguous())))
    if ok:
        HW = x.shape[2] * x.shape[3]
        ok = HW % 2 == 0 and HW <= int(_lib.load().pm_groupnorm_4x8_max_hw())
    if not ok:
        y = norm(x)
        if res is not None:
            y = y + res
        return torch.nn.functional.gelu(y) if gelu else y
    return _GroupNorm4x8.apply(x, res, norm.weight, norm.bias, norm.eps, gelu)


class FlatAdamEMA:
    """``clip_grad_norm_`` + ``Adam.step`` + the EMA copy of the policy (`pacman_mappo_resnet.py:588-595`) as two
    kernels over flat buffers.  The module's parameters (and the EMA module's) are re-pointed to views of one flat
    fp32 buffer each; gradients come from :class:`sharding.FlatGradBucket`'s flat buffer.  Same arithmetic as
    ``torch.optim.Adam(lr, eps=1e-5)`` followed by ``ema.mul_(d).add_(p, alpha=1-d)``."""

    def __init__(self, module, ema_module, grad_flat, lr, betas=(0.9, 0.999), eps=1e-5, ema_decay=0.995, max_grad_norm=0.5):
        self.L = _lib.load()
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.flat = self._flatten(self.params)
        self.ema_flat = self._flatten([p for p in ema_module.parameters() if p.requires_grad]) if ema_module is not None else None
        if grad_flat.numel() != self.flat.numel():
            raise ValueError("gradient bucket and parameters differ in size")
        self.grad = grad_flat
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.lr, self.betas, self.eps, self.ema_decay, self.max_grad_norm = lr, betas, eps, ema_decay, max_grad_norm
        self.step_count = 0
        dev = self.flat.device
        self.norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.scratch = torch.zeros(int(self.L.pm_adam_scratch_bytes()), dtype=torch.uint8, device=dev)

    @staticmethod
    def _flatten(params):
        """Re-point every parameter at a view of one flat buffer, same aligned layout as sharding.FlatGradBucket."""
        from .sharding import flat_offsets
        offs, n = flat_offsets(params)
        flat = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
        for p, o in zip(params, offs):
            flat[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = flat[o:o + p.numel()].view_as(p)
        return flat

    def step(self, grad_scale: float = 1.0):
        """One optimizer step on the gradient bucket (scaled by ``grad_scale``, e.g. 1/world after a sum all-reduce).
        Returns the device scalar holding the gradient norm before clipping."""
        self.step_count += 1
        dev = self.flat.device
        with torch.cuda.device(dev):
            _lib.check(self.L.pm_adam_ema(self.flat.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                          self.ema_flat.data_ptr() if self.ema_flat is not None else None,
                                          self.flat.numel(), self.step_count, float(self.lr), float(self.betas[0]),
                                          float(self.betas[1]), float(self.eps), float(self.max_grad_norm), float(grad_scale),
                                          float(self.ema_decay), self.norm.data_ptr(), self.scratch.data_ptr(), _stream(dev)),
                       "pm_adam_ema")
        return self.norm
