"""SURVEY §8 row f3 as kernels - the PPO minibatch pipeline around the (PyTorch) networks,
`pacman_mappo_resnet.py:555-600`: uint8 gather + widen, the fused loss + gradient, clip + Adam + EMA over flat
buffers - against the reference's own formula (tests/golden/mappo_golden.npz, recorded by executing the reference's
update lines on seeded tensors) and against the same steps written as eager PyTorch ops.  Tolerance 1e-5 relative
(north_star's fp32 tolerance)."""
import os

import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(helpers.GOLDEN, "mappo_golden.npz"))


def torch_ppo(logits, vals, acts, old, adv, ret, clip_eps, vf, ec):
    lg = torch.log_softmax(logits, -1)
    lps = lg.gather(1, acts[:, None]).squeeze(1)
    ent = -(lg.exp() * lg).sum(-1)
    ratio = (lps - old).exp()
    pg = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - clip_eps, 1 + clip_eps)).mean()
    vl = 0.5 * ((vals - ret) ** 2).mean()
    return pg + vf * vl - ec * ent.mean(), torch.stack([pg, vl, ent.mean(), ((ratio - 1).abs() > clip_eps).float().mean()])


def test_fused_ppo_loss_matches_the_references_update_lines():
    from pacman_marl_2025_b200.mappo_ops import ppo_loss
    t = lambda k, dt=torch.float32: torch.as_tensor(G[k]).to(dt).cuda()
    logits, vals = t("ppo_logits").requires_grad_(True), t("ppo_values").requires_grad_(True)
    clip_eps, vf, ec = (float(x) for x in G["ppo_coefs"])
    loss, st = ppo_loss(logits, vals, t("ppo_actions", torch.int64), t("ppo_old_logp"), t("ppo_norm_adv"), t("ppo_ret"),
                        clip_eps, vf, ec)
    loss.backward()
    np.testing.assert_allclose(st[:5].cpu().numpy(), G["ppo_scalars"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(logits.grad.cpu().numpy(), G["ppo_dlogits"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(vals.grad.cpu().numpy(), G["ppo_dvalues"], rtol=1e-5, atol=1e-9)
    assert float(G["ppo_scalars"][4]) > 0.05                     # the fixture really has clipped samples


@pytest.mark.parametrize("B,pair", [(8192, True), (8192, False), (1000, False), (66, True), (300000, False)])
def test_fused_ppo_loss_vs_autograd(B, pair):
    from pacman_marl_2025_b200.mappo_ops import ppo_loss
    g = torch.Generator(device="cuda").manual_seed(B)
    r = lambda *s: torch.randn(*s, generator=g, device="cuda")
    logits, nv = (r(B, 5) * 2), (B // 2 if pair else B)
    vals = r(nv)
    acts = torch.randint(0, 5, (B,), generator=g, device="cuda")
    old = torch.log_softmax(logits + 0.5 * r(B, 5), -1).gather(1, acts[:, None]).squeeze(1)
    adv, ret = r(B), r(B)
    adv[::17] = 0.0
    la, va = logits.clone().requires_grad_(True), vals.clone().requires_grad_(True)
    lb, vb = logits.clone().requires_grad_(True), vals.clone().requires_grad_(True)
    loss, st = ppo_loss(la, va, acts, old, adv, ret, 0.15, 0.5, 0.02)
    (loss * 3.0).backward()                                       # a non-unit upstream gradient
    want, wst = torch_ppo(lb, torch.cat([vb, vb]) if pair else vb, acts, old, adv, ret, 0.15, 0.5, 0.02)
    (want * 3.0).backward()
    np.testing.assert_allclose(loss.item(), want.item(), rtol=2e-5)
    np.testing.assert_allclose(st[1:5].cpu().numpy(), wst.detach().cpu().numpy(), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(la.grad.cpu().numpy(), lb.grad.cpu().numpy(), rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(va.grad.cpu().numpy(), vb.grad.cpu().numpy(), rtol=1e-4, atol=1e-10)
    # deterministic: the same launch twice gives bit-identical scalars
    loss2, st2 = ppo_loss(la.detach(), va.detach(), acts, old, adv, ret, 0.15, 0.5, 0.02)
    assert torch.equal(st2[:5], st[:5])


@pytest.mark.parametrize("shape", [(8, 11, 14), (8, 7, 20), (8, 20, 20), (3, 5, 8)])
def test_gather_rows_is_indexing_plus_float(shape):
    from pacman_marl_2025_b200.mappo_ops import gather_rows
    g = torch.Generator(device="cuda").manual_seed(1)
    src = torch.randint(0, 256, (5000,) + shape, dtype=torch.uint8, device="cuda", generator=g)
    for n in (1, 31, 4096, 20000):
        idx = torch.randint(0, 5000, (n,), device="cuda", generator=g)
        assert torch.equal(gather_rows(src, idx), src[idx].float())
        assert torch.equal(gather_rows(src, idx, torch.bfloat16), src[idx].to(torch.bfloat16))
    with pytest.raises(RuntimeError, match="PM_EINVAL"):
        gather_rows(torch.zeros((10, 7), dtype=torch.uint8, device="cuda"), torch.zeros(3, dtype=torch.int64, device="cuda"))


def test_flat_adam_ema_matches_torch_adam_clip_and_ema():
    """pm_adam_ema == nn.utils.clip_grad_norm_ + torch.optim.Adam(eps=1e-5).step() + ema.mul_(d).add_(p, alpha=1-d)
    (`pacman_mappo_resnet.py:588-595`), over 30 steps with gradients that do and do not get clipped."""
    import copy
    from pacman_marl_2025_b200 import sharding
    from pacman_marl_2025_b200.mappo_ops import FlatAdamEMA
    torch.backends.cuda.matmul.allow_tf32 = False       # (a trainer test earlier in the process may have switched TF32 on:
    torch.backends.cudnn.allow_tf32 = False             #  two copies of a net at different addresses then round differently)
    # the bucket holds the SUM over ranks in real use: double one rank's gradient and step with grad_scale = 1/2
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.GELU(), torch.nn.Linear(64, 5)).cuda()
    ref_net, ref_ema, ema = copy.deepcopy(net), copy.deepcopy(net), copy.deepcopy(net)
    bucket = sharding.FlatGradBucket(net.parameters())
    opt = FlatAdamEMA(net, ema, bucket.flat, lr=2e-4, eps=1e-5, ema_decay=0.995, max_grad_norm=0.5)
    ref_opt = torch.optim.Adam(ref_net.parameters(), lr=2e-4, eps=1e-5)
    clipped = unclipped = 0
    for step in range(30):
        x = torch.randn(256, 37, device="cuda") * (0.05 if step % 3 == 0 else 3.0)
        bucket.zero_()
        net(x).square().mean().backward()
        bucket.flat.mul_(2.0)                                             # "sum over 2 ranks"
        norm = opt.step(grad_scale=0.5)
        ref_opt.zero_grad()
        ref_net(x).square().mean().backward()
        want_norm = torch.nn.utils.clip_grad_norm_(ref_net.parameters(), 0.5)
        ref_opt.step()
        with torch.no_grad():
            for e, p in zip(ref_ema.parameters(), ref_net.parameters()):
                e.mul_(0.995).add_(p, alpha=0.005)
        np.testing.assert_allclose(norm.item(), want_norm.item(), rtol=1e-5)
        clipped += want_norm.item() > 0.5
        unclipped += want_norm.item() <= 0.5
        for a, b in zip(net.parameters(), ref_net.parameters()):
            np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=1e-5, atol=1e-7, err_msg=f"step {step}")
    for a, b in zip(ema.parameters(), ref_ema.parameters()):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=1e-5, atol=1e-7)
    assert clipped > 3 and unclipped > 3


def test_trainer_update_fused_pipeline_equals_eager_pipeline():
    """One PPO update (3 epochs x 4 minibatches) from the same weights, rollout and shuffles: the kernels of row f3
    (gather + widen, fused loss + gradient, clip + Adam + EMA) leave the same policy, EMA and logged losses as the
    eager PyTorch formulation of `pacman_mappo_resnet.py:563-600`."""
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    res = []
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    for fused in (True, False):
        cfg = M.TrainConfig(num_envs=64, num_steps=16, num_minibatches=4, update_epochs=3, total_updates=10, length=12,
                            seed=21, checkpoint_every=0, fused_pipeline=fused, fast_math=False)
        tr = M.MAPPOTrainer(cfg, device="cuda:0")
        torch.manual_seed(77)
        lv, _ = tr.collect("randomTeam", play_as_red=False)
        torch.manual_seed(78)
        losses = tr.update(lv, tr.schedule(0))
        res.append((torch.cat([p.detach().reshape(-1) for p in tr.agent.parameters()]).cpu(),
                    torch.cat([p.detach().reshape(-1) for p in tr.ema_agent.parameters()]).cpu(),
                    {k: float(v) for k, v in losses.items()}))
    (pa, ea, la), (pb, eb, lb) = res
    for k in la:
        np.testing.assert_allclose(la[k], lb[k], rtol=2e-4, atol=1e-6, err_msg=k)
    # 12 Adam steps of size 2e-4: identical up to fp32 summation order in the losses' gradients
    assert float((pa - pb).abs().max()) < 2e-5, float((pa - pb).abs().max())
    assert float((ea - eb).abs().max()) < 2e-6


def test_trainer_rollout_against_the_references_cpu_bots():
    """Curriculum wiring of the host-bot hook: with TrainConfig.reference_root the reference's CPU teams
    (`pacman_mappo_resnet.py:40-42`) play red through gymPacMan_parallel_env(bot_team=...), each bot asked at its own
    sub-move on the reference's own GameState.  Needs the reference copy that build() puts in baseline/_ref."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    if not os.path.exists(os.path.join(root, "capture.py")):
        pytest.skip("no copy of the reference in baseline/_ref")
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    cfg = M.TrainConfig(num_envs=3, num_steps=8, num_minibatches=2, update_epochs=1, total_updates=4, length=20, seed=2,
                        checkpoint_every=0, reference_root=root, easy_until=0, selfplay_after=100)
    tr = M.MAPPOTrainer(cfg, device="cuda:0")
    assert M.curriculum_pools(root)[1] == ["MCTSTeam", "heuristicTeam"]
    lv, stats = tr.collect("heuristicTeam", play_as_red=False)
    env = tr._host_envs["heuristicTeam"]
    assert env._bot_team is not None and not isinstance(env.agents[0], int)
    ro = tr.rollout
    assert bool(torch.isfinite(ro.reward).all()) and bool(((ro.obs[:, :, :, 1] > 0).sum((-1, -2)) == 1).all())
    losses = tr.update(lv, tr.schedule(1))
    assert all(np.isfinite(float(v)) for v in losses.values())


@pytest.mark.parametrize("S,B,h", [(154, 64, 4), (140, 33, 4), (1, 5, 2), (400, 7, 4), (1024, 2, 1), (33, 1, 8)])
def test_attention_kernels_match_multihead_attention_math(S, B, h):
    """pm_attn_fwd / pm_attn_bwd == softmax(q k^T / sqrt(8)) v of nn.MultiheadAttention (the reference critic's
    nn.TransformerEncoder, `pacman_mappo_resnet.py:131-136`), forward and gradient, against an fp64 evaluation of the
    same formula; 1e-5 relative (of the tensor's scale)."""
    from pacman_marl_2025_b200.mappo_ops import attention_hd8
    g = torch.Generator(device="cuda").manual_seed(S * 131 + B)
    E = 8 * h
    qkv = (torch.randn((S, B, 3 * E), generator=g, device="cuda") * 1.7).requires_grad_(True)
    dout = torch.randn((S, B, E), generator=g, device="cuda")
    out = attention_hd8(qkv, h)
    out.backward(dout)

    x = qkv.detach().double().requires_grad_(True)
    q, k, v = (t.reshape(S, B * h, 8).transpose(0, 1) for t in x.chunk(3, dim=-1))
    p = torch.softmax(torch.bmm(q / np.sqrt(8.0), k.transpose(1, 2)), dim=-1)
    ref = torch.bmm(p, v).transpose(0, 1).reshape(S, B, E)
    ref.backward(dout.double())
    tol = lambda t: 1e-5 * float(t.abs().max())
    assert float((out.detach().double() - ref.detach()).abs().max()) <= tol(ref.detach())
    assert float((qkv.grad.double() - x.grad).abs().max()) <= tol(x.grad)


def test_critic_with_attention_kernel_equals_the_module_path():
    """MAPPOAgent.get_value with attn_impl 'kernel' == nn.TransformerEncoder.forward (the reference's call) in fp32:
    values and every parameter gradient."""
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(5)
        agent = M.MAPPOAgent((8, 11, 14), 5, 2).cuda()
        x = (torch.rand(96, 8, 11, 14, device="cuda") < 0.2).float()
        res = {}
        for impl in ("module", "kernel"):
            agent.attn_impl = impl
            agent.zero_grad(set_to_none=True)
            v = agent.get_value(x)
            (v * torch.linspace(-1, 1, 96, device="cuda")).sum().backward()
            res[impl] = (v.detach().clone(), {n: p.grad.clone() for n, p in agent.named_parameters() if p.grad is not None})
        np.testing.assert_allclose(res["kernel"][0].cpu().numpy(), res["module"][0].cpu().numpy(), rtol=2e-4, atol=2e-5)
        assert res["kernel"][1].keys() == res["module"][1].keys() and len(res["kernel"][1]) > 10
        for n, gm in res["module"][1].items():
            gk = res["kernel"][1][n]
            assert float((gk - gm).abs().max()) <= 2e-4 * float(gm.abs().max()) + 1e-6, n
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
        M.MAPPOAgent.attn_impl = "auto"


@pytest.mark.parametrize("rows,with_b", [(154 * 64, True), (7, True), (1, False), (100003, False)])
def test_add_layernorm32_kernel_matches_torch(rows, with_b):
    """pm_add_layernorm32_fwd / _bwd == nn.LayerNorm(32)(a + b) (the post-norm residual of the reference critic's encoder
    layers, `pacman_mappo_resnet.py:131-136`): output and every gradient against an fp64 evaluation, 1e-5 of the scale."""
    from pacman_marl_2025_b200.mappo_ops import add_layer_norm_d32
    g = torch.Generator(device="cuda").manual_seed(rows)
    norm = torch.nn.LayerNorm(32).cuda()
    with torch.no_grad():
        norm.weight.copy_(torch.randn(32, generator=g, device="cuda") * 0.5 + 1.0)
        norm.bias.copy_(torch.randn(32, generator=g, device="cuda") * 0.3)
    a = (torch.randn((rows, 32), generator=g, device="cuda") * 2 + 0.5).requires_grad_(True)
    b = torch.randn((rows, 32), generator=g, device="cuda").requires_grad_(True) if with_b else None
    dy = torch.randn((rows, 32), generator=g, device="cuda")
    y = add_layer_norm_d32(a, b, norm)
    y.backward(dy)
    got = [y.detach(), a.grad, norm.weight.grad.clone(), norm.bias.grad.clone()] + ([b.grad] if with_b else [])
    norm.zero_grad()
    n64 = torch.nn.LayerNorm(32).cuda().double()
    n64.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
    a64 = a.detach().double().requires_grad_(True)
    b64 = b.detach().double().requires_grad_(True) if with_b else None
    r = n64(a64 + b64 if with_b else a64)
    r.backward(dy.double())
    want = [r.detach(), a64.grad, n64.weight.grad, n64.bias.grad] + ([b64.grad] if with_b else [])
    for k, (x, w) in enumerate(zip(got, want)):
        assert float((x.double() - w).abs().max()) <= 2e-5 * max(1.0, float(w.abs().max())), k


@pytest.mark.parametrize("H,W,B", [(11, 14, 96), (7, 20, 33), (2, 4, 5), (1, 2, 3), (12, 16, 17), (20, 20, 21), (32, 32, 3), (13, 18, 9)])
@pytest.mark.parametrize("res,gelu", [(False, True), (True, True), (False, False)])
def test_groupnorm_4x8_kernel_matches_torch(H, W, B, res, gelu):
    """pm_groupnorm_4x8_fwd / _bwd == [gelu](nn.GroupNorm(4, 32)(x) [+ res]) - the two halves of the reference actor's
    ResidualBlock (`:49-68`) - output and every gradient against fp64; H*W = 154 makes float4s straddle channels, boards
    above 192 cells (20 x 20, 32 x 32, 13 x 18) take the CTA-per-group kernels."""
    from pacman_marl_2025_b200.mappo_ops import group_norm_4x8
    g = torch.Generator(device="cuda").manual_seed(H * 100 + W + B)
    norm = torch.nn.GroupNorm(4, 32).cuda()
    with torch.no_grad():
        norm.weight.copy_(torch.randn(32, generator=g, device="cuda") * 0.5 + 1.0)
        norm.bias.copy_(torch.randn(32, generator=g, device="cuda") * 0.3)
    x = (torch.randn((B, 32, H, W), generator=g, device="cuda") * 1.5 + 0.2).requires_grad_(True)
    r = torch.randn((B, 32, H, W), generator=g, device="cuda").requires_grad_(True) if res else None
    dy = torch.randn((B, 32, H, W), generator=g, device="cuda")
    y = group_norm_4x8(x, norm, res=r, gelu=gelu)
    y.backward(dy)
    got = [y.detach(), x.grad, norm.weight.grad.clone(), norm.bias.grad.clone()] + ([r.grad] if res else [])
    n64 = torch.nn.GroupNorm(4, 32).cuda().double()
    n64.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
    x64 = x.detach().double().requires_grad_(True)
    r64 = r.detach().double().requires_grad_(True) if res else None
    z = n64(x64)
    if res:
        z = z + r64
    if gelu:
        z = torch.nn.functional.gelu(z)
    z.backward(dy.double())
    want = [z.detach(), x64.grad, n64.weight.grad, n64.bias.grad] + ([r64.grad] if res else [])
    for k, (a, w) in enumerate(zip(got, want)):
        assert float((a.double() - w).abs().max()) <= 2e-5 * max(1.0, float(w.abs().max())), k


def test_networks_with_fused_norm_kernels_equal_the_plain_modules():
    """MAPPOAgent with the library's normalisation / attention kernels in the loop == the same modules evaluated by
    PyTorch alone (attn_impl 'module', ResidualBlock.fused_norm False): logits, values and every parameter gradient;
    on smallCapture's 11 x 14 (warp-per-group GroupNorm) and on a 20 x 20 board (CTA-per-group)."""
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        for H, W in ((11, 14), (20, 20)):
            torch.manual_seed(9)
            agent = M.MAPPOAgent((8, H, W), 5, 2).cuda()
            x = (torch.rand(48, 8, H, W, device="cuda") < 0.2).float()
            res = {}
            for fused in (False, True):
                M.ResidualBlock.fused_norm = fused
                agent.attn_impl = "kernel" if fused else "module"
                agent.zero_grad(set_to_none=True)
                lg, v = agent.logits(x), agent.get_value(x)
                ((lg * torch.linspace(-1, 1, 5, device="cuda")).sum() + (v * torch.linspace(-1, 1, 48, device="cuda")).sum()).backward()
                res[fused] = (lg.detach().clone(), v.detach().clone(), {n: p.grad.clone() for n, p in agent.named_parameters()})
            np.testing.assert_allclose(res[True][0].cpu().numpy(), res[False][0].cpu().numpy(), rtol=2e-4, atol=2e-5)
            np.testing.assert_allclose(res[True][1].cpu().numpy(), res[False][1].cpu().numpy(), rtol=2e-4, atol=2e-5)
            for n, gm in res[False][2].items():
                gk = res[True][2][n]
                assert float((gk - gm).abs().max()) <= 3e-4 * float(gm.abs().max()) + 1e-6, (H, W, n)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
        M.ResidualBlock.fused_norm = True
        M.MAPPOAgent.attn_impl = "auto"
