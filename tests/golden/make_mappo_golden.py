"""Records network / view-function known answers from the unmodified reference trainer
(`/root/reference/pacman_mappo_resnet.py`) into tests/golden/mappo_golden.npz.

Run in the build container (the reference is importable there; matplotlib is stubbed):
    python tests/golden/make_mappo_golden.py
The fixture holds: state_dict keys + shapes and parameter counts of MAPPOAgent for the three
shipped board sizes; actor logits / critic value for DETERMINISTIC weights (every tensor filled
by `det_fill`, so no checkpoint needs to be stored); canonicalize_obs / canonicalize_action /
merge_obs_for_critic outputs on a seeded observation pair.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
for m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(m, types.ModuleType(m))


def det_fill(state_dict):
    """Same formula in the test: value = 0.05 * sin(0.37 * i + k) over the flattened tensor, k = key index."""
    out = {}
    for k, (name, t) in enumerate(state_dict.items()):
        if not t.dtype.is_floating_point:
            out[name] = t
            continue
        i = torch.arange(t.numel(), dtype=torch.float64)
        out[name] = (0.05 * torch.sin(0.37 * i + k)).to(t.dtype).reshape(t.shape)
    return out


def det_obs(B, H, W, seed):
    g = np.random.default_rng(seed)
    o = (g.random((B, 8, H, W)) < 0.2).astype(np.float32)
    o[:, 1] = 0
    for b in range(B):
        o[b, 1, g.integers(H), g.integers(W)] = 1 + g.integers(0, 4)
    return torch.from_numpy(o)


def main():
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        import pacman_mappo_resnet as ref
    out = {}
    for name, (H, W) in dict(tiny=(7, 20), small=(11, 14), blox=(20, 20)).items():
        torch.manual_seed(0)
        agent = ref.MAPPOAgent((8, H, W), 5, 2)
        sd = agent.state_dict()
        out[f"{name}_keys"] = np.array(list(sd.keys()))
        out[f"{name}_shapes"] = np.array([",".join(map(str, v.shape)) for v in sd.values()])
        out[f"{name}_nparams"] = np.array(sum(p.numel() for p in agent.parameters()))
        if name == "blox":
            continue
        agent.load_state_dict(det_fill(sd))
        agent.eval()
        obs = det_obs(3, H, W, 1)
        with torch.no_grad():
            logits = agent.actor_head(agent.actor_backbone(obs))
            merged = torch.stack([ref.merge_obs_for_critic([obs[0], obs[1]]), ref.merge_obs_for_critic([obs[1], obs[2]])])
            value = agent._forward_critic(merged).squeeze(-1)
            v2, lp, ent = agent.evaluate(obs[:2], merged, torch.tensor([1, 4]))
        out[f"{name}_logits"], out[f"{name}_value"] = logits.numpy(), value.numpy()
        out[f"{name}_eval"] = np.stack([v2.numpy(), lp.numpy(), ent.numpy()])
        out[f"{name}_merged"] = merged.numpy()
        out[f"{name}_canon"] = ref.canonicalize_obs(obs, True).numpy()
    # The minibatch loss of the reference's update loop (`pacman_mappo_resnet.py:563-600`).  It is inline code in
    # train(), so the lines are executed here verbatim on seeded tensors (Categorical as in MAPPOAgent.evaluate,
    # `:199-207`), with autograd giving d loss / d logits and d loss / d values.
    from torch.distributions import Categorical
    g = torch.Generator().manual_seed(7)
    B = 512
    logits = (torch.randn(B, 5, generator=g) * 1.5).requires_grad_(True)
    vals = torch.randn(B, generator=g).requires_grad_(True)
    acts = torch.randint(0, 5, (B,), generator=g)
    b_logp = Categorical(logits=logits.detach() + 0.4 * torch.randn(B, 5, generator=g)).log_prob(acts)   # an older policy
    b_adv = torch.randn(B, generator=g) * 2 + 0.3
    b_ret = torch.randn(B, generator=g)
    clip_eps, ent_coef = ref.CLIP_EPS, ref.ENT_COEF_START
    dist = Categorical(logits=logits)
    lps, ent = dist.log_prob(acts), dist.entropy()
    norm_adv = (b_adv - b_adv.mean()) / (b_adv.std() + 1e-8)
    ratio = (lps - b_logp).exp()
    pg = -torch.min(norm_adv * ratio, norm_adv * torch.clamp(ratio, 1 - clip_eps, 1 + clip_eps)).mean()
    vl = 0.5 * ((vals - b_ret) ** 2).mean()
    loss = pg + ref.VF_COEF * vl - ent_coef * ent.mean()
    loss.backward()
    out["ppo_logits"], out["ppo_values"], out["ppo_actions"] = logits.detach().numpy(), vals.detach().numpy(), acts.numpy()
    out["ppo_old_logp"], out["ppo_adv"], out["ppo_norm_adv"], out["ppo_ret"] = b_logp.numpy(), b_adv.numpy(), norm_adv.numpy(), b_ret.numpy()
    out["ppo_scalars"] = np.array([loss.item(), pg.item(), vl.item(), ent.mean().item(),
                                   ((ratio - 1).abs() > clip_eps).float().mean().item()], np.float32)
    out["ppo_dlogits"], out["ppo_dvalues"] = logits.grad.numpy(), vals.grad.numpy()
    out["ppo_coefs"] = np.array([clip_eps, ref.VF_COEF, ent_coef], np.float64)
    out["canon_action"] = np.array([ref.canonicalize_action(a, True) for a in range(5)])
    out["canon_action_t"] = ref.canonicalize_action(torch.arange(5), True).numpy()
    np.savez_compressed(os.path.join(HERE, "mappo_golden.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if "keys" not in k and "shapes" not in k})


if __name__ == "__main__":
    main()
