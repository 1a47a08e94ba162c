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
    out["canon_action"] = np.array([ref.canonicalize_action(a, True) for a in range(5)])
    out["canon_action_t"] = ref.canonicalize_action(torch.arange(5), True).numpy()
    np.savez_compressed(os.path.join(HERE, "mappo_golden.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if "keys" not in k and "shapes" not in k})


if __name__ == "__main__":
    main()
