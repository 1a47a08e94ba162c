"""CPU checks of the trainer's host side against known answers recorded from the unmodified
reference trainer (tests/golden/make_mappo_golden.py -> mappo_golden.npz): MAPPOAgent has the
reference's state_dict keys / shapes / parameter counts (checkpoints interchange), gives the same
logits / values for the same weights, and the observation / action view functions agree."""
import os

import numpy as np
import pytest
import torch

import helpers

G = np.load(os.path.join(helpers.GOLDEN, "mappo_golden.npz"))
SHAPES = dict(tiny=(7, 20), small=(11, 14), blox=(20, 20))


def det_fill(state_dict):
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


@pytest.mark.parametrize("name", ["tiny", "small", "blox"])
def test_agent_state_dict_is_the_references(name):
    from pacman_marl_2025_b200.pacman_mappo_resnet import MAPPOAgent
    H, W = SHAPES[name]
    agent = MAPPOAgent((8, H, W), 5, 2)
    sd = agent.state_dict()
    assert list(sd.keys()) == [str(k) for k in G[f"{name}_keys"]]
    assert [",".join(map(str, v.shape)) for v in sd.values()] == [str(s) for s in G[f"{name}_shapes"]]
    n = sum(p.numel() for p in agent.parameters())
    assert n == int(G[f"{name}_nparams"]) == dict(tiny=2404694, small=2634070, blox=6664534)[name]


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_agent_forward_and_views_match_reference(name):
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    H, W = SHAPES[name]
    agent = M.MAPPOAgent((8, H, W), 5, 2)
    agent.load_state_dict(det_fill(agent.state_dict()))
    agent.eval()
    obs = det_obs(3, H, W, 1)
    merged = torch.stack([M.merge_obs_for_critic(obs[0], obs[1]), M.merge_obs_for_critic([obs[1], obs[2]])])
    np.testing.assert_array_equal(merged.numpy(), G[f"{name}_merged"])
    np.testing.assert_array_equal(M.canonicalize_obs(obs, True).numpy(), G[f"{name}_canon"])
    assert M.canonicalize_obs(obs, False) is obs
    with torch.no_grad():
        np.testing.assert_allclose(agent.logits(obs).numpy(), G[f"{name}_logits"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(agent.get_value(merged).numpy(), G[f"{name}_value"], rtol=1e-5, atol=1e-6)
        v, lp, ent = agent.evaluate(obs[:2], merged, torch.tensor([1, 4]))
    np.testing.assert_allclose(np.stack([v.numpy(), lp.numpy(), ent.numpy()]), G[f"{name}_eval"], rtol=1e-5, atol=1e-6)
    # batched merge == per-pair merge
    a, b = det_obs(5, H, W, 2), det_obs(5, H, W, 3)
    mb = M.merge_obs_for_critic(a, b)
    for i in range(5):
        assert torch.equal(mb[i], M.merge_obs_for_critic([a[i], b[i]]))


def test_action_canonicalisation_and_schedule():
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    assert [M.canonicalize_action(a, True) for a in range(5)] == list(G["canon_action"])
    assert M.canonicalize_action(torch.arange(5), True).tolist() == list(G["canon_action_t"])
    assert M.canonicalize_action(3, False) == 3
    # hyper-parameters of pacman_mappo_resnet.py:17-47
    assert (M.NUM_STEPS, M.BATCH_SIZE, M.GAMMA, M.GAE_LAMBDA, M.CLIP_EPS, M.VF_COEF, M.MAX_GRAD_NORM, M.UPDATE_EPOCHS,
            M.TOTAL_UPDATES, M.LR_START, M.LR_END, M.ENT_COEF_START, M.ENT_COEF_END, M.OPPONENT_POOL_SIZE,
            M.OPPONENT_UPDATE_FREQ, M.SHAPING_SCALE, M.EMA_DECAY) == (
        2048, 512, 0.99, 0.95, 0.15, 0.5, 0.5, 3, 2000, 2e-4, 4e-5, 0.02, 0.004, 80, 15, 0.1, 0.995)


def test_critic_written_out_with_bmm_equals_the_module():
    """MAPPOAgent.attn_impl = "bmm" evaluates the critic's nn.TransformerEncoder layers with explicit ops (same
    parameters): values and gradients equal nn.TransformerEncoder.forward's."""
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    torch.manual_seed(0)
    agent = M.MAPPOAgent((8, 11, 14), 5, 2)
    merged = det_obs(6, 11, 14, 5)
    outs, grads = {}, {}
    for impl in ("module", "bmm"):
        agent.attn_impl = impl
        agent.zero_grad()
        v = agent.get_value(merged)
        v.square().sum().backward()
        outs[impl] = v.detach().clone()
        grads[impl] = torch.cat([p.grad.reshape(-1) for n, p in agent.named_parameters() if n.startswith("critic")]).clone()
    np.testing.assert_allclose(outs["bmm"].numpy(), outs["module"].numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(grads["bmm"].numpy(), grads["module"].numpy(), rtol=1e-4, atol=1e-6)
    agent.eval()
    with torch.no_grad():
        agent.attn_impl = "module"; a = agent.get_value(merged)
        agent.attn_impl = "bmm"; b = agent.get_value(merged)
    np.testing.assert_allclose(b.numpy(), a.numpy(), rtol=1e-5, atol=1e-6)


def test_curriculum_pools_are_the_references_when_its_bots_are_available():
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    assert (M.EASY_TEAMS, M.MEDIUM_TEAMS, M.HARD_TEAMS) == (
        ["randomTeam"], ["MCTSTeam", "heuristicTeam"], ["baselineTeam", "AstarTeam", "approxQTeam"])   # `:40-42`
    assert M.curriculum_pools("/some/checkout") == (M.EASY_TEAMS, M.MEDIUM_TEAMS, M.HARD_TEAMS)
    assert M.curriculum_pools("") == (["randomTeam"], [], ["baselineTeam"])        # what runs inside the step kernel
