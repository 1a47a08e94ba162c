"""Batched MAPPO trainer on the B200 capture env (BASELINE config 5).

Mirrors `/root/reference/pacman_mappo_resnet.py` — same hyper-parameter names
(`:17-47`), same network (`MAPPOAgent`, `:49-211`; identical ``state_dict`` keys,
so checkpoints move both ways), same reward (`team_reward + SHAPING_SCALE *
shaping`, `:518-522`), GAE (`:277-290`), PPO-clip loss with per-minibatch
advantage normalisation (`:563-590`), EMA policy (`:593-595`), opponent pool and
LR / entropy / clip schedules (`:385-447`) — but the rollout is N envs wide and
never leaves the GPU:

  reference (1 env, CPU loop)                      here
  -----------------------------------------------  ---------------------------------------------
  env.step() in Python, obs -> .to(device) x4       gymPacMan_parallel_env.step_batched (one launch,
  per step (`:461-538`)                             finished envs restart inside it)
  compute_heuristic_shaping on obs pairs (`:513`)   fused into the step kernel (shaped_reward)
  compute_gae python reverse loop (`:548-553`)      pm_gae kernel over [T, N*2] columns
  (adv - mean) / (std + 1e-8) (`:577`)              pm_adv_norm kernel
  single process                                    one process per GPU, env shards, flat-bucket
                                                    gradient all-reduce + stats all-reduce (NCCL)

The policy / value networks stay plain PyTorch (north_star).  Opponents: the
on-device ``randomTeam`` and self-play against the current / pooled EMA policy;
the reference's CPU search bots (`agents/*.py`) are outside the GPU path.
"""
from __future__ import annotations

import copy
import math
import os
import time
from collections import deque
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import mappo_ops, sharding
from .gymPacMan import gymPacMan_parallel_env

# ---- hyper-parameters: values of pacman_mappo_resnet.py:17-47 ------------------------------------
NUM_STEPS = 2048
BATCH_SIZE = 512
GAMMA = 0.99
GAE_LAMBDA = 0.95
CLIP_EPS = 0.15
VF_COEF = 0.5
MAX_GRAD_NORM = 0.5
UPDATE_EPOCHS = 3
TOTAL_UPDATES = 2000
LR_START, LR_END = 2e-4, 4e-5
ENT_COEF_START, ENT_COEF_END = 0.02, 0.004
OPPONENT_POOL_SIZE = 80
OPPONENT_UPDATE_FREQ = 15
SHAPING_SCALE = 0.1
EVAL_FREQ = 50
EVAL_EPISODES = 10
EMA_DECAY = 0.995
# the reference's curriculum pools (`:40-42`).  randomTeam and baselineTeam run inside the step kernel; the other four
# are the reference's CPU search bots: they join the pools when TrainConfig.reference_root points at a checkout of
# the reference, and are then served through the host-bot hook (gymPacMan_parallel_env(bot_team=...), each bot asked
# at its own sub-move on the reference's own GameState - reference_bridge.py).
EASY_TEAMS = ["randomTeam"]
MEDIUM_TEAMS = ["MCTSTeam", "heuristicTeam"]
HARD_TEAMS = ["baselineTeam", "AstarTeam", "approxQTeam"]
ON_DEVICE = ("randomTeam", "baselineTeam")

# canonical (blue-perspective) view of a red agent: mirror x, swap the capsule and food plane pairs
_CANON_PLANES = [0, 1, 3, 2, 4, 5, 7, 6]
_CANON_ACTION = [0, 3, 2, 1, 4]


# ---- networks (architecture of pacman_mappo_resnet.py:49-211) --------------------------------------
class ResidualBlock(nn.Module):
    """conv3x3 - GroupNorm(4) - GELU - conv3x3 - GroupNorm(4), skip, GELU (`:49-68`)."""

    def __init__(self, channels: int):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, 3, padding=1)
        self.gn1 = nn.GroupNorm(4, channels)
        self.act = nn.GELU()
        self.conv2 = nn.Conv2d(channels, channels, 3, padding=1)
        self.gn2 = nn.GroupNorm(4, channels)

    fused_norm = True      # GroupNorm + GELU (+ skip) as one kernel each (mappo_ops.group_norm_4x8) where it applies

    def forward(self, x):
        if self.fused_norm and x.is_cuda:
            from .mappo_ops import group_norm_4x8
            h = group_norm_4x8(self.conv1(x), self.gn1, gelu=True)
            return group_norm_4x8(self.conv2(h), self.gn2, res=x, gelu=True)
        y = self.gn2(self.conv2(self.act(self.gn1(self.conv1(x)))))
        return self.act(y + x)


class PositionalEncoding2D(nn.Module):
    """Fixed sin/cos code: first half of the channels encodes y, second half x (`:70-95`)."""

    def __init__(self, d_model: int, max_h: int = 50, max_w: int = 50):
        super().__init__()
        half = d_model // 2
        freq = torch.exp(torch.arange(0, half, 2) * -(math.log(10000.0) / half))

        def table(n):
            pos = torch.arange(n).unsqueeze(1)
            t = torch.zeros(n, half)
            t[:, 0::2] = torch.sin(pos * freq)
            t[:, 1::2] = torch.cos(pos * freq)
            return t
        self.register_buffer("y_enc", table(max_h))
        self.register_buffer("x_enc", table(max_w))

    def forward(self, x):
        _, _, H, W = x.shape
        pos = torch.cat([self.y_enc[:H, None, :].expand(H, W, -1), self.x_enc[None, :W, :].expand(H, W, -1)], dim=2)
        return x + pos.permute(2, 0, 1).unsqueeze(0)


class MAPPOAgent(nn.Module):
    """CNN/ResNet actor + small transformer critic (`:97-211`).  Batched entry points take
    ``obs [B,8,H,W]`` and ``merged [M,8,H,W]`` tensors directly."""

    def __init__(self, obs_shape, action_dim: int = 5, num_agents: int = 2):
        super().__init__()
        self.obs_shape = tuple(obs_shape)
        c_in, H, W = self.obs_shape
        width = 32
        self.actor_backbone = nn.Sequential(
            nn.Conv2d(c_in, 16, 3, padding=1), nn.GELU(),
            nn.Conv2d(16, width, 3, padding=1), nn.GELU(),
            ResidualBlock(width), ResidualBlock(width), ResidualBlock(width),
            nn.Flatten())
        self.actor_head = nn.Sequential(nn.Linear(width * H * W, 512), nn.LayerNorm(512), nn.GELU(),
                                        nn.Linear(512, action_dim))
        self.d_model = 32
        self.critic_projector = nn.Sequential(nn.Conv2d(c_in, self.d_model, 3, padding=1))
        self.pos_encoder = PositionalEncoding2D(self.d_model)
        layer = nn.TransformerEncoderLayer(d_model=self.d_model, nhead=4, dim_feedforward=128, dropout=0.0,
                                           batch_first=False)
        self.critic_transformer = nn.TransformerEncoder(layer, num_layers=2, enable_nested_tensor=False)
        self.critic_head = nn.Sequential(nn.Linear(self.d_model, 512), nn.GELU(), nn.Linear(512, 1))
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                nn.init.orthogonal_(m.weight, gain=math.sqrt(2))
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        nn.init.orthogonal_(self.actor_head[-1].weight, gain=0.01)
        nn.init.orthogonal_(self.critic_head[-1].weight, gain=1.0)

    # -- pieces --
    def logits(self, obs):
        return self.actor_head(self.actor_backbone(obs))

    # How the critic's two encoder layers are evaluated.  The PARAMETERS are nn.TransformerEncoder's (identical
    # state_dict keys); only the op sequence differs:
    #   "module"  nn.TransformerEncoder.forward (the reference's call, `:168-176`)
    #   "bmm"     the same post-norm layer written out with torch.bmm: scores = (q / sqrt(hd)) k^T, softmax, P v.
    #             With 4 heads of dimension 8 over H*W tokens, PyTorch's fused fp32 attention kernels
    #             (mem-efficient fmha) spend 22 ms per layer on a 4 096-sample minibatch, most of a PPO update;
    #             batched TF32 GEMMs + one softmax do the same arithmetic in a fraction of that.
    #   "kernel"  the same layer with the attention core (softmax(q k^T / sqrt(8)) v and its gradient) in the library's
    #             own kernels (mappo_ops.attention_hd8: one CTA per (sample, head), K / V in shared memory, plain
    #             fp32, nothing S x S materialised)
    #   "auto"    "module" under bf16 autocast (PyTorch's bf16 flash kernel), otherwise "kernel" when the heads have
    #             dimension 8, else "bmm".  B200, smallCapture, 4 096-sample minibatch, fwd+bwd of the whole critic:
    #             fp32 36.5 ms module / 27.9 ms bmm / 16.1 ms kernel; bf16 autocast 14.5 / 29.6 / 15.2 ms
    attn_impl = "auto"

    def _encoder_layer_bmm(self, layer, x, kernel=False):
        """One nn.TransformerEncoderLayer (norm_first=False, relu, dropout 0) on x [S, B, E] with explicit ops."""
        sa = layer.self_attn
        S, B, E = x.shape
        h = sa.num_heads
        hd = E // h
        qkv = nn.functional.linear(x, sa.in_proj_weight, sa.in_proj_bias)             # [S, B, 3E]
        if kernel:
            from .mappo_ops import attention_hd8
            o = attention_hd8(qkv, h)
        else:
            q, k, v = (t.reshape(S, B * h, hd).transpose(0, 1) for t in qkv.chunk(3, dim=-1))   # [B*h, S, hd]
            p = torch.softmax(torch.bmm(q * (1.0 / math.sqrt(hd)), k.transpose(1, 2)), dim=-1)
            o = torch.bmm(p, v).transpose(0, 1).reshape(S, B, E)
        if kernel:
            from .mappo_ops import add_layer_norm_d32
            x = add_layer_norm_d32(x, sa.out_proj(o), layer.norm1)
            return add_layer_norm_d32(x, layer.linear2(nn.functional.relu(layer.linear1(x))), layer.norm2)
        x = layer.norm1(x + sa.out_proj(o))
        return layer.norm2(x + layer.linear2(nn.functional.relu(layer.linear1(x))))

    def _forward_critic(self, merged):
        x = self.pos_encoder(self.critic_projector(merged))
        x = x.flatten(2).permute(2, 0, 1)                                  # [H*W tokens, B, d]
        impl = self.attn_impl
        if impl == "auto":
            sa = self.critic_transformer.layers[0].self_attn
            if torch.is_autocast_enabled("cuda"):
                impl = "module"
            elif x.is_cuda and sa.embed_dim // sa.num_heads == 8 and x.shape[0] <= 1024:
                impl = "kernel"
            else:
                impl = "bmm"
        if impl in ("bmm", "kernel"):
            for layer in self.critic_transformer.layers:
                x = self._encoder_layer_bmm(layer, x, kernel=impl == "kernel")
        else:
            x = self.critic_transformer(x)
        return self.critic_head(x.mean(dim=0))

    def get_value(self, merged):
        return self._forward_critic(merged).squeeze(-1)

    # -- rollout: sample --
    def get_action_and_value(self, obs, merged):
        """obs [B,8,H,W] (B = agents x envs), merged [M,8,H,W].  Returns action, log_prob, entropy per obs
        row and one value per merged row."""
        lg = torch.log_softmax(self.logits(obs), dim=-1)
        action = torch.multinomial(lg.exp(), 1).squeeze(-1)
        logp = lg.gather(1, action.unsqueeze(1)).squeeze(1)
        ent = -(lg.exp() * lg).sum(-1)
        return action, logp, self.get_value(merged), ent

    # -- update: evaluate stored actions --
    def evaluate(self, obs, merged_obs, action):
        lg = torch.log_softmax(self.logits(obs), dim=-1)
        logp = lg.gather(1, action.unsqueeze(1)).squeeze(1)
        ent = -(lg.exp() * lg).sum(-1)
        return (self.get_value(merged_obs) if merged_obs is not None else None), logp, ent

    @torch.no_grad()
    def get_deterministic_action(self, obs):
        return self.logits(obs).argmax(dim=-1)


# ---- observation / action views (pacman_mappo_resnet.py:215-238, 267-274) ------------------------------
_index_cache: Dict[tuple, torch.Tensor] = {}


def _index_tensor(name: str, values, device) -> torch.Tensor:
    """The small constant index tensors of the canonical views, one per device (a fresh torch.tensor(...) per call is a
    host-to-device copy: slow in the rollout loop and illegal while a CUDA graph is being captured)."""
    key = (name, str(device))
    if key not in _index_cache:
        _index_cache[key] = torch.tensor(values, device=device)
    return _index_cache[key]


def canonicalize_obs(obs: torch.Tensor, is_red_agent: bool) -> torch.Tensor:
    """Red agents see the board mirrored with the capsule (2,3) and food (6,7) plane pairs swapped, so one
    policy serves both sides.  obs [...,8,H,W]."""
    if not is_red_agent:
        return obs
    return obs.flip(-1).index_select(-3, _index_tensor("planes", _CANON_PLANES, obs.device))


def canonicalize_action(action, is_red_agent: bool):
    if not is_red_agent:
        return action
    if isinstance(action, torch.Tensor):
        return _index_tensor("action", _CANON_ACTION, action.device)[action]
    return _CANON_ACTION[int(action)]


def merge_obs_for_critic(obs_a: torch.Tensor, obs_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Critic input: the first learner's view with plane 1 = both team-mates' cells and plane 4 cleared.
    Accepts two [...,8,H,W] tensors or a list of them (the reference's calling convention)."""
    if obs_b is None:
        obs_a, obs_b = obs_a[0], obs_a[1]
    merged = obs_a.clone()
    merged[..., 1, :, :] = torch.maximum(obs_a[..., 1, :, :], obs_b[..., 1, :, :])
    merged[..., 4, :, :] = 0
    return merged


def compute_gae(rewards, values, dones, last_value, gamma=GAMMA):
    """`:277-290` on the GPU (kernel pm_gae); columns are independent."""
    return mappo_ops.compute_gae(rewards, values, dones, last_value, gamma, GAE_LAMBDA)


def curriculum_pools(reference_root: str = ""):
    """(easy, medium, hard) opponent pools of the curriculum (`:40-42, 396-438`).  With a checkout of the reference
    they are the reference's lists; without one only the teams that run inside the step kernel remain."""
    have = lambda teams: [t for t in teams if t in ON_DEVICE or reference_root]
    return have(EASY_TEAMS), have(MEDIUM_TEAMS), have(HARD_TEAMS)


# ---- the trainer ---------------------------------------------------------------------------------------
@dataclass
class TrainConfig:
    layout_file: str = ""
    num_envs: int = 1024              # per GPU
    num_steps: int = 128              # rollout length T (reference: 2048 x 1 env)
    num_minibatches: int = 8          # reference: 4096 samples / 512
    update_epochs: int = UPDATE_EPOCHS
    total_updates: int = TOTAL_UPDATES
    length: int = 300
    seed: int = 0
    amp_bf16: bool = False            # networks in bf16 autocast (fp32 master weights); default = reference fp32
    fast_math: bool = True            # cuDNN autotune + TF32 tensor-core matmuls for the fp32 networks
    use_views: bool = True            # learner views written by the step kernel (False: torch flip / merge per step)
    pair_agents: bool = True          # minibatches are drawn over env-steps and hold BOTH learners of each, so the critic
                                      # runs once per env-step (the reference shuffles (step, agent) samples and runs it
                                      # once per sample on the same merged observation, `:563-575`); False = reference order
    fused_pipeline: bool = True       # the minibatch pipeline around the nets as kernels (row f3): uint8 gather + widen,
                                      # fused PPO loss + gradient, clip + Adam + EMA over flat buffers; False = the
                                      # same steps as eager PyTorch ops (kept as the in-repo cross-check)
    graph_rollout: bool = True        # a rollout step (policy forward, sampling, env step, bookkeeping: ~150 small launches)
                                      # is captured into a CUDA graph per (opponent, side, t) on its second use and
                                      # replayed afterwards; host-side bots and False = eager launches
    easy_until: int = 200             # updates <= this: EASY_TEAMS only (`:396-402`)
    selfplay_after: int = 800         # update index after which self-play opponents enter the mix (`:404-438`)
    log_every: int = 1
    checkpoint_every: int = 250
    checkpoint_dir: str = "."
    reference_root: str = ""          # a checkout of the reference: enables its CPU bot teams (MCTSTeam, heuristicTeam,
                                      # AstarTeam, approxQTeam) as host-side opponents in the curriculum (`:40-42`);
                                      # every env then asks a Python bot twice per step - use a modest num_envs
    load_checkpoint: str = ""         # LOAD_CHECKPOINT (`:44, 368-372`): a state_dict saved by this trainer or the reference's
    start_update: int = 0             # START_UPDATES (`:45`)


class Rollout:
    """On-device rollout buffers, agent-major ([T, 2, N, ...]) so that the step kernel can write each step's learner
    views straight into slot t+1 (slot T holds the bootstrap observation).  Observations are stored as uint8 (they
    hold small non-negative integers: 0/1 planes and 1+numCarrying), a quarter of the fp32 footprint; converted at
    minibatch time."""

    def __init__(self, T, N, obs_shape, device):
        self.obs = torch.zeros((T + 1, 2, N) + tuple(obs_shape), dtype=torch.uint8, device=device)
        self.merged = torch.zeros((T + 1, N) + tuple(obs_shape), dtype=torch.uint8, device=device)
        self.action = torch.zeros((T, 2, N), dtype=torch.int64, device=device)
        self.logp = torch.zeros((T, 2, N), dtype=torch.float32, device=device)
        self.reward = torch.zeros((T, 2, N), dtype=torch.float32, device=device)
        self.done = torch.zeros((T, 2, N), dtype=torch.float32, device=device)
        self.value = torch.zeros((T, 2, N), dtype=torch.float32, device=device)


class MAPPOTrainer:
    def __init__(self, cfg: TrainConfig, device=None):
        self.cfg = cfg
        self.rank, self.world = sharding.world_info()
        if cfg.fast_math:
            torch.backends.cudnn.benchmark = True
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        layout_file = cfg.layout_file or os.path.join(os.path.dirname(os.path.abspath(__file__)), "layouts",
                                                      "smallCapture.lay")
        base = self.rank * cfg.num_envs
        common = dict(layout_file=layout_file, length=cfg.length, reward_forLegalAction=True, defenceReward=True,
                      num_envs=cfg.num_envs, device=self.device, auto_reset=True, env_index_base=base,
                      shaping_scale=SHAPING_SCALE, opponent_seed=cfg.seed)
        self.env_selfplay = gymPacMan_parallel_env(self_play=True, **common)
        self.env_bot = gymPacMan_parallel_env(self_play=False, enemieName="randomTeam", **common)
        self.obs_shape = self.env_selfplay.obs_shape
        torch.manual_seed(cfg.seed)                                   # same initial weights on every rank
        self.agent = MAPPOAgent(self.obs_shape, 5, 2).to(self.device)
        if cfg.load_checkpoint:
            self.agent.load_state_dict(torch.load(cfg.load_checkpoint, map_location=self.device))
        sharding.broadcast_module_(self.agent)
        self.ema_agent = copy.deepcopy(self.agent)
        self.opponent = copy.deepcopy(self.agent).eval()
        self._params = list(self.agent.parameters())
        self._ema_params = list(self.ema_agent.parameters())
        self.bucket = sharding.FlatGradBucket(self.agent.parameters())
        if cfg.fused_pipeline:
            # parameters / EMA / Adam moments as flat fp32 buffers: clip + Adam + EMA are two kernels per step
            self.optimizer = mappo_ops.FlatAdamEMA(self.agent, self.ema_agent, self.bucket.flat, lr=LR_START, eps=1e-5,
                                                   ema_decay=EMA_DECAY, max_grad_norm=MAX_GRAD_NORM)   # `:366, 588-595`
        else:
            self.optimizer = torch.optim.Adam(self.agent.parameters(), lr=LR_START, eps=1e-5, fused=cfg.fast_math)   # `:366`
        self.opponent_pool = deque(maxlen=OPPONENT_POOL_SIZE)
        self.opponent_pool.append(copy.deepcopy(self.ema_agent.state_dict()))
        self.rollout = Rollout(cfg.num_steps, cfg.num_envs, self.obs_shape, self.device)
        self._pair = cfg.pair_agents and (cfg.num_steps * cfg.num_envs) % cfg.num_minibatches == 0
        self.rng = np.random.default_rng(cfg.seed)                    # curriculum draws: identical on every rank
        torch.manual_seed(cfg.seed * 1000003 + self.rank)             # action sampling: different per rank
        self.log: Dict[str, List[float]] = {k: [] for k in (
            "update", "reward", "ep_return", "train_winrate", "entropy", "clip_frac", "pg_loss", "v_loss", "lr",
            "ent_coef", "sps")}
        N = cfg.num_envs
        self._ep_return = torch.zeros(N, dtype=torch.float64, device=self.device)
        self._score = torch.zeros(N, dtype=torch.int32, device=self.device)       # running game score per env
        # persistent rollout scratch (the CUDA graphs of the rollout steps hold on to these addresses)
        self._stats = torch.zeros(4, dtype=torch.float64, device=self.device)
        self._all_envs = torch.ones(N, dtype=torch.uint8, device=self.device)
        self._actions = torch.zeros((N, 4), dtype=torch.int64, device=self.device)
        self._graphs: Dict[tuple, Dict[int, "torch.cuda.CUDAGraph"]] = {}
        self._graph_uses: Dict[tuple, int] = {}
        self._graph_pool = None

    # ------------------------------------------------------------------ helpers
    def _autocast(self):
        return torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.cfg.amp_bf16)

    def _views(self, obs4, ids, as_red):
        """obs4 [4,N,8,H,W] -> (actor batch [2N,8,H,W] agent-major, merged critic batch [N,8,H,W])."""
        a, b = canonicalize_obs(obs4[ids[0]], as_red), canonicalize_obs(obs4[ids[1]], as_red)
        return torch.cat([a, b], dim=0), merge_obs_for_critic(a, b)

    def schedule(self, update):
        """LR / entropy / clip annealing and the opponent curriculum (`:385-447`), restricted to opponents that
        exist on the device: randomTeam, baselineTeam, the current policy, a pooled EMA snapshot."""
        cfg = self.cfg
        progress = update / max(cfg.total_updates, 1)
        lr = LR_START - (LR_START - LR_END) * progress
        ent_coef = ENT_COEF_START - (ENT_COEF_START - ENT_COEF_END) * progress
        clip_eps = 0.1 if update > cfg.selfplay_after else CLIP_EPS
        easy, medium, hard = curriculum_pools(cfg.reference_root)
        pick = lambda pool: str(pool[int(self.rng.integers(len(pool)))])
        if update <= cfg.easy_until:
            mode, play_as_red = pick(easy), False
        elif update <= cfg.selfplay_after:
            mode, play_as_red = pick(easy + medium + hard * 5), False
        else:
            r = self.rng.random()
            if r < 0.40:
                mode, play_as_red = "Self(Curr)", bool(self.rng.random() > 0.5)
            elif r < 0.60:
                mode, play_as_red = "Self(Old)", bool(self.rng.random() > 0.5)
            else:
                mode, play_as_red = pick(hard * 5 + medium + easy), False
        return dict(lr=lr, ent_coef=ent_coef, clip_eps=clip_eps, mode=mode, play_as_red=play_as_red)

    # ------------------------------------------------------------------ rollout
    @torch.no_grad()
    def collect(self, mode: str, play_as_red: bool):
        """T steps of N envs (`:457-538`).  Returns (last_value [N], stats tensor)."""
        cfg, ro = self.cfg, self.rollout
        T, N = cfg.num_steps, cfg.num_envs
        selfplay = mode.startswith("Self")
        if selfplay:
            env = self.env_selfplay
        elif mode in ON_DEVICE:
            env = self.env_bot
            if env.enemieName != mode:
                env.reset(enemieName=mode)                            # `:399, 406, 436`: env.reset(enemieName=opp_name)
        else:
            env = self._host_bot_env(mode)                            # the reference's CPU bots, asked at their own sub-move
        use_views = cfg.use_views and env._bot_team is None           # (a step split over two launches has no fused views)
        if mode == "Self(Curr)":
            self.opponent.load_state_dict(self.agent.state_dict())
        elif mode == "Self(Old)":
            self.opponent.load_state_dict(self.opponent_pool[int(self.rng.integers(len(self.opponent_pool)))])
        learner_ids, opp_ids = ([0, 2], [1, 3]) if play_as_red else ([1, 3], [0, 2])
        # `:457`: every rollout starts from fresh games.  After an env's first rollout the reset is a MASKED reset of all
        # envs, which advances the per-env episode counters instead of restarting them: the counters key the in-kernel
        # opponents' draws, so a later rollout does not replay the random choices of an earlier one (the reference's
        # bots draw from Python's global RNG, which never restarts either)
        if getattr(env, "_rolled_out", False):
            obs4, _ = env.reset_batched(self._all_envs)
        else:
            obs4, _ = env.reset_batched()
            env._rolled_out = True
        self._ep_return.zero_()
        self._score.zero_()
        stats = self._stats.zero_()                                   # episodes, return sum, wins, reward sum
        actions = self._actions.zero_()
        team_row = 0 if play_as_red else 1
        shp = self.obs_shape

        def store_views(slot, o4):                                    # torch path (`:462-469`) when the kernel views are off
            a_in, m = self._views(o4, learner_ids, play_as_red)
            ro.obs[slot] = a_in.view(2, N, *shp)
            ro.merged[slot] = m
        if use_views:
            env.observe_views(play_as_red, ro.obs[0], ro.merged[0])
        else:
            store_views(0, obs4)

        def step_body(t):
            """One rollout step (`:470-538`).  No host synchronisation and no per-call allocation of anything that
            outlives the step, so the whole body can be captured into a CUDA graph."""
            actor_in, merged = ro.obs[t].view(2 * N, *shp).float(), ro.merged[t].float()
            with self._autocast():
                act, logp, value, _ = self.agent.get_action_and_value(actor_in, merged)
            act = act.view(2, N)
            ro.action[t] = act
            ro.logp[t] = logp.float().view(2, N)
            ro.value[t] = value.float().unsqueeze(0).expand(2, N)
            env_act = canonicalize_action(act, play_as_red)
            actions[:, learner_ids[0]] = env_act[0]
            actions[:, learner_ids[1]] = env_act[1]
            if selfplay:
                o_in, o_merged = self._views(env._obs, opp_ids, not play_as_red)
                with self._autocast():
                    o_act, _, _, _ = self.opponent.get_action_and_value(o_in, o_merged)
                o_act = canonicalize_action(o_act.view(2, N), not play_as_red)
                actions[:, opp_ids[0]] = o_act[0]
                actions[:, opp_ids[1]] = o_act[1]
            if use_views:                                         # next step's network inputs land in slot t+1;
                env.set_view_outputs(play_as_red, ro.obs[t + 1], ro.merged[t + 1], raw_obs=selfplay)   # raw obs only for the opponent net
            out = env._step_with_host_bots(actions) if env._bot_team is not None else env.step_batched(actions)
            if not use_views:
                store_views(t + 1, out["obs"])                    # finished envs already show the new episode
            ro.reward[t, 0] = out["shaped_reward"][learner_ids[0]]
            ro.reward[t, 1] = out["shaped_reward"][learner_ids[1]]
            ro.done[t] = out["done"].to(torch.float32).unsqueeze(0).expand(2, N)
            # episode bookkeeping (`:516-538`): team_reward = sum over both learners = 2 x team reward
            team_reward = 2.0 * out["reward"][team_row]
            self._ep_return.add_(team_reward)
            db = out["done"].bool()
            # final score of a finished episode (`:525-531`): the env restarted in-kernel, so the score is
            # tracked here from the per-step score_change output
            self._score.add_(out["score_change"])
            won = (self._score > 0) if play_as_red else (self._score < 0)
            stats[0] += db.sum()
            stats[1] += (self._ep_return * db).sum()
            stats[2] += (db & won).sum()
            stats[3] += team_reward.sum()
            self._ep_return.masked_fill_(db, 0.0)
            self._score.masked_fill_(db, 0)

        # CUDA graphs: a step is ~150 small launches (two network forwards, sampling, the env kernel, bookkeeping) and
        # the loop is bound by their launch cost, not by the GPU.  The first rollout against an opponent runs eagerly
        # (cuDNN autotune, lazy initialisation); on the second the T step bodies are captured - one graph per t, because
        # step t reads slot t and writes slot t + 1 of the rollout buffer - and replayed from then on.
        key = (id(env), mode if not selfplay else "self", bool(play_as_red), bool(use_views), T)
        graphs = None
        if cfg.graph_rollout and env._bot_team is None:
            self._graph_uses[key] = self._graph_uses.get(key, 0) + 1
            if self._graph_uses[key] >= 2:
                graphs = self._graphs.setdefault(key, {})
        try:
            for t in range(T):
                if graphs is None:
                    step_body(t)
                    continue
                if t not in graphs:
                    g = torch.cuda.CUDAGraph()
                    if self._graph_pool is None:
                        self._graph_pool = torch.cuda.graph_pool_handle()
                    steps0, ver0 = env.steps, env._version
                    try:
                        with torch.cuda.graph(g, pool=self._graph_pool):
                            step_body(t)
                    except RuntimeError as e:                        # something in the step cannot be captured: stay eager
                        print(f"[MAPPOTrainer] rollout step not captured ({str(e).splitlines()[0][:120]}); eager launches", flush=True)
                        env.steps, env._version = steps0, ver0
                        self.cfg.graph_rollout, graphs = False, None
                        self._graphs.pop(key, None)
                        step_body(t)
                        continue
                    env.steps, env._version = steps0, ver0          # capturing runs the host code but not the kernels
                    graphs[t] = g
                if use_views:                                        # (host-side descriptor, for callers that look at it)
                    env.set_view_outputs(play_as_red, ro.obs[t + 1], ro.merged[t + 1], raw_obs=selfplay)
                graphs[t].replay()
                env.steps += 1
                env._version += 1
        finally:
            env.set_view_outputs(None)
        with self._autocast():
            last_value = self.agent.get_value(ro.merged[T].float()).float()
        return last_value, stats

    def _host_bot_env(self, team: str):
        """An env whose red side is played by the reference's own CPU bots of `team` (agents/<team>.py), one pair of
        bot objects serving every env of the batch through the host-bot hook."""
        if not hasattr(self, "_host_envs"):
            self._host_envs = {}
        if team not in self._host_envs:
            from . import reference_bridge as RB
            mods = RB.load_reference(self.cfg.reference_root)
            e0 = self.env_bot
            self._host_envs[team] = gymPacMan_parallel_env(
                layouts=[e0.layout], length=self.cfg.length, reward_forLegalAction=True, defenceReward=True, self_play=False,
                enemieName=team, bot_team=RB.make_bot_team(mods, team), state_factory=RB.state_factory(mods),
                num_envs=self.cfg.num_envs, device=self.device, auto_reset=True, env_index_base=e0.env_index_base,
                shaping_scale=SHAPING_SCALE)
        env = self._host_envs[team]
        env.reset()
        return env

    # ------------------------------------------------------------------ update
    def update(self, last_value, sched):
        """GAE + PPO epochs (`:540-600`).  Returns dict of mean losses."""
        cfg, ro = self.cfg, self.rollout
        T, N = cfg.num_steps, cfg.num_envs
        adv, ret = compute_gae(ro.reward.view(T, 2 * N), ro.value.view(T, 2 * N), ro.done.view(T, 2 * N),
                               last_value.repeat(2), GAMMA)          # columns are (agent, env): one critic value per env
        b_obs = ro.obs[:T].view(T * 2 * N, *self.obs_shape)            # sample s = (t, agent, env)
        b_merged = ro.merged[:T].view(T * N, *self.obs_shape)
        b_act, b_logp = ro.action.view(-1), ro.logp.view(-1)
        b_adv, b_ret = adv.reshape(-1), ret.reshape(-1)
        S = T * N * 2
        mb = S // cfg.num_minibatches
        fused = cfg.fused_pipeline
        if fused:
            self.optimizer.lr = sched["lr"]
        else:
            for g in self.optimizer.param_groups:
                g["lr"] = sched["lr"]
        clip_eps, ent_coef = sched["clip_eps"], sched["ent_coef"]
        sums = torch.zeros(4, device=self.device)
        steps = 0
        in_dtype = torch.bfloat16 if (fused and cfg.amp_bf16) else torch.float32
        widen = (lambda buf, idx: mappo_ops.gather_rows(buf, idx, in_dtype)) if fused else (lambda buf, idx: buf[idx].float())
        for _ in range(cfg.update_epochs):
            if self._pair:
                # shuffle env-steps e = (t, env); sample indices of its two learners: (t, 0, env) and (t, 1, env)
                pe = torch.randperm(T * N, device=self.device)
                s0 = torch.div(pe, N, rounding_mode="floor") * (2 * N) + pe % N
                half = mb // 2
                perm = torch.stack([s0.view(-1, half), (s0 + N).view(-1, half)], dim=1).reshape(-1)   # [mb] blocks: agent 0 | agent 1
            else:
                perm = torch.randperm(S, device=self.device)
            norm_adv_all = mappo_ops.normalize_advantages(b_adv, perm[: mb * cfg.num_minibatches], mb)
            for k in range(cfg.num_minibatches):
                idx = perm[k * mb:(k + 1) * mb]
                norm_adv = norm_adv_all[k * mb:(k + 1) * mb]
                with self._autocast():
                    if self._pair:
                        vals = self.agent.get_value(widen(b_merged, pe[k * half:(k + 1) * half]))      # one value per env-step
                    else:
                        midx = torch.div(idx, 2 * N, rounding_mode="floor") * N + idx % N   # (t, env) of sample (t, agent, env)
                        vals = self.agent.get_value(widen(b_merged, midx))
                    logits = self.agent.logits(widen(b_obs, idx))
                vals, logits = vals.float(), logits.float()
                if fused:
                    # log-softmax, ratio, clipped surrogate, value loss, entropy, the four logged scalars AND the
                    # gradient w.r.t. logits / values: one launch (`:575-585, 597-600`)
                    loss, st = mappo_ops.ppo_loss(logits, vals, b_act[idx], b_logp[idx], norm_adv, b_ret[idx], clip_eps,
                                                  VF_COEF, ent_coef)
                    self.bucket.zero_()
                    loss.backward()
                    self.bucket.all_reduce_sum_()                     # the one data-path collective of the trainer
                    self.optimizer.step(grad_scale=1.0 / self.world)  # mean over ranks, clip, Adam, EMA: two kernels
                    sums += st[1:5]
                else:
                    lg = torch.log_softmax(logits, dim=-1)
                    lps = lg.gather(1, b_act[idx].unsqueeze(1)).squeeze(1)
                    ent = -(lg.exp() * lg).sum(-1)
                    if self._pair:
                        vals = torch.cat([vals, vals])
                    ratio = (lps - b_logp[idx]).exp()
                    pg = -torch.min(norm_adv * ratio, norm_adv * torch.clamp(ratio, 1 - clip_eps, 1 + clip_eps)).mean()
                    vl = 0.5 * ((vals - b_ret[idx]) ** 2).mean()
                    loss = pg + VF_COEF * vl - ent_coef * ent.mean()
                    self.bucket.zero_()
                    loss.backward()
                    self.bucket.all_reduce_mean_()
                    self.bucket.clip_grad_norm_(MAX_GRAD_NORM)
                    self.optimizer.step()
                    with torch.no_grad():
                        torch._foreach_mul_(self._ema_params, EMA_DECAY)
                        torch._foreach_add_(self._ema_params, self._params, alpha=1 - EMA_DECAY)
                        sums += torch.stack([pg.detach(), vl.detach(), ent.mean().detach(),
                                             ((ratio - 1).abs() > clip_eps).float().mean()])
                steps += 1
        sums /= max(steps, 1)
        return dict(pg_loss=sums[0], v_loss=sums[1], entropy=sums[2], clip_frac=sums[3])

    # ------------------------------------------------------------------ loop
    def train_one(self, update: int):
        sched = self.schedule(update)
        t0 = time.perf_counter()
        last_value, stats = self.collect(sched["mode"], sched["play_as_red"])
        losses = self.update(last_value, sched)
        sharding.all_reduce_stats_(stats)
        torch.cuda.synchronize(self.device)
        dt = time.perf_counter() - t0
        if update % OPPONENT_UPDATE_FREQ == 0:
            self.opponent_pool.append(copy.deepcopy(self.ema_agent.state_dict()))
        eps, ret_sum, wins, rew_sum = stats.tolist()
        samples = 2 * self.cfg.num_envs * self.cfg.num_steps * self.world
        rec = dict(update=update, reward=rew_sum / (self.cfg.num_envs * self.world), ep_return=ret_sum / eps if eps else 0.0,
                   train_winrate=wins / eps if eps else 0.0, lr=sched["lr"], ent_coef=sched["ent_coef"],
                   sps=samples / dt, **{k: float(v) for k, v in losses.items()})
        for k, v in rec.items():
            self.log[k].append(v)
        rec.update(mode=sched["mode"], side="Red " if sched["play_as_red"] else "Blue", seconds=dt, episodes=eps)
        return rec

    def train(self, total_updates: Optional[int] = None, start: Optional[int] = None):
        n = self.cfg.total_updates if total_updates is None else total_updates
        start = self.cfg.start_update if start is None else start
        for update in range(start, n + 1):
            rec = self.train_one(update)
            if self.rank == 0 and update % self.cfg.log_every == 0:
                print(f"Upd {update:4d} [{rec['side']}|{rec['mode']:10s}] | Win: {rec['train_winrate'] * 100:5.1f}% | "
                      f"Rew: {rec['reward']:7.1f} | EpRet: {rec['ep_return']:6.1f} | Ent: {rec['entropy']:.3f} | "
                      f"Clip: {rec['clip_frac']:.3f} | LR: {rec['lr']:.2e} | SPS: {rec['sps']:.0f}", flush=True)
            if self.rank == 0 and self.cfg.checkpoint_every and update % self.cfg.checkpoint_every == 0:
                torch.save(self._ema_state(), os.path.join(self.cfg.checkpoint_dir, f"mappo_resnet_{update}.pt"))
        if self.rank == 0:
            torch.save(self._ema_state(), os.path.join(self.cfg.checkpoint_dir, "mappo_resnet_final.pt"))
        return self.log

    def _ema_state(self):
        """EMA weights as independent tensors (with the fused pipeline the live ones are views of one flat buffer)."""
        return {k: v.detach().clone() for k, v in self.ema_agent.state_dict().items()}

    @torch.no_grad()
    def evaluate_vs_bots(self, agent: Optional[MAPPOAgent] = None, num_envs: int = 256, team: str = "baselineTeam"):
        """`:293-337` against an on-device bot team: deterministic blue learners, one episode per env.
        Returns (mean return, std, win rate)."""
        agent = (agent or self.ema_agent).eval()
        env = gymPacMan_parallel_env(layouts=[self.env_bot.layout], length=self.cfg.length, self_play=False,
                                     enemieName=team, num_envs=num_envs, device=self.device,
                                     opponent_seed=self.cfg.seed + 7919)
        obs4, _ = env.reset_batched()
        ret = torch.zeros(num_envs, dtype=torch.float64, device=self.device)
        alive = torch.ones(num_envs, dtype=torch.bool, device=self.device)
        won = torch.zeros(num_envs, dtype=torch.bool, device=self.device)
        score = torch.zeros(num_envs, dtype=torch.int32, device=self.device)
        actions = torch.zeros((num_envs, 4), dtype=torch.int64, device=self.device)
        for _ in range(self.cfg.length + 2):
            a = agent.get_deterministic_action(torch.cat([obs4[1], obs4[3]], dim=0)).view(2, num_envs)
            actions[:, 1], actions[:, 3] = a[0], a[1]
            out = env.step_batched(actions)
            obs4 = out["obs"]
            ret += torch.where(alive, 2.0 * out["reward"][1], torch.zeros_like(ret))
            score += torch.where(alive, out["score_change"], torch.zeros_like(score))
            d = out["done"].bool()
            won |= alive & d & (score < 0)                            # blue wins on a negative final score (`:331-333`)
            alive &= ~d
            if not bool(alive.any()):
                break
        agent.train()
        return float(ret.mean()), float(ret.std()), float(won.double().mean())


def train(**kw):
    """Entry point with the reference's name (`:340`); kwargs override :class:`TrainConfig`."""
    return MAPPOTrainer(TrainConfig(**kw)).train()


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    for f, v in TrainConfig().__dict__.items():
        ap.add_argument("--" + f.replace("_", "-"), type=type(v) if not isinstance(v, bool) else int, default=v)
    a = ap.parse_args()
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", 1)) > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    train(**{k: (bool(v) if isinstance(getattr(TrainConfig(), k), bool) else v) for k, v in vars(a).items()})
