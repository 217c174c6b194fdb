"""On-device policy plumbing for BASELINE config 5 (self-play rollout with a TD3 actor) -- SURVEY.md section 8f rank 2.

The actor is the reference's `ActorNetwork` (rl/td3/networks.py:6-20: 18 -> 256 -> 256 -> 4, tanh after every layer);
it consumes the env's observation tensor in place (no host round trip) and its output feeds `HockeyVecEnv.step`.
The dense layers are plain library matmuls (torch/cuBLAS): they are the op adjacent to the hot path, not the path.
"""
import torch
from torch import nn


class ActorNetwork(nn.Module):
    def __init__(self, obs_dim=18, action_dim=4, hidden=256):
        super().__init__()
        self.fc1 = nn.Linear(obs_dim, hidden)
        self.fc2 = nn.Linear(hidden, hidden)
        self.fc3 = nn.Linear(hidden, action_dim)

    def forward(self, obs):
        x = torch.tanh(self.fc1(obs))
        x = torch.tanh(self.fc2(x))
        return torch.tanh(self.fc3(x))


def load_td3_actor(path, device="cuda:0", name=None):
    """The reference's trained policy as an on-device module.  `path` is either a reference checkpoint written by
    `TD3Agent.save` (rl/td3/agent.py:269-275: a dict whose `policy` entry is the ActorNetwork state_dict) or an
    `.npz` of float32 arrays `<name>.fc1.weight` ... `<name>.fc3.bias` (tests/golden/td3_actors.npz, extracted from
    those checkpoints by tests/golden/make_actor_fixtures.py)."""
    if str(path).endswith(".npz"):
        import numpy as np
        z = np.load(path)
        names = sorted({k.split(".")[0] for k in z.files})
        if name is None:
            if len(names) != 1:
                raise ValueError(f"{path} holds several actors {names}: pass name=")
            name = names[0]
        if name not in names:
            raise ValueError(f"no actor {name!r} in {path}; have {names}")
        sd = {k[len(name) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + ".")}
    else:
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        sd = ckpt["policy"] if "policy" in ckpt else ckpt
    hidden, obs_dim = sd["fc1.weight"].shape
    actor = ActorNetwork(obs_dim=obs_dim, action_dim=sd["fc3.weight"].shape[0], hidden=hidden)
    actor.load_state_dict(sd)
    return actor.to(device).eval()


@torch.no_grad()
def actor_rollout(env, actor, steps, opponent_actor=None):
    """`steps` ticks of `env` (a HockeyVecEnv with p1 external) driven by `actor`; player 2 is the env's in-kernel
    BasicOpponent, or `opponent_actor` acting on obs_agent_two() (the PolicyOpponent pattern, hockey_env.py:908-922)
    when the env was built with p2=None.  Returns the env's episode statistics."""
    obs = env.obs
    for _ in range(steps):
        a1 = actor(obs)
        if opponent_actor is not None:
            a2 = opponent_actor(env.obs_agent_two())
            obs, *_ = env.step(torch.cat([a1, a2], dim=1).contiguous())
        else:
            obs, *_ = env.step(a1.contiguous())
    return env.stats()
