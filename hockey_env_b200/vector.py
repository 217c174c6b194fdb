"""gymnasium-VectorEnv-shaped wrapper over HockeyVecEnv.

Mirrors what `gymnasium.vector.SyncVectorEnv([lambda: HockeyEnv_BasicOpponent(...)] * n)` gives a caller of the
reference (hockey_env.py:875-903): `num_envs`, `single_observation_space`, `single_action_space`,
`reset(seed=, options=) -> (obs, infos)`, `step(actions) -> (obs, rewards, terminations, truncations, infos)` with
SAME-STEP auto-reset (the returned obs of a finished env is the first obs of its next episode; the terminal one is in
`infos["final_obs"]`, rows selected by `infos["_final_obs"]`).  Works without gymnasium installed.
"""
import numpy as np
import torch

from .env import HockeyVecEnv, Mode, _mkbox


class HockeyGymVectorEnv:
    metadata = {"autoreset_mode": "same_step"}

    def __init__(self, num_envs, mode=Mode.NORMAL, opponent="strong", device="cuda:0", seed=0, numpy_io=True):
        """opponent: 'weak' / 'strong' (HockeyEnv_BasicOpponent semantic, 4-d actions) or None (HockeyEnv, 8-d actions)."""
        self.env = HockeyVecEnv(num_envs, mode=mode, device=device, seed=seed, auto_reset=True, p2=opponent)
        self.num_envs = int(num_envs)
        self.numpy_io = bool(numpy_io)
        self.single_observation_space = _mkbox(-np.inf, np.inf, (18,))
        self.single_action_space = _mkbox(-1, +1, (4 if opponent else 8,))
        self.observation_space = _mkbox(-np.inf, np.inf, (self.num_envs, 18))
        self.action_space = _mkbox(-1, +1, (self.num_envs, 4 if opponent else 8))
        self.closed = False

    def _out(self, t):
        return t.detach().cpu().numpy() if self.numpy_io else t

    def _infos(self, done=None):
        e = self.env
        infos = {"winner": self._out(e.info[:, 0]), "reward_closeness_to_puck": self._out(e.info[:, 1]),
                 "reward_touch_puck": self._out(e.info[:, 2]), "reward_puck_direction": self._out(e.info[:, 3])}
        if done is not None:
            infos["final_obs"] = self._out(e.final_obs)
            infos["_final_obs"] = self._out(done.to(torch.bool))
        return infos

    def reset(self, *, seed=None, options=None):
        obs, _ = self.env.reset()
        return self._out(obs), self._infos()

    def step(self, actions):
        a = torch.as_tensor(actions, dtype=torch.float32, device=self.env.device)
        obs, reward, done, trunc, _ = self.env.step(a.contiguous())
        return (self._out(obs), self._out(reward), self._out(done.to(torch.bool)), self._out(trunc), self._infos(done))

    def close(self):
        if not self.closed:
            self.env.close()
            self.closed = True
