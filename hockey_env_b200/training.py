"""Device-side plumbing around the hot path for the reference's training loop -- SURVEY.md section 8f, ranks 2-4.

Everything here consumes and produces the env's zero-copy device tensors; nothing makes a host round trip per tick.
The dense networks are plain torch modules (library matmuls): they are the ops next to the hot path, not the path.

* `OpponentPool`        per-env opponent selection: weak / strong BasicOpponent (in-kernel) or self-play snapshots
                        (batched actor inference on obs_agent_two()), re-drawn per episode
                        (rl/training/opponent_manager.py:19-91, rl/training/self_play.py:7-68, hockey_env.py:908-922)
* `DeviceReplayBuffer`  ring buffer in HBM fed straight from step outputs (rl/replay/base_buffer.py:21-41,
                        rl/replay/uniform_buffer.py) -- `push` is one batched copy per tick instead of N Python calls
* `evaluate`            the Evaluator protocol (rl/utils/evaluator.py:10-35): win / draw / loss rates and mean return
                        of a policy against a fixed opponent over n COMPLETE episodes (fixed quota per env)
* `evaluate_model`      the ModelEvaluator protocol (model_evaluation/model_evaluator.py:81-108): 300 episodes, seed
                        123, against the weak and the strong BasicOpponent
"""
import torch

from . import _lib
from .env import HockeyVecEnv, Mode


class OpponentPool:
    """Draws, per env and per episode, one opponent out of {weak, strong, snapshot_0..k-1} with the given
    probabilities (the reference draws one per episode for its single env; here every env of the batch holds its own
    draw).  BasicOpponents run inside the step kernel; snapshot opponents are torch modules evaluated on
    `obs_agent_two()` for all envs at once, their actions are only read by the envs that drew them."""

    WEAK, STRONG, SNAPSHOT0 = 0, 1, 2

    def __init__(self, env, p_weak=0.5, p_strong=0.5, snapshots=(), p_snapshot=0.0, seed=0):
        if env.p2 != _lib.POLICY_PER_ENV:
            raise ValueError("OpponentPool needs HockeyVecEnv(..., p2='per_env')")
        self.env = env
        self.snapshots = list(snapshots)
        if p_snapshot > 0 and not self.snapshots:
            raise ValueError("p_snapshot > 0 without snapshots")
        k = len(self.snapshots)
        probs = [p_weak, p_strong] + [p_snapshot / k] * k
        self.probs = torch.tensor(probs, dtype=torch.float32, device=env.device)
        self.probs = self.probs / self.probs.sum()
        self.gen = torch.Generator(device=env.device)
        self.gen.manual_seed(seed)
        self.choice = torch.zeros(env.num_envs, dtype=torch.long, device=env.device)
        self.codes = env.opponent_codes
        self._a2 = torch.zeros((env.num_envs, 4), dtype=torch.float32, device=env.device)
        self.resample(None)

    def add_snapshot(self, actor, p_snapshot=None):
        """SelfPlayManager.add_snapshot (rl/training/self_play.py:30-45): later draws may pick `actor`."""
        self.snapshots.append(actor)
        k = len(self.snapshots)
        ps = float(self.probs[2:].sum()) if p_snapshot is None else float(p_snapshot)
        base = self.probs[:2] / self.probs[:2].sum() * (1.0 - ps)
        self.probs = torch.cat([base, torch.full((k,), ps / k, device=self.env.device)])

    def resample(self, done):
        """Re-draw the opponent of every env whose episode just ended (`done` = the step's done tensor, None = all)."""
        n = self.env.num_envs
        draw = torch.multinomial(self.probs.expand(n, -1), 1, generator=self.gen).squeeze(1)
        if done is None:
            self.choice.copy_(draw)
        else:
            self.choice.copy_(torch.where(done.to(torch.bool), draw, self.choice))
        code = torch.where(self.choice == self.WEAK, _lib.POLICY_BASIC_WEAK,
                           torch.where(self.choice == self.STRONG, _lib.POLICY_BASIC_STRONG, _lib.POLICY_EXTERNAL))
        self.codes.copy_(code.to(torch.uint8))  # in place: the step kernel reads this very tensor

    @torch.no_grad()
    def opponent_actions(self):
        """[N,4] actions of the snapshot opponents (zeros where an in-kernel BasicOpponent plays)."""
        if not self.snapshots:
            return self._a2
        obs2 = self.env.obs_agent_two()
        self._a2.zero_()
        for k, actor in enumerate(self.snapshots):
            m = self.choice == self.SNAPSHOT0 + k
            # evaluated for the whole batch (dense, no gather/scatter); only the rows that drew this snapshot are kept
            self._a2 = torch.where(m.unsqueeze(1), actor(obs2), self._a2)
        return self._a2

    def step(self, a1):
        """One env tick with player-1 actions `a1` [N,4]; returns the env's step tuple.  Opponents of finished
        episodes are re-drawn afterwards."""
        out = self.env.step(torch.cat([a1, self.opponent_actions()], dim=1).contiguous())
        self.resample(out[2])
        return out


class DeviceReplayBuffer:
    """Uniform replay buffer in device memory (rl/replay/base_buffer.py, uniform_buffer.py): `push` appends a whole
    batch of transitions (one per env) with five strided copies, `sample` gathers a minibatch with one index tensor."""

    def __init__(self, capacity, obs_dim=18, action_dim=4, device="cuda:0", seed=0):
        self.capacity, self.device = int(capacity), torch.device(device)
        self.obs = torch.empty((self.capacity, obs_dim), dtype=torch.float32, device=self.device)
        self.action = torch.empty((self.capacity, action_dim), dtype=torch.float32, device=self.device)
        self.reward = torch.empty(self.capacity, dtype=torch.float32, device=self.device)
        self.next_obs = torch.empty((self.capacity, obs_dim), dtype=torch.float32, device=self.device)
        self.done = torch.empty(self.capacity, dtype=torch.float32, device=self.device)
        self.pos, self.size = 0, 0
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)

    def __len__(self):
        return self.size

    def push(self, obs, action, reward, next_obs, done):
        """Batch of n transitions (n <= capacity); wraps around like the reference's ring buffer."""
        n = obs.shape[0]
        if n > self.capacity:
            raise ValueError("batch larger than the buffer")
        first = min(n, self.capacity - self.pos)
        for dst, src in ((self.obs, obs), (self.action, action), (self.reward, reward), (self.next_obs, next_obs),
                         (self.done, done.to(torch.float32))):
            dst[self.pos:self.pos + first].copy_(src[:first])
            if first < n:
                dst[:n - first].copy_(src[first:])
        self.pos = (self.pos + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def sample(self, batch_size):
        idx = torch.randint(0, self.size, (batch_size,), device=self.device, generator=self.gen)
        return self.obs[idx], self.action[idx], self.reward[idx], self.next_obs[idx], self.done[idx]


@torch.no_grad()
def collect(pool_or_env, actor, buffer, steps):
    """`steps` ticks of experience collection: actor -> env -> buffer, all on the device.  With auto-reset the
    transition of a finished episode stores the TERMINAL observation (`final_obs`) as next_obs, as the reference's
    loop does (rl/training/train.py:136-160)."""
    env = pool_or_env.env if isinstance(pool_or_env, OpponentPool) else pool_or_env
    obs = env.obs.clone()
    for _ in range(steps):
        a1 = actor(obs)
        nobs, reward, done, _, _ = pool_or_env.step(a1) if isinstance(pool_or_env, OpponentPool) else env.step(a1.contiguous())
        nxt = torch.where(done.to(torch.bool).unsqueeze(1), env.final_obs, nobs) if env.final_obs is not None else nobs
        buffer.push(obs, a1, reward, nxt, done)
        obs = nobs.clone()
    return buffer


@torch.no_grad()
def evaluate(actor, n_episodes=1000, opponent="strong", mode=Mode.NORMAL, num_envs=4096, device="cuda:0", seed=0,
             max_ticks=100000):
    """Evaluator.evaluate (rl/utils/evaluator.py:10-35) batched: `actor` (obs [N,18] -> actions [N,4]) plays COMPLETE
    episodes of Hockey-One-v0 against the in-kernel weak/strong BasicOpponent; returns the win rate (`winner == 1`) and
    the mean return (sum of step rewards) the reference reports, plus draw/loss rates and the mean episode length.

    Every env plays the same quota of k = ceil(n_episodes / num_envs) complete episodes and only those are counted (an
    env that has filled its quota keeps stepping but is ignored), so k * num_envs >= n_episodes episodes are played --
    stopping all envs once a global count is reached would keep only the shortest episodes and lose the 251-tick
    draws.  Sides alternate like the reference's reset (hockey_env.py:359-362): even envs start with player 1."""
    env = HockeyVecEnv(num_envs, mode=mode, device=device, seed=seed, p2=opponent)
    n = env.num_envs
    k = -(-int(n_episodes) // n)
    dev = env.device
    obs, _ = env.reset(one_starting=(torch.arange(n, device=dev) % 2 == 0).to(torch.int8))
    count = torch.zeros(n, dtype=torch.int64, device=dev)
    ret = torch.zeros(n, dtype=torch.float64, device=dev)
    length = torch.zeros(n, dtype=torch.int64, device=dev)
    acc = torch.zeros(6, dtype=torch.float64, device=dev)  # wins, draws, losses, sum return, sum return^2, sum length
    ticks = 0
    while ticks < max_ticks:
        obs, reward, done, _, info = env.step(actor(obs).contiguous())
        ticks += 1
        ret += reward.to(torch.float64)
        length += 1
        fin = done.to(torch.bool) & (count < k)
        w = info["winner"]
        f64 = fin.to(torch.float64)
        acc += torch.stack([(f64 * (w == 1)).sum(), (f64 * (w == 0)).sum(), (f64 * (w == -1)).sum(), (f64 * ret).sum(),
                            (f64 * ret * ret).sum(), (f64 * length).sum()])
        count += fin
        ret = torch.where(done.to(torch.bool), torch.zeros_like(ret), ret)
        length = torch.where(done.to(torch.bool), torch.zeros_like(length), length)
        if ticks % 16 == 0 and bool((count >= k).all().item()):  # one small host read every 16 ticks
            break
    env.close()
    a = acc.cpu().tolist()
    m = max(int(count.sum().item()), 1)
    mean_ret = a[3] / m
    return {"episodes": m, "win_rate": a[0] / m, "draw_rate": a[1] / m, "loss_rate": a[2] / m, "mean_return": mean_ret,
            "std_return": max(a[4] / m - mean_ret * mean_ret, 0.0) ** 0.5, "mean_length": a[5] / m, "ticks": ticks,
            "episodes_per_env": k}


@torch.no_grad()
def evaluate_model(actor, episodes=300, seed=123, num_envs=None, device="cuda:0"):
    """ModelEvaluator._eval_once for both opponents (model_evaluation/model_evaluator.py:81-108, defaults :234-235:
    300 complete episodes, seed 123): win rate and mean return of `actor` on Hockey-One-v0 against the weak and the
    strong BasicOpponent -- the four numbers of one row of the reference's final-evaluation table."""
    n = int(num_envs) if num_envs else int(episodes)
    weak = evaluate(actor, n_episodes=episodes, opponent="weak", num_envs=n, device=device, seed=seed)
    strong = evaluate(actor, n_episodes=episodes, opponent="strong", num_envs=n, device=device, seed=seed)
    return {"wr_weak": weak["win_rate"], "wr_strong": strong["win_rate"], "ret_weak": weak["mean_return"],
            "ret_strong": strong["mean_return"], "episodes": weak["episodes"]}
