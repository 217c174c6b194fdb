"""The notebook's 1000-game protocol (Hockey-Env.ipynb cells 50-59) as a statistical parity check.

The reference plays 1000 COMPLETE strong-vs-strong NORMAL games back to back (sides alternating, the two BasicOpponent
phases carried over from game to game) and prints: the number of steps, the 18 column means of every observation
returned by step() (the terminal one on the last tick of a game), the winners, and both players' reward sums.
Those 24 numbers are ONE sample of that protocol.  Here every env of a batch plays `games_per_env` complete games under
the same protocol; envs are pooled into replicas of 1000 games, which gives the sampling distribution of each statistic
under OUR engine, and the reference's sample is placed in it as a z-score.  (Sampling fixed-length time windows instead
of complete games would be length-biased -- long games over-represented -- and compare the wrong quantity.)
"""
import numpy as np

STAT_NAMES = [f"obs_mean[{k}]" for k in range(18)] + ["total_steps", "winners_plus1", "winners_zero", "winners_minus1",
                                                      "reward_sum_p1", "reward_sum_p2"]


def reference_sample(fx):
    return np.array(list(fx["obs_mean"]) + [fx["total_steps"], fx["winners_plus1"], fx["winners_zero"], fx["winners_minus1"],
                                            fx["reward_sums"][0], fx["reward_sums"][1]], np.float64)


def replicas_from_env_sums(obs_sum, steps, wdl, rsum, games_per_env, games=1000):
    """Per-env accumulators over `games_per_env` complete games -> [replicas, 24] statistics of `games`-game samples."""
    n = obs_sum.shape[0]
    per = games // games_per_env
    assert per * games_per_env == games and n % per == 0
    g = lambda a: a.reshape(n // per, per, -1).sum(1)
    o, s, w, r = g(obs_sum), g(steps.astype(np.float64))[:, 0], g(wdl), g(rsum)
    assert np.all(w.sum(1) == games)
    return np.concatenate([o / s[:, None], s[:, None], w, r], 1)


def zscores(fx, reps):
    ref = reference_sample(fx)
    return (ref - reps.mean(0)) / reps.std(0, ddof=1)
