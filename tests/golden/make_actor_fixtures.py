"""Extracts the reference's trained TD3 actors and their recorded evaluation numbers into small fixtures.

Run in the build container (where /root/reference exists):  python tests/golden/make_actor_fixtures.py

The reference has no unit tests; besides the notebook outputs (make_notebook_fixtures.py) the only numbers it records
for the step path are the win rates / mean returns of its own checkpoints against the weak and the strong
BasicOpponent on the real pybox2d engine (Evaluator, rl/utils/evaluator.py:10-35: 100 complete episodes of
Hockey-One-v0 per evaluation, deterministic actor).  A policy trained on the real engine is a sensitive probe of the
whole step path (contacts, keep/shoot, TOI, rewards): it only reaches its recorded win rate on an engine that behaves
like the one it was trained on.  tests/test_actor_golden.py plays these actors on the CPU oracle and on the CUDA path.

Written: td3_actors.npz  -- float32 `policy` weights (ActorNetwork 18-256-256-4, rl/td3/networks.py:6-20) of
                            pretrained/stage_3/models/td3_best.pt and of the competition run's td3_best.pt
         td3_actors.json -- the recorded evaluation series (metrics/metrics.json) and best_winrate (run_info.json)
"""
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
CKPTS = {
    "stage_3": "pretrained/stage_3",
    "competition": "runs/20260216_113850_single_dual_eval_abcdefg_3(1)",
}


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    arrays, meta = {}, {"source": "julilili42/hockey-env checkpoints (policy state_dict) and metrics/metrics.json",
                        "protocol": "rl/utils/evaluator.py:10-35 -- 100 complete episodes per evaluation, Hockey-One-v0, "
                                    "deterministic actor, win = info['winner'] == 1, return = sum of step rewards"}
    for name, rel in CKPTS.items():
        ck = torch.load(os.path.join(REF, rel, "models", "td3_best.pt"), map_location="cpu", weights_only=False)
        for k, v in ck["policy"].items():
            arrays[f"{name}.{k}"] = v.detach().cpu().numpy().astype(np.float32)
        m = json.load(open(os.path.join(REF, rel, "metrics", "metrics.json")))
        info = {k: m[k] for k in ("winrates_strong", "winrates_weak", "reward_strong", "reward_weak") if k in m}
        ri = os.path.join(REF, rel, "config", "run_info.json")
        if os.path.exists(ri):
            info["best_winrate"] = json.load(open(ri)).get("run_result", {}).get("best_winrate")
        # the checkpoint is saved at the evaluation with the best min(strong, weak) win rate (rl/training/train.py)
        wm = np.minimum(np.array(info["winrates_strong"]), np.array(info["winrates_weak"]))
        b = int(np.argmax(wm))
        info["best_eval"] = {"index": b, "winrate_strong": info["winrates_strong"][b], "winrate_weak": info["winrates_weak"][b],
                             "reward_strong": info["reward_strong"][b], "reward_weak": info["reward_weak"][b]}
        info["checkpoint"] = rel + "/models/td3_best.pt"
        meta[name] = info
    np.savez_compressed(os.path.join(here, "td3_actors.npz"), **arrays)
    json.dump(meta, open(os.path.join(here, "td3_actors.json"), "w"), indent=1)
    print("wrote td3_actors.npz", {k: v.shape for k, v in arrays.items()})
    for name in CKPTS:
        print(name, meta[name]["best_eval"], "best_winrate", meta[name].get("best_winrate"))


if __name__ == "__main__":
    sys.exit(main())
