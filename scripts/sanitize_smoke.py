"""Small workload that touches every kernel of the library (for compute-sanitizer; one tool per gpurun call)."""
import os
import sys
import torch
sys.path.insert(0, ".")
import hockey_env_b200 as hk

for tiers in ("2", "3"):
    os.environ["HK_TIERS"] = tiers
    for mode in (0, 1, 2):
        env = hk.HockeyVecEnv(700, mode=hk.Mode(mode), device="cuda:0", seed=mode, p1="strong", p2="weak", want_agent_two=True)
        for _ in range(120):
            env.step()
        env.rollout(8, "strong", "strong")
        s = env.get_full_state()
        env.set_full_state(s)
        env.set_state(env.obs.clone())
        env.reset(mask=torch.arange(700, device="cuda:0") % 3 == 0)
        ext = hk.HockeyVecEnv(300, mode=hk.Mode(mode), device="cuda:0", seed=9)
        for _ in range(60):
            ext.step(torch.rand(300, 8, device="cuda:0") * 2 - 1)
        torch.cuda.synchronize()
        print("tiers", tiers, "mode", mode, env.stats()["episodes"], env.stats()["overflows"])
