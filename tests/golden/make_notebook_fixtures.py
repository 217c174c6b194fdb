"""Extracts the reference's result-pinning artefacts from its executed notebook into a small JSON fixture.

Run in the build container (where /root/reference exists):  python tests/golden/make_notebook_fixtures.py
The reference has no tests; these printed outputs are the only numbers it pins for the step path.
`reset_draws` is NOT in the notebook: it is the initial state recovered by fitting the three unknown reset draws
(puck x, puck y, aim y) to the six free-flight rewards with the oracle in the loop (residual <= 6e-8); player 2's
position is not observable in the trace and is parked out of the way.
"""
import json
import os
import re
import sys

NB = "/root/reference/Hockey-Env.ipynb"


def cell_text(c):
    out = ""
    for o in c.get("outputs", []):
        if "text" in o:
            out += "".join(o["text"])
        elif "data" in o and "text/plain" in o["data"]:
            out += "".join(o["data"]["text/plain"])
    return out


def main():
    nb = json.load(open(NB))
    cells = nb["cells"]
    fx = {"source": "julilili42/hockey-env Hockey-Env.ipynb (executed outputs)"}
    # cell 20: TRAIN_DEFENSE reward trace
    c20 = cells[20]
    assert "a1 = [0.1,0,0,1]" in "".join(c20["source"])
    rewards = [float(x) for x in cell_text(c20).split()]
    puck0 = (6.6970242293, 1.7652901733)
    aim = 4.2377813024
    fx["train_defense_trace"] = {
        "mode": "TRAIN_DEFENSE", "action": [0.1, 0, 0, 1, 0, 0, 0, 0], "rewards": rewards,
        # r_uniform() return values in draw order: p2 dx, p2 dy, puck dx, puck dy (pre 0.8 factor), aim (pre 0.6 factor)
        "reset_draws": [9.6 - 8.0, 2.0 - 4.0, puck0[0] - 5.0, (puck0[1] - 4.0) / 0.8, (aim - 4.0) / 0.6],
    }
    # cells 53-59: 1000 strong-vs-strong NORMAL games
    num = lambda s: [float(x) for x in re.findall(r"-?\d+\.\d+(?:e[-+]?\d+)?|-?\d+", s.replace("np.float64", ""))]
    winners = [int(x) for x in re.findall(r"-?\d+", cell_text(cells[56]))]
    fx["strong_vs_strong_1000_games"] = {
        "total_steps": int(num(cell_text(cells[53]))[0]),
        "obs_mean": num(cell_text(cells[54])),
        "winners_plus1": winners.count(1), "winners_zero": winners.count(0), "winners_minus1": winners.count(-1),
        "winner_mean": num(cell_text(cells[57]))[0], "winner_std": num(cell_text(cells[58]))[0],
        "reward_sums": num(cell_text(cells[59]))[:2],
    }
    fx["after_reset_info_agent_two_closeness"] = -0.11766339645208586  # cell 11
    fx["weak_vs_strong_game"] = {"sum_reward": num(cell_text(cells[39]))[0], "final": cell_text(cells[40])}
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "notebook_fixtures.json")
    json.dump(fx, open(out, "w"), indent=1)
    print("wrote", out, "| games:", len(winners), "| trace:", len(rewards))


if __name__ == "__main__":
    sys.exit(main())
