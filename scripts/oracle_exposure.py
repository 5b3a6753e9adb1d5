"""What "parity unpinned" can cost: the real Box2D / gym / pyglet stack is installable neither in the
build container nor on the GPU box (profiles/r02_probe_deps.txt), so the third-party arithmetic the oracle restates cannot be
checked against the real thing.  This script quantifies the exposure of the three choices SURVEY.md flags as unverifiable:

  D1   same-step shared-tile tie break (higher car id first)        -> flipped: lower car id first
  A.2  island joint order [j3, j2, j1, j0]                          -> reversed: [j0, j1, j2, j3]
  H2   fp32 rounding details (libm sinf ulps, contraction, ...)     -> proxy: one fp32 ulp added to a spawn coordinate

for the default two-agent episode (random policy, 1000 steps) over several seeds, CPU oracle only.
    python scripts/oracle_exposure.py [out.json] [seeds] [steps]"""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mcr_oracle as mo
from helpers import action_tape


def make(seed, A, variant=None, ulp=False):
    tr, _ = mo.generate_track(np.random.RandomState(9000 + seed))
    w = mo.OracleWorld(A)
    if variant:
        w.set_variant(**variant)
    w.set_track(tr, False)
    poses = mo.spawn_poses([tuple(r) for r in tr.nodes], {i: i for i in range(A)}, 'CCW')
    poses = np.array(poses, np.float64)
    if ulp:
        x32 = np.float32(poses[0, 1])
        poses[0, 1] = float(np.nextafter(x32, np.float32(np.inf)))
    w.spawn(poses)
    w.step(None)
    return w, tr


def run(seed, A, steps, **kw):
    w, tr = make(seed, A, **kw)
    tape = action_tape(100 + seed, steps, 1, A, brake_p=0.1)
    rewards = np.zeros(A); frames = []
    for s in range(steps):
        obs, r, done = w.step(tape[s, 0].astype(np.float64))
        rewards += r
        if s in (0, steps // 2, steps - 1):
            frames.append(obs.copy())
    vis, _ = w.visited()
    sc = w.scores()
    return dict(pose=w.bodies()[:, 0, :3].astype(np.float64), reward=np.array(sc[0]), counts=np.array(sc[1]),
                visited=np.array(vis), frames=frames, T=tr.T)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_oracle_exposure.json")
    seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    A = 2
    rows = []
    for seed in range(seeds):
        base = run(seed, A, steps)
        row = {"seed": seed, "T": base["T"], "reward": base["reward"].tolist(), "tiles_visited": base["counts"].tolist()}
        for name, kw in (("tie_break_flipped", dict(variant=dict(tie_ascending=True))),
                         ("joint_order_reversed", dict(variant=dict(joint_ascending=True))),
                         ("one_ulp_spawn_x", dict(ulp=True))):
            v = run(seed, A, steps, **kw)
            dpos = np.abs(v["pose"][:, :2] - base["pose"][:, :2]).max()
            rel = (np.abs(v["pose"] - base["pose"]) / np.maximum(np.abs(base["pose"]), 1.0)).max()
            row[name] = {
                "reward_delta_per_agent": (v["reward"] - base["reward"]).tolist(),
                "reward_sum_delta": float(v["reward"].sum() - base["reward"].sum()),
                "tiles_visited_delta": (v["counts"] - base["counts"]).tolist(),
                "visit_flags_differing": int((v["visited"] != base["visited"]).sum()),
                "hull_position_max_abs_diff": float(dpos), "pose_max_rel_diff": float(rel),
                "pixels_differing_last_frame": int((v["frames"][-1] != base["frames"][-1]).any(axis=-1).sum()),
            }
        rows.append(row)
        print(json.dumps(row), flush=True)
    summ = {}
    for name in ("tie_break_flipped", "joint_order_reversed", "one_ulp_spawn_x"):
        summ[name] = {
            "max_abs_reward_delta_per_agent": max(max(abs(x) for x in r[name]["reward_delta_per_agent"]) for r in rows),
            "max_abs_reward_sum_delta": max(abs(r[name]["reward_sum_delta"]) for r in rows),
            "max_tiles_visited_delta": max(max(abs(x) for x in r[name]["tiles_visited_delta"]) for r in rows),
            "max_pose_rel_diff": max(r[name]["pose_max_rel_diff"] for r in rows),
            "median_pose_rel_diff": float(np.median([r[name]["pose_max_rel_diff"] for r in rows])),
            "episodes_with_identical_visit_flags": sum(r[name]["visit_flags_differing"] == 0 for r in rows),
        }
    json.dump({"what": __doc__.split("\n\n")[0], "num_agents": A, "steps": steps, "episodes": seeds, "summary": summ, "rows": rows},
              open(out, "w"), indent=1)
    print(json.dumps(summ, indent=1))


if __name__ == "__main__":
    main()
