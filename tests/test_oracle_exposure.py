"""Exposure of the choices the oracle could not check against a real Box2D (the probe of the GPU box,
profiles/r02_probe_deps.txt, found no Box2D / gym / pyglet / shapely and no wheel to install).  These tests pin the
QUALITATIVE consequences; scripts/oracle_exposure.py writes the numbers (profiles/r02_oracle_exposure.json).

  * D1 tie break flipped: only WHO gets the first-visitor share of the tiles both cars touch in the same step changes --
    the sum over agents, every visit flag / count and every pose stay identical.
  * joint order reversed / one-ulp perturbation: Gauss-Seidel order and rounding do change the trajectory (the system
    amplifies ulp-level differences), which is why pose parity is defined bit-exactly against the oracle's arithmetic."""
import numpy as np

from helpers import action_tape


def _run(oracle, seed, steps, variant=None, ulp=False, A=2):
    tr, _ = oracle.generate_track(np.random.RandomState(9000 + seed))
    w = oracle.OracleWorld(A)
    if variant:
        w.set_variant(**variant)
    w.set_track(tr, False)
    poses = np.array(oracle.spawn_poses([tuple(r) for r in tr.nodes], {i: i for i in range(A)}, 'CCW'), np.float64)
    if ulp:
        poses[0, 1] = float(np.nextafter(np.float32(poses[0, 1]), np.float32(np.inf)))
    w.spawn(poses)
    w.step(None)
    tape = action_tape(100 + seed, steps, 1, A, brake_p=0.1)
    total = np.zeros(A)
    for s in range(steps):
        total += w.step(tape[s, 0].astype(np.float64), render=False)[1]
    sc = w.scores()
    return dict(pose=w.bodies()[:, 0, :3].astype(np.float64), reward=np.array(sc[0]), counts=np.array(sc[1]),
                visited=np.array(w.visited()[0]), step_rewards=total, T=tr.T)


def test_tie_break_flip_only_moves_the_first_visitor_share(oracle):
    for seed in (0, 1, 2):
        a = _run(oracle, seed, 150)
        b = _run(oracle, seed, 150, variant=dict(tie_ascending=True))
        assert np.array_equal(a["pose"], b["pose"]) and np.array_equal(a["visited"], b["visited"]) and np.array_equal(a["counts"], b["counts"])
        assert abs(a["reward"].sum() - b["reward"].sum()) < 1e-9, "reward is conserved across agents"
        d = b["reward"] - a["reward"]
        # both cars sit on the same tiles at spawn: the higher car id is the first visitor there (D1); flipped, the lower one
        assert d[0] > 0 and abs(d[0] + d[1]) < 1e-9
        # a shared tile is worth 1000/T to its first visitor and (1 - 1/A) * 1000/T to the second: the share that moves
        share = 1000.0 / a["T"] / 2
        assert abs(d[0] / share - round(d[0] / share)) < 1e-6 and 1 <= round(d[0] / share) <= 12


def test_joint_order_and_ulp_perturbations_change_the_trajectory(oracle):
    base = _run(oracle, 0, 200)
    rev = _run(oracle, 0, 200, variant=dict(joint_ascending=True))
    ulp = _run(oracle, 0, 200, ulp=True)
    assert not np.array_equal(base["pose"], rev["pose"]), "Gauss-Seidel order is observable"
    assert not np.array_equal(base["pose"], ulp["pose"]), "a one-ulp difference is observable after 200 steps"
    for v in (rev, ulp):
        assert np.isfinite(v["pose"]).all() and np.abs(v["pose"][:, :2] - base["pose"][:, :2]).max() < 50.0
