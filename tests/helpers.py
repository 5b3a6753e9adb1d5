"""Shared helpers: drive the CUDA path and the CPU oracle with identical inputs."""
import numpy as np


def action_tape(seed, steps, B, A, brake_p=0.3):
    """Seeded random policy inside the action bounds (steer [-1,1], gas [0,1], brake [0,1]);
    float32 like the reference's action_space dtype."""
    rs = np.random.RandomState(seed)
    a = np.empty((steps, B, A, 3), np.float32)
    a[..., 0] = rs.uniform(-1, 1, (steps, B, A))
    a[..., 1] = rs.uniform(0, 1, (steps, B, A))
    br = rs.uniform(0, 1, (steps, B, A))
    a[..., 2] = np.where(rs.uniform(0, 1, (steps, B, A)) < brake_p, br, 0.0)
    return a


def make_oracle_worlds(oracle, tracks, orders, directions, A, **kw):
    worlds = []
    for tr, order, d in zip(tracks, orders, directions):
        w = oracle.OracleWorld(A, **kw)
        w.set_track(tr, d == 'CW')
        poses = oracle.spawn_poses([tuple(r) for r in tr.nodes], {i: order[i] for i in range(A)}, d)
        w.spawn(poses)
        worlds.append(w)
    return worlds


def gpu_state(venv):
    """Host copies of the CUDA path's state in the oracle's getter layouts."""
    B, A = venv.batch_envs, venv.num_agents
    N = B * A
    body = venv.buffers["body"].cpu().numpy()          # (5, 10, N)
    awake = venv.buffers["awake"].cpu().numpy()        # (5, N)
    # oracle bodies: p.x p.y angle v.x v.y w c.x c.y awake
    ob = np.empty((N, 5, 9), np.float32)
    for i in range(5):
        ob[:, i, 0] = body[i, 6]; ob[:, i, 1] = body[i, 7]; ob[:, i, 2] = body[i, 2]
        ob[:, i, 3] = body[i, 3]; ob[:, i, 4] = body[i, 4]; ob[:, i, 5] = body[i, 5]
        ob[:, i, 6] = body[i, 0]; ob[:, i, 7] = body[i, 1]; ob[:, i, 8] = awake[i]
    wheel = venv.buffers["wheel"].cpu().numpy()        # (4, 2, N)
    ctrl = venv.buffers["ctrl"].cpu().numpy()          # (3, N)
    on_road = venv.buffers["on_road"].cpu().numpy()    # (4, N)
    ow = np.empty((N, 4, 6), np.float64)
    for k in range(4):
        ow[:, k, 0] = wheel[k, 0]; ow[:, k, 1] = wheel[k, 1]
        ow[:, k, 2] = ctrl[0] if k >= 2 else 0.0
        ow[:, k, 3] = ctrl[1]
        ow[:, k, 4] = ctrl[2] if k < 2 else 0.0
        ow[:, k, 5] = on_road[k]
    joint = venv.buffers["joint"].cpu().numpy()        # (4, 4, N)
    lim = venv.buffers["limit_state"].cpu().numpy()    # (4, N)
    oj = np.empty((N, 4, 5), np.float32)
    for k in range(4):
        for f in range(4):
            oj[:, k, f] = joint[k, f]
        oj[:, k, 4] = lim[k]
    return dict(
        bodies=ob.reshape(B, A, 5, 9), wheels=ow.reshape(B, A, 4, 6), joints=oj.reshape(B, A, 4, 5),
        reward=venv.buffers["reward"].cpu().numpy().reshape(B, A),
        counts=venv.buffers["visit_count"].cpu().numpy().reshape(B, A),
        backward=venv.buffers["backward"].cpu().numpy().reshape(B, A),
        grass=venv.buffers["on_grass"].cpu().numpy().reshape(B, A),
        visited=venv.buffers["visited"].cpu().numpy(), touched=venv.buffers["touched"].cpu().numpy())


def oracle_state(worlds):
    A = worlds[0].A
    bodies = np.stack([w.bodies() for w in worlds])
    wheels = np.stack([w.wheels() for w in worlds])
    wheels[..., 5] = (wheels[..., 5] > 0)
    joints = np.stack([w.joints() for w in worlds])[..., :5]
    sc = [w.scores() for w in worlds]
    vis = [w.visited() for w in worlds]
    return dict(bodies=bodies, wheels=wheels, joints=joints,
                reward=np.stack([s[0] for s in sc]), counts=np.stack([s[1] for s in sc]),
                backward=np.stack([s[2] for s in sc]), grass=np.stack([w.grass() for w in worlds]),
                visited=[v[0] for v in vis], touched=[v[1] for v in vis])


def visited_bits(vis_words, T, A):
    """(Tmax,) uint32 bitmask words -> (T, A) uint8 like the oracle's road_visited."""
    w = vis_words[:T].astype(np.int64)
    return np.stack([(w >> c) & 1 for c in range(A)], axis=1).astype(np.uint8)
