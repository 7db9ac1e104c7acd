"""Seeded synthetic worlds and scans for the named configurations: used by bench.py and the
full-size parity tests.  Pure numpy, host side only.
"""
import numpy as np

LETHAL = 254


def warehouse_static(size_x, size_y, clutter=0.0, seed=1):
    """1-cell border + shelves: 6-cell-thick bars every 80 rows, 160 cells long per 200 columns (axis-aligned, thick:
    the class on which the reference's inflation does not depend on its priority-queue tie order)."""
    g = np.zeros((size_y, size_x), np.uint8)
    g[0, :] = g[-1, :] = LETHAL
    g[:, 0] = g[:, -1] = LETHAL
    for y in range(40, size_y - 10, 80):
        for x in range(20, size_x - 10, 200):
            g[y:y + 6, x:min(x + 160, size_x - 1)] = LETHAL
    if clutter > 0:
        rng = np.random.default_rng(seed)
        g[rng.random((size_y, size_x)) < clutter] = LETHAL
    return g


def pallets(size_x, size_y, centers, half=3):
    """Axis-aligned boxes that exist only in the scans (>= 3 cells thick)."""
    g = np.zeros((size_y, size_x), bool)
    for (cx, cy) in centers:
        g[max(0, cy - half):cy + half + 1, max(0, cx - half):cx + half + 1] = True
    return g


def raycast_scan(occupied, resolution, origin_xy, sensor_xy, n_beams=360, max_range=10.0, z=0.3, phase=0.0):
    """Casts n_beams rays from sensor_xy through the boolean world `occupied`; returns float32 (n,3) end points: the
    centre of the first occupied cell hit, or the point at max_range when nothing is hit."""
    sy, sx = occupied.shape
    ang = phase + np.arange(n_beams) * (2 * np.pi / n_beams)
    step = 0.5 * resolution
    n_steps = int(max_range / step)
    t = (np.arange(1, n_steps + 1) * step)[None, :]
    px = sensor_xy[0] + np.cos(ang)[:, None] * t
    py = sensor_xy[1] + np.sin(ang)[:, None] * t
    mx = np.floor((px - origin_xy[0]) / resolution).astype(np.int64)
    my = np.floor((py - origin_xy[1]) / resolution).astype(np.int64)
    inside = (mx >= 0) & (mx < sx) & (my >= 0) & (my < sy)
    hit = np.zeros_like(inside)
    hit[inside] = occupied[my[inside], mx[inside]]
    stop = hit | ~inside
    first = np.where(stop.any(axis=1), stop.argmax(axis=1), n_steps - 1)
    rows = np.arange(n_beams)
    hx, hy = mx[rows, first], my[rows, first]
    was_hit = hit[rows, first]
    ex = np.where(was_hit, origin_xy[0] + (hx + 0.5) * resolution, px[rows, first])
    ey = np.where(was_hit, origin_xy[1] + (hy + 0.5) * resolution, py[rows, first])
    return np.stack([ex, ey, np.full(n_beams, z)], axis=1).astype(np.float32)


def warehouse_c3(size=4000, resolution=0.05, n_obs=8, n_beams=360, scan_range=10.0, seed=1, cycle=0):
    """Config C3: static warehouse + K observations x 360 beams obtained by ray-casting the world (static + pallets)
    from sensor poses along an aisle.  Returns (static_grid, observations, robot_pose, footprint)."""
    static = warehouse_static(size, size)
    rng = np.random.default_rng(seed)
    aisle_y = (40 + 6 + 80 + 40) // 2 + 80 * (size // 160)  # middle of an aisle near the map centre
    x_start = size // 2 - 200 + 7 * cycle
    centers = [(x_start + int(rng.integers(-60, 160)), aisle_y + int(rng.integers(-20, 20))) for _ in range(6)]
    world = (static == LETHAL) | pallets(size, size, centers)
    obs = []
    for k in range(n_obs):
        sx_cell = x_start + 12 * k
        sensor = ((sx_cell + 0.5) * resolution, (aisle_y + 0.5) * resolution)
        while world[int(sensor[1] / resolution), int(sensor[0] / resolution)]:
            sensor = (sensor[0] + 7 * resolution, sensor[1])
        pts = raycast_scan(world, resolution, (0.0, 0.0), sensor, n_beams, scan_range, phase=0.001 * k)
        obs.append(dict(origin=(sensor[0], sensor[1], 0.3), points=pts, obstacle_range=scan_range,
                        raytrace_range=scan_range, marking=True, clearing=True))
    robot = (obs[-1]["origin"][0], obs[-1]["origin"][1], 0.0)
    half = 0.325
    footprint = [(half, half), (half, -half), (-half, -half), (-half, half)]
    return static, obs, robot, footprint


def blocks_c1(size=400, seed=1, adversarial=False):
    """Config C1: 400x400 border walls + rectangular blocks >= 3 cells thick; adversarial adds 1 % single cells."""
    rng = np.random.default_rng(seed)
    g = np.zeros((size, size), np.uint8)
    g[0:3, :] = g[-3:, :] = LETHAL
    g[:, 0:3] = g[:, -3:] = LETHAL
    for _ in range(25):
        w, h = int(rng.integers(3, 40)), int(rng.integers(3, 40))
        x, y = int(rng.integers(0, size - w)), int(rng.integers(0, size - h))
        g[y:y + h, x:x + w] = LETHAL
    if adversarial:
        g[rng.random((size, size)) < 0.01] = LETHAL
    return g


def fleet_robot(robot_id, n=120, res=0.05):
    """Config C5: the independent inputs of one robot of the fleet, seeded by its id -- a raw (un-inflated) n x n local
    map made of thick axis-aligned structure (corridor walls and boxes; the class on which the reference's inflation
    is tie-order independent), its world origin, pose, velocity and a plan through it."""
    rng = np.random.default_rng(1_000_003 * 7 + robot_id)
    raw = np.zeros((n, n), np.uint8)
    lo, hi = int(n * rng.uniform(0.15, 0.3)), int(n * rng.uniform(0.7, 0.85))
    raw[lo - 4:lo - 1, :] = 254
    raw[hi + 1:hi + 4, :] = 254
    for _ in range(int(rng.integers(1, 4))):
        bx, by = int(rng.integers(n // 2, n - 12)), int(rng.integers(lo + 6, hi - 12))
        w, h = int(rng.integers(3, 9)), int(rng.integers(3, 9))
        raw[by:by + h, bx:bx + w] = 254
    ox, oy = float(rng.uniform(-50, 50)), float(rng.uniform(-50, 50))
    size = n * res
    mid = oy + (lo + hi) / 2 * res
    pose = (ox + size * rng.uniform(0.15, 0.35), mid + rng.uniform(-0.3, 0.3), float(rng.uniform(-0.4, 0.4)))
    vel = (float(rng.uniform(0.0, 0.5)), 0.0, float(rng.uniform(-0.4, 0.4)))
    t = np.linspace(-0.05, 1.2, int(rng.integers(40, 160)))
    plan = np.stack([pose[0] + t * size * 0.75, mid + 0.25 * np.sin(t * rng.uniform(1, 3)) * rng.uniform(0, 1)], 1)
    return dict(raw=raw, origin=(ox, oy), pose=pose, vel=vel, plan=plan)
