"""Pins the CPU checkers against the reference's own known-answer tests.

Runs for the restatement ("port") always and for the compiled reference ("ref") where oracle/_ref exists.
Each test names the reference test it restates.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
import scenarios as sc

MAX_Z = 1.0  # costmap_2d/include/costmap_2d/testing_helper.h:10


def ten_by_ten():
    """costmap_2d/test/TenByTen.pgm as drawn in obstacle_tests.cpp:45-69 (row 0 first)."""
    t = np.zeros((10, 10), np.uint8)
    t[0, 7:10] = t[1, 7:10] = 254
    t[2, 3:6] = 254
    t[5, 4] = t[6, 4] = 254
    t[5, 7:10] = t[6, 7:10] = t[7, 7:10] = 254
    return t


def radii(length, width):  # inflation_tests.cpp:49-72 setRadii
    return [(width, length), (width, -length), (-width, -length), (-width, length)]


def obs(x, y, z=0.0, ox=0.0, oy=0.0, oz=MAX_Z):  # testing_helper.h:75-91 addObservation
    return dict(origin=(ox, oy, oz), points=[(x, y, z)], obstacle_range=100.0, raytrace_range=100.0)


@pytest.fixture(params=["port", "ref"])
def api(request):
    return request.getfixturevalue(request.param)


def static_stack(api, track_unknown=False, inflation=None, footprint=None):
    cm = api.costmap(10, 10, 1.0, track_unknown=track_unknown)
    s = cm.add_grid_layer(po.TRUE_OVERWRITE)
    o = cm.add_obstacle_layer(1, True, 2.0)
    i = cm.add_inflation_layer(*inflation) if inflation else None
    if footprint:
        cm.set_footprint(footprint)
    cm.set_grid_layer(s, ten_by_ten())
    return cm, s, o, i


def test_fixture_has_20_occupied_cells():
    assert int((ten_by_ten() == 254).sum()) == 20


def test_raytracing(api):  # obstacle_tests.cpp:77-91 testRaytracing
    cm, s, o, _ = static_stack(api)
    cm.set_observations(o, [obs(0.0, 0.0, MAX_Z / 2, 0, 0, MAX_Z / 2)])
    cm.update_map()
    assert int((cm.get() == 254).sum()) == 21


def test_raytracing2(api):  # obstacle_tests.cpp:96-130 testRaytracing2
    cm, s, o, _ = static_stack(api)
    cm.update_map()
    assert int((cm.get() == 254).sum()) == 20
    cm.set_observations(o, [obs(9.5, 9.5, MAX_Z / 2, 0.5, 0.5, MAX_Z / 2)])
    cm.update_map()
    assert int((cm.get() == 254).sum()) == 21
    # the test then paints the layer's diagonal lethal; clearing along the ray must erase it again.  We cannot poke a
    # layer grid through the C ABI, so mark the diagonal through marking-only observations placed first.
    diag = [dict(origin=(0.5, 0.5, MAX_Z / 2), points=[(i + 0.5, i + 0.5, MAX_Z / 2) for i in range(10)],
                 obstacle_range=100.0, raytrace_range=100.0, clearing=False)]
    cm.set_observations(o, diag)
    cm.update_map()
    cm.set_observations(o, [obs(9.5, 9.5, MAX_Z / 2, 0.5, 0.5, MAX_Z / 2)])
    cm.update_map()
    g = cm.get()
    assert int((g == 254).sum()) == 21
    assert int((g == 0).sum()) == 79


def test_wave_interference(api):  # obstacle_tests.cpp:135-152
    cm = api.costmap(10, 10, 1.0, track_unknown=True)
    o = cm.add_obstacle_layer(1, True, 2.0)
    cm.set_observations(o, [obs(3.0, 3.0, MAX_Z), obs(5.0, 5.0, MAX_Z), obs(7.0, 7.0, MAX_Z)])
    cm.update_map()
    g = cm.get()
    assert (int((g == 254).sum()), int((g == 255).sum()), int((g == 0).sum())) == (3, 92, 5)


def test_z_threshold(api):  # obstacle_tests.cpp:157-172
    cm = api.costmap(10, 10, 1.0, track_unknown=True)
    o = cm.add_obstacle_layer(1, True, 2.0)
    cm.set_observations(o, [obs(0.0, 5.0, 0.4), obs(1.0, 5.0, 2.2)])
    cm.update_map()
    assert int((cm.get() == 254).sum()) == 1


def test_dynamic_obstacles_and_multiple_additions(api):  # obstacle_tests.cpp:178-222
    cm, s, o, _ = static_stack(api)
    cm.set_observations(o, [obs(0.0, 0.0)] * 3)
    cm.update_map()
    assert int((cm.get() == 254).sum()) == 21
    cm2, s2, o2, _ = static_stack(api)
    cm2.set_observations(o2, [obs(9.5, 0.0)])
    cm2.update_map()
    assert int((cm2.get() == 254).sum()) == 20


def test_adjacent_to_obstacle_can_still_move(api):  # inflation_tests.cpp:130-154
    cm = api.costmap(10, 10, 1.0)
    o = cm.add_obstacle_layer(1, True, 2.0)
    cm.add_inflation_layer(4.1, 1.0)
    cm.set_footprint(radii(2.1, 2.3))
    cm.set_observations(o, [obs(0, 0, MAX_Z)])
    cm.update_map()
    g = cm.get()  # g[y, x]
    assert g[0, 0] == 254 and g[0, 1] == 253 and g[0, 2] == 253 and g[1, 1] == 253
    assert g[0, 3] < 253 and g[1, 2] < 253
    assert int((g == 255).sum()) == 0  # testInflationShouldNotCreateUnknowns :156-175


def test_cost_function_correctness(api):  # inflation_tests.cpp:181-222
    cm = api.costmap(100, 100, 1.0)
    o = cm.add_obstacle_layer(1, True, 2.0)
    il = cm.add_inflation_layer(10.5, 1.0)
    cm.set_footprint(radii(5.0, 6.25))
    cm.set_observations(o, [obs(50, 50, MAX_Z)])
    cm.update_map()
    g = cm.get()
    for i in range(0, 6):
        assert g[50, 50 + i] >= 253 and g[50, 50 - i] >= 253 and g[50 + i, 50] >= 253 and g[50 - i, 50] >= 253
    R, costs, dists = cm.inflation_tables(il)
    assert R == 11
    assert [int(g[50, 50 + i]) for i in range(6, 12)] == [int(costs[i, 0]) for i in range(6, 12)] == [92, 34, 12, 4, 1, 0]
    assert np.array_equal(dists, np.hypot(*np.meshgrid(np.arange(13.0), np.arange(13.0), indexing="ij")))


def test_priority_queue_use_correctness(api):  # inflation_tests.cpp:249-272 (+ validatePointInflation :74-128)
    cm = api.costmap(10, 10, 1.0)
    o = cm.add_obstacle_layer(1, True, 2.0)
    il = cm.add_inflation_layer(4.1, 1.0)
    cm.set_footprint(radii(2.1, 2.3))
    cm.set_observations(o, [obs(4, 4, MAX_Z), obs(5, 5, MAX_Z)])
    cm.update_map()
    g = cm.get()
    R, costs, _ = cm.inflation_tables(il)
    for (sx, sy) in ((4, 4), (5, 5)):  # lower bound: every cell is at least the single-source cost
        for y in range(10):
            for x in range(10):
                dx, dy = abs(x - sx), abs(y - sy)
                if np.hypot(dx, dy) <= 4.1:
                    assert g[y, x] >= costs[dx, dy]


def test_inflation(api):  # inflation_tests.cpp:277-336 testInflation
    cm, s, o, i = static_stack(api, inflation=(1.0, 1.0), footprint=radii(1, 1))
    cm.update_map()
    g = cm.get()
    assert (int((g == 254).sum()), int((g == 253).sum())) == (20, 28)
    observations = [obs(0, 0, 0.4)]
    cm.set_observations(o, observations)
    cm.update_map()
    assert int((cm.get() >= 253).sum()) == 51
    observations.append(obs(2, 0))
    cm.set_observations(o, observations)
    cm.update_map()
    assert int((cm.get() >= 253).sum()) == 54
    observations.append(obs(1, 9))
    cm.set_observations(o, observations)
    cm.update_map()
    g = cm.get()
    assert g[9, 1] == 254 and g[9, 0] == 253 and g[9, 2] == 253
    observations.append(obs(0, 9))
    cm.set_observations(o, observations)
    cm.update_map()
    assert cm.get()[9, 0] == 254


def test_inflation2(api):  # inflation_tests.cpp:341-365
    cm, s, o, i = static_stack(api, inflation=(1.0, 1.0), footprint=radii(1, 1))
    cm.set_observations(o, [obs(1, 1, MAX_Z), obs(2, 1, MAX_Z), obs(2, 2, MAX_Z)])
    cm.update_map()
    g = cm.get()
    assert g[3, 2] == 253 and g[3, 3] == 253


def test_inflation3(api):  # inflation_tests.cpp:370-403
    cm = api.costmap(10, 10, 1.0)
    o = cm.add_obstacle_layer(1, True, 2.0)
    cm.add_inflation_layer(3.0, 1.0)
    cm.set_footprint(radii(1, 1.75))
    g = cm.get()
    assert int((g >= 253).sum()) == 0
    cm.set_observations(o, [obs(5, 5, MAX_Z)])
    for _ in range(2):  # idempotent on re-update
        cm.update_map()
        g = cm.get()
        assert (int((g != 0).sum()), int((g == 254).sum()), int((g == 253).sum())) == (29, 1, 4)


def test_tricky_propagation(api):  # module_tests.cpp:1055-1158 (dead test kept as a known answer: shape of the wavefront)
    # a single lethal cell inflates to a disc of every cell within R, each with the single-source table cost
    cm = api.costmap(21, 21, 1.0)
    s = cm.add_grid_layer(po.TRUE_OVERWRITE)
    il = cm.add_inflation_layer(6.0, 1.0)
    cm.set_footprint(radii(1, 1))
    g0 = np.zeros((21, 21), np.uint8)
    g0[10, 10] = 254
    cm.set_grid_layer(s, g0)
    cm.update_map()
    g = cm.get()
    R, costs, _ = cm.inflation_tables(il)
    yy, xx = np.mgrid[0:21, 0:21]
    dx, dy = np.abs(xx - 10), np.abs(yy - 10)
    expect = np.where(dx * dx + dy * dy <= R * R, costs[np.minimum(dx, R + 1), np.minimum(dy, R + 1)], 0)
    assert np.array_equal(g, expect.astype(np.uint8))


# ---- Path B known answers
def test_mapgrid_bfs_empty(api):  # base_local_planner/test/map_grid_test.cpp:137-160
    d = api.mapgrid_bfs(np.zeros((10, 10), np.uint8), [(0, 0)])
    assert (d[0, 0], d[1, 1], d[0, 4], d[4, 0], d[9, 9]) == (0.0, 2.0, 4.0, 4.0, 18.0)


def test_mapgrid_bfs_obstacles(api):  # base_local_planner/test/utest.cpp:104-166 + wavefront_map_accessor.h
    c = np.zeros((10, 10), np.uint8)
    # utest marks cells (x,y) lethal via occ_dist == 1 in MapGrid index space
    lethal = [(4, 6), (5, 6), (6, 6), (7, 6), (3, 6), (3, 7), (3, 8), (1, 1), (1, 2), (2, 2), (3, 2), (3, 1), (2, 0)]
    # our own layout (the reference's uses a legacy MapCell field): a wall with a boxed-in pocket
    for (x, y) in lethal:
        c[y, x] = 254
    d = api.mapgrid_bfs(c, [(4, 9)])
    n = 100.0
    assert d[9, 4] == 0 and d[9, 5] == 1 and d[9, 6] == 2
    assert d[6, 4] == n          # obstacle sentinel = size_x*size_y
    assert d[2, 2] == n
    assert d[1, 2] == n + 1      # boxed-in free cell (2,1): unreachable sentinel
    assert d[8, 5] == 2 and d[5, 9] == 9


def test_line_iterator(api):  # base_local_planner/test/line_iterator_test.cpp:34-76
    assert api.line_cells(0, 0, 5, 3).tolist() == [[0, 0], [1, 1], [2, 1], [3, 2], [4, 2], [5, 3]]
    assert api.line_cells(0, 0, -5, -3).tolist() == [[0, 0], [-1, -1], [-2, -1], [-3, -2], [-4, -2], [-5, -3]]
    assert api.line_cells(0, 0, 3, 5).tolist() == [[0, 0], [1, 1], [1, 2], [2, 3], [2, 4], [3, 5]]
    assert api.line_cells(2, 2, 2, 2).tolist() == [[2, 2]]


def test_velocity_iterator(api):  # base_local_planner/test/velocity_iterator_test.cpp:45-176
    np.testing.assert_allclose(api.velocity_samples(0.0, 1.0, 2), [0.0, 1.0])
    np.testing.assert_allclose(api.velocity_samples(0.0, 1.0, 3), [0.0, 0.5, 1.0])
    np.testing.assert_allclose(api.velocity_samples(-1.0, 1.0, 2), [-1.0, 0.0, 1.0])      # zero inserted
    np.testing.assert_allclose(api.velocity_samples(-1.0, 1.0, 4), [-1.0, -1 / 3, 0.0, 1 / 3, 1.0])
    np.testing.assert_allclose(api.velocity_samples(-1.0, 1.0, 3), [-1.0, 0.0, 1.0])      # zero already a sample
    np.testing.assert_allclose(api.velocity_samples(0.3, 0.3, 5), [0.3])                  # min == max
    np.testing.assert_allclose(api.velocity_samples(-1.0, -0.5, 3), [-1.0, -0.75, -0.5])
    np.testing.assert_allclose(api.velocity_samples(0.0, 1.0, 0), [0.0, 1.0])             # n clamped to 2


def test_interpret_value(api):  # static_layer.cpp:149-163 with the defaults of :70-80
    v = np.arange(256, dtype=np.uint8)
    out = api.interpret_values(v, True, 255, 100, True)
    assert out[255] == 255 and out[100] == 254 and out[99] == 0 and out[0] == 0 and out[254] == 254
    out = api.interpret_values(v, False, 255, 100, False)
    assert out[255] == 0 and out[50] == 127 and out[99] == int(99 / 100 * 254)


def test_footprint_radii(api):  # costmap_2d/test/footprint_tests.cpp expectations re-derived: square of half-width a
    i, c = api.footprint_radii([(1, 1), (1, -1), (-1, -1), (-1, 1)])
    assert i == 1.0 and abs(c - np.sqrt(2)) < 1e-15


def test_voxel_grid_basic_marking_and_clearing(api):  # voxel_grid/test/voxel_grid_tests.cpp:40-128
    """The tabletop of basicMarkingAndClearing: 11 marked lines of 4 voxels at z = 12 (44 voxels), one cleared row of
    11, a vertical column of 16 -- expressed through VoxelGrid::raytraceLine's visited voxels on a 50-wide grid."""
    size_x, table_z = 50, 12
    table = set()
    for x in range(5, 16):  # markVoxelLine(x, 0, 12, x, 3, 12)
        cells = api.voxel_line_cells(size_x, (x, 0, table_z), (x, 3, table_z))
        assert [tuple(c) for c in cells] == [(y * size_x + x, table_z) for y in range(4)]
        table |= {tuple(c) for c in cells}
    assert len(table) == 44
    row = api.voxel_line_cells(size_x, (5, 0, table_z), (15, 0, table_z))  # clearVoxelLine along the table's first row
    assert [tuple(c) for c in row] == [(x, table_z) for x in range(5, 16)]
    assert len(table - {tuple(c) for c in row}) == 33
    col = api.voxel_line_cells(size_x, (0, 0, 0), (0, 0, 15))  # clearVoxelLine(0, 0, 0, 0, 0, sizeZ - 1)
    assert [tuple(c) for c in col] == [(0, z) for z in range(16)]


def test_voxel_line_length_limit_and_diagonal(api):  # voxel_grid.h:226-297 (raytraceLine + bresenham3D)
    cells = api.voxel_line_cells(50, (2.7, 1.2, 0.5), (12.9, 6.1, 9.8))
    assert tuple(cells[0]) == (1 * 50 + 2, 0) and tuple(cells[-1]) == (6 * 50 + 12, 9) and len(cells) == 11
    short = api.voxel_line_cells(50, (2.7, 1.2, 0.5), (12.9, 6.1, 9.8), max_length=4)
    assert len(short) == int(min(1.0, 4 / np.sqrt(10.2 ** 2 + 4.9 ** 2 + 9.3 ** 2)) * 10) + 1
    assert np.array_equal(short, cells[:len(short)])


def test_scan_ingest_restatement_known_values(port):
    """oracle/scan_ingest_restated.h (the checker of navgpu_obstacle_set_scans): rays at 0 / 90 / 180 degrees land on
    (r, 0), (~0, r), (-r, ~0); NaN, negative and >= range_max ranges are dropped; +inf survives only with inf_is_valid,
    as range_max - 0.0001f (obstacle_layer.cpp:277-292); the sensor origin is the transform's translation and the height
    filter of ObservationBuffer::bufferCloud (observation_buffer.cpp:170-177) keeps min <= z <= max."""
    base = dict(angle_min=np.float32(0.0), angle_increment=np.float32(np.pi / 2), range_min=np.float32(0.1),
                range_max=np.float32(4.0), translation=(1.0, -2.0, 0.2), rotation_xyzw=(0, 0, 0, 1),
                min_obstacle_height=0.0, max_obstacle_height=2.0)
    org, a = port.project_scan(dict(base, ranges=[1.0, 2.0, 3.0, np.nan, -1.0, 4.0, np.inf], inf_is_valid=0))
    assert org == (1.0, -2.0, 0.2) and a.shape == (3, 3)
    assert a[0, 0] == 2.0 and a[0, 1] == -2.0 and a[1, 1] == 0.0 and a[2, 0] == -2.0 and np.all(a[:, 2] == np.float32(0.2))
    _, b = port.project_scan(dict(base, ranges=[np.inf, 0.05], inf_is_valid=1))
    assert b.shape == (1, 3) and b[0, 0] == np.float32(1.0) + (np.float32(4.0) - np.float32(0.0001))
    _, c = port.project_scan(dict(base, ranges=[1.0], translation=(0, 0, 2.5)))
    assert len(c) == 0  # above max_obstacle_height
    # a quarter turn about z: the ray along the sensor's x axis points along the global y axis
    q = (0.0, 0.0, np.sin(np.pi / 4), np.cos(np.pi / 4))
    _, d = port.project_scan(dict(base, ranges=[1.0], rotation_xyzw=q, translation=(0, 0, 0.2)))
    assert abs(d[0, 0]) < 1e-6 and abs(d[0, 1] - 1.0) < 1e-6


def test_tp_utest_footprint_obstacles(port):
    """The reference's own TrajectoryPlanner known answer (utest.cpp:86-110) through the CPU restatement."""
    a, b, c = sc.tp_utest_footprint_obstacles(port)
    assert a == -1.0 and b == -1.0 and c >= 0.0
