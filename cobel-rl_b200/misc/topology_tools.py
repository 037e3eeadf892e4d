"""Topology-graph builders (reference: misc/topology_tools.py:14-172, 275-373).

Same node dictionaries as the reference (``{'id','pose','terminal','reward','neighbors'}``,
ids ``str(n)``, neighbour order left / up / right / down, missing neighbours point to the
node itself), built from integer lattice arithmetic instead of all-pairs distance scans.
"""
import numpy as np


def _lattice_nodes(cells, spacing):
    """Nodes on integer lattice cells ``[(x, y), ...]`` with 4-neighbourhoods (up = larger y)."""
    index = {c: str(n) for n, c in enumerate(cells)}
    nodes = {}
    for n, (x, y) in enumerate(cells):
        nid = str(n)
        nb = [index.get((x - 1, y), nid), index.get((x, y + 1), nid),
              index.get((x + 1, y), nid), index.get((x, y - 1), nid)]
        nodes[nid] = {'id': nid, 'pose': (float(x) * spacing, float(y) * spacing, 0.0, 0.0, 0.0, 0.0),
                      'terminal': False, 'reward': 0.0, 'neighbors': nb}
    return nodes


def linear_track(nb_nodes_track, nb_nodes_width, spacing=1.0, reward=1.0, location='right'):
    """topology_tools.py:14-100: nodes numbered row-major ``n = j*L + i``; reward and terminal
    flag on the ``location`` end of every row, starting nodes on the opposite end."""
    assert nb_nodes_track > 1, 'Track has to be at least 2 states long!'
    assert nb_nodes_width > 0, 'Track has to be at least 1 state wide!'
    assert spacing > 0, 'Node spacing must be positive!'
    assert location in ['left', 'right'], 'Invalid reward location!'
    L, Wd = nb_nodes_track, nb_nodes_width
    nodes = _lattice_nodes([(i, Wd - j - 1) for j in range(Wd) for i in range(L)], spacing)
    starting_nodes = []
    for j in range(Wd):
        goal = str(j * L + (L - 1) * (location == 'right'))
        nodes[goal].update({'terminal': True, 'reward': reward})
        starting_nodes.append(str(j * L + (L - 1) * (location == 'left')))
    return nodes, starting_nodes


def grid(nb_nodes, limits=(0.0, 1.0), reward=1.0, location=None):
    """topology_tools.py:103-172: rectangular grid graph; the goal defaults to the top-right node."""
    nx = nb_nodes if isinstance(nb_nodes, int) else nb_nodes[0]
    ny = nb_nodes if isinstance(nb_nodes, int) else nb_nodes[1]
    assert (nx > 1 and ny >= 1) or (nx >= 1 and ny > 1), 'Invalid environment dimensions!'
    lim_x = limits if isinstance(limits[0], float) else limits[0]
    lim_y = limits if isinstance(limits[0], float) else limits[1]
    assert lim_x[1] > lim_x[0], 'Invalid x coordinate range!'
    assert lim_y[1] > lim_y[0], 'Invalid y coordinate range!'
    xs, ys = np.linspace(lim_x[0], lim_x[1], nx), np.linspace(lim_y[0], lim_y[1], ny)
    nodes = {}
    for n in range(nx * ny):
        j, i = divmod(n, nx)
        nb = [str(j * nx + max(i - 1, 0)), str(max(j - 1, 0) * nx + i),
              str(j * nx + min(i + 1, nx - 1)), str(min(j + 1, ny - 1) * nx + i)]
        pose = (float(xs[i]), float(lim_y[1] - (ys[j] - lim_y[0])), 0.0, 0.0, 0.0, 0.0)
        nodes[str(n)] = {'id': str(n), 'pose': pose, 'terminal': False, 'reward': 0.0, 'neighbors': nb}
    if location is None or location not in nodes:
        location = str(nx - 1)
    nodes[location].update({'terminal': True, 'reward': reward})
    starting_nodes = [n for n in nodes if n != location]
    return nodes, starting_nodes


def t_maze(nb_nodes_stem, nb_nodes_arm, nb_nodes_width, spacing=1.0, reward=1.0, location='right'):
    """topology_tools.py:275-373: arms first (row-major from the top), then the stem; goal at
    the end of the ``location`` arm, starting nodes = bottom row of the stem."""
    assert nb_nodes_arm > 0, 'Invalid arm length!'
    assert nb_nodes_stem > 0, 'Invalid stem length!'
    assert nb_nodes_width > 0, 'Invalid corridor width!'
    assert location in ['left', 'right'], 'The goal can only be located left or right!'
    span = nb_nodes_arm * 2 + nb_nodes_width
    cells = [(j, nb_nodes_stem + nb_nodes_width - 1 - i) for i in range(nb_nodes_width) for j in range(span)]
    cells += [(nb_nodes_arm + j, nb_nodes_stem - 1 - i) for i in range(nb_nodes_stem) for j in range(nb_nodes_width)]
    nodes = _lattice_nodes(cells, spacing)
    for i in range(nb_nodes_width):
        nodes[str(span * i + (span - 1) * int(location == 'right'))].update({'terminal': True, 'reward': reward})
    starting_nodes = list(nodes.keys())[-nb_nodes_width:]
    return nodes, starting_nodes


def hexagonal(nb_nodes, limits=(0.0, 1.0), reward=1.0, location=None):
    """topology_tools.py:175-272: hexagonal lattice (odd rows shifted by half a spacing, nodes beyond
    the upper x limit dropped) with 6 neighbour slots.  A neighbour is any node closer than 1.5
    spacings; it is put into the slot whose nominal direction (in the reference's slot order
    240, 300, 0, 60, 120, 180 degrees, first minimum of the absolute angular difference without
    wrap-around) is closest to its bearing, empty slots point to the node itself, and the slot
    list is finally reversed -- all as in the reference, including its float behaviour."""
    assert nb_nodes > 1, 'Invalid number of nodes!'
    assert limits[1] > limits[0], 'Invalid coordinate range!'
    spacing = (limits[1] - limits[0]) / (nb_nodes - 1)
    ticks = np.linspace(limits[0], limits[1], nb_nodes)
    xy = np.array([[x, y] for y in ticks for x in ticks])
    odd = np.repeat(np.arange(nb_nodes) % 2 == 1, nb_nodes)
    xy[odd, 0] += spacing / 2
    xy = xy[xy[:, 0] <= limits[1]]
    pose_xy = [(x, limits[1] - (y - limits[0])) for x, y in xy]
    n = len(pose_xy)
    P = np.array(pose_xy)
    dist = np.sqrt(((P[:, None, :] - P[None, :, :]) ** 2).sum(axis=2))
    slots = np.array([(i * 60 - 120) % 360 for i in range(6)])
    nodes = {}
    for a in range(n):
        nid = str(a)
        assigned = [nid] * 6
        # the reference also runs its own id (six times) through the slot search: bearing 0 -> slot of 0 degrees
        candidates = [a] * 6 + [b for b in range(n) if b != a and dist[a, b] < spacing * 1.5]
        for b in candidates:
            bearing = np.angle(complex(P[b, 0] - P[a, 0], P[b, 1] - P[a, 1]), deg=True) % 360
            assigned[int(np.argmin(np.abs(slots - bearing)))] = str(b)
        nodes[nid] = {'id': nid, 'pose': (pose_xy[a][0], pose_xy[a][1], 0.0, 0.0, 0.0, 0.0), 'terminal': False,
                      'reward': 0.0, 'neighbors': assigned[::-1]}
    if location is None or location not in nodes:
        location = str(nb_nodes - 1)
    nodes[location].update({'terminal': True, 'reward': reward})
    starting_nodes = [k for k in nodes if k != location]
    return nodes, starting_nodes


def cross(nb_nodes_arm, nb_nodes_width, spacing=1.0, rotation=0.0):
    """topology_tools.py:376-469: plus-shaped maze (top arm, middle band, bottom arm), no goal; every
    node is a starting node; poses are centred, scaled by ``spacing`` and rotated by ``rotation`` degrees."""
    assert nb_nodes_arm > 0, 'The arm must be at least 1 node long!'
    assert nb_nodes_width > 0, 'The corridors must be at least 1 node wide!'
    assert spacing > 0, 'Node spacing must be positive!'
    arm, wid = nb_nodes_arm, nb_nodes_width
    D = 2 * arm + wid
    cells = [(arm + j, D - 1 - i) for i in range(arm) for j in range(wid)]
    cells += [(i, D - 1 - arm - j) for j in range(wid) for i in range(D)]
    cells += [(arm + j, D - 1 - (i + arm + wid)) for i in range(arm) for j in range(wid)]
    nodes = _lattice_nodes(cells, 1.0)
    ticks = np.linspace(0, 1.0, D)
    scale = (D - 1) * spacing
    offset = spacing * (D - 1) / 2
    theta = np.deg2rad(rotation)
    R = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
    for (ix, iy), node in zip(cells, nodes.values()):
        x, y = R @ (np.array((ticks[ix], ticks[iy])) * scale - offset)
        node['pose'] = (float(x), float(y), 0.0, 0.0, 0.0, 0.0)
    return nodes, list(nodes.keys())


# ---------------------------------------------------------------------------
# remove_obstructed_neighbors (topology_tools.py:472-505).  The reference delegates the geometry to shapely
# (intersects(LineString, buffer(MultiPolygon, d))); shapely is not part of this image, so the predicate is written
# out: an edge is obstructed if its segment comes within ``buffer_distance`` of an obstacle polygon (closed set;
# d = 0: touches or crosses it).  shapely's buffer approximates the rounded corners of the offset region by 8
# segments per quarter circle, so the two predicates can differ only for an edge that passes a convex obstacle
# corner at a distance within 0.5 % below ``buffer_distance``.
# ---------------------------------------------------------------------------
def _rings(obstacle):
    """[exterior, hole, ...] vertex arrays of an obstacle: an ``[n, 2]`` vertex list, or any object with shapely's
    ``exterior.coords`` / ``interiors`` attributes."""
    if hasattr(obstacle, 'exterior'):
        rings = [np.asarray(obstacle.exterior.coords, dtype=np.float64)[:, :2]]
        rings += [np.asarray(r.coords, dtype=np.float64)[:, :2] for r in getattr(obstacle, 'interiors', [])]
    else:
        rings = [np.asarray(obstacle, dtype=np.float64).reshape(-1, 2)]
    return [r[:-1] if len(r) > 1 and np.array_equal(r[0], r[-1]) else r for r in rings]


def _inside(pt, ring):
    """Even-odd rule; points on the boundary are caught by the distance test."""
    x, y = pt
    x0, y0 = ring[:, 0], ring[:, 1]
    x1, y1 = np.roll(x0, -1), np.roll(y0, -1)
    cross = (y0 > y) != (y1 > y)
    with np.errstate(divide='ignore', invalid='ignore'):
        xs = x0 + (y - y0) * (x1 - x0) / (y1 - y0)
    return bool(np.count_nonzero(cross & (x < xs)) % 2)


def _point_segment_distance(p, a, b):
    ab = b - a
    den = float(ab @ ab)
    t = 0.0 if den == 0.0 else min(1.0, max(0.0, float((p - a) @ ab) / den))
    return float(np.linalg.norm(p - (a + t * ab)))


def _segments_cross(p1, p2, q1, q2):
    def orient(a, b, c):
        return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
    d1, d2, d3, d4 = orient(q1, q2, p1), orient(q1, q2, p2), orient(p1, p2, q1), orient(p1, p2, q2)
    return ((d1 > 0) != (d2 > 0)) and ((d3 > 0) != (d4 > 0)) and d1 != 0 and d2 != 0 and d3 != 0 and d4 != 0


def _segment_polygon_distance(p1, p2, rings):
    """0 if the segment touches, crosses or lies inside the polygon (exterior minus holes), else the distance."""
    inside = _inside(p1, rings[0]) and not any(_inside(p1, h) for h in rings[1:])
    if inside:
        return 0.0
    best = np.inf
    for ring in rings:
        for a, b in zip(ring, np.roll(ring, -1, axis=0)):
            if _segments_cross(p1, p2, a, b):
                return 0.0
            best = min(best, _point_segment_distance(p1, a, b), _point_segment_distance(p2, a, b),
                       _point_segment_distance(a, p1, p2), _point_segment_distance(b, p1, p2))
    return best


def remove_obstructed_neighbors(nodes, obstacles, buffer_distance=0.0):
    """topology_tools.py:472-505: edges whose straight line comes within ``buffer_distance`` of an obstacle are
    redirected to the node itself.  ``obstacles``: polygons as ``[n, 2]`` vertex lists (or shapely-like objects);
    the z-coordinate is ignored, as in the reference."""
    import copy
    assert buffer_distance >= 0, 'The buffer distance has to be non-negative!'
    nodes_updated = copy.deepcopy(nodes)
    polys = [_rings(o) for o in obstacles]
    for n, node in nodes_updated.items():
        pos_1 = np.array(node['pose'], dtype=np.float64)[:2]
        for i, neighbor in enumerate(node['neighbors']):
            pos_2 = np.array(nodes_updated[neighbor]['pose'][:2], dtype=np.float64)
            if any(_segment_polygon_distance(pos_1, pos_2, rings) <= buffer_distance for rings in polys):
                node['neighbors'][i] = n
    return nodes_updated
