"""Synthetic "Didi-shaped" cities and demand tables.

Host-side, one-time setup (NumPy).  The per-replica order streams themselves
are generated ON DEVICE from these tables by vds_generate_orders.
"""
from math import asin, cos, radians, sin, sqrt

import numpy as np

from .engine import City

DEFAULT_BOUND = (104.011, 104.125, 30.618, 30.703)

# orders per 10-minute slot of the shipped 2016-11-01 day (UTC order), SURVEY 8d
SHIPPED_SLOT_COUNTS = np.array([
    432, 727, 782, 707, 872, 728, 555, 500, 476, 411, 435, 477, 373, 292, 277, 281, 284, 287, 210, 202, 175, 156, 232,
    242, 170, 204, 189, 156, 147, 183, 150, 147, 156, 171, 241, 242, 244, 266, 306, 336, 422, 528, 767, 923, 1182,
    1425, 1599, 1793, 1426, 1433, 1602, 1765, 2052, 2218, 1918, 1945, 2113, 2328, 2564, 2621, 1857, 1921, 1970, 2008,
    2397, 2440, 1720, 1738, 1766, 1943, 2255, 2464, 1719, 1662, 1770, 1885, 2316, 2447, 1850, 2007, 2158, 2491, 2863,
    3055, 2209, 2215, 2297, 2377, 2756, 2782, 1948, 1973, 1895, 2050, 2532, 2692, 1885, 1766, 1691, 2380, 2469, 2621,
    1880, 1929, 2018, 2340, 2630, 2438, 2003, 1951, 1656, 2090, 2471, 2246, 1845, 1825, 1848, 1900, 2149, 2232, 1518,
    1644, 1648, 1824, 2000, 2205, 1639, 1567, 1552, 1676, 1913, 2066, 1406, 1447, 1284, 1403, 1511, 1458, 1058, 959,
    910, 898, 701, 400], np.float64)


def haversine_m(lon1, lat1, lon2, lat2):
    """Great-circle distance in metres (same formula as simulator.py:266-278)."""
    lon1, lat1, lon2, lat2 = map(radians, [lon1, lat1, lon2, lat2])
    a = sin((lat2 - lat1) / 2) ** 2 + cos(lat1) * cos(lat2) * sin((lon2 - lon1) / 2) ** 2
    return 2 * asin(sqrt(a)) * 6371 * 1000


def grid_scale(bound, side_m, service_m):
    """CalculateTheScaleOfDivision (simulator.py:280-291), including its
    AverageLatitude = (N-S)/2 quirk."""
    w, e, s, n = bound
    avg_lon, avg_lat = (e - w) / 2, (n - s) / 2
    gw = int(haversine_m(w, avg_lat, e, avg_lat) / side_m + 1)
    gh = int(haversine_m(avg_lon, s, avg_lon, n) / side_m + 1)
    depth = int((service_m - 0.5 * side_m) // side_m)
    return gw, gh, depth


def grid_neighbors(gw, gh):
    """8-neighbour lists in the reference's fixed order Up, Down, Left, Right,
    LeftUp, LeftDown, RightUp, RightDown (simulator.py:463-511)."""
    off = [0]
    idx = []
    for i in range(gw * gh):
        up = i < gw * (gh - 1)
        down = i >= gw
        left = i % gw != 0
        right = (i + 1) % gw != 0
        if up: idx.append(i + gw)
        if down: idx.append(i - gw)
        if left: idx.append(i - 1)
        if right: idx.append(i + 1)
        if left and up: idx.append(i + gw - 1)
        if left and down: idx.append(i - gw - 1)
        if right and up: idx.append(i + gw + 1)
        if right and down: idx.append(i - gw + 1)
        off.append(len(idx))
    return np.array(off, np.int32), np.array(idx, np.int32)


def grid_assign(lon, lat, bound, gw, gh):
    """Node -> grid cell with the reference's strict inequalities
    (simulator.py:417-454); -1 if a node sits exactly on a cell boundary
    (the reference raises there)."""
    w, e, s, n = bound
    iw, ih = (e - w) / gw, (n - s) / gh
    cell = np.full(len(lon), -1, np.int64)
    for k in range(len(lon)):
        gx = gy = None
        for i in range(gw):
            if lon[k] > w + i * iw and lon[k] < w + (i + 1) * iw:
                gx = i
                break
        for i in range(gh):
            if lat[k] > s + i * ih and lat[k] < s + (i + 1) * ih:
                gy = i
                break
        if gx is not None and gy is not None:
            cell[k] = gw * gy + gx
    return cell


def synthetic_cost_table(lon, lat, seed=11, block=1024):
    """uint8 table indexed [end, start]: min(40, floor(5.6 min/km *
    haversine)), +1 on 1 % of off-diagonal entries, 0.18 % overwritten with
    uniform 41..79, zero diagonal."""
    n = len(lon)
    rng = np.random.default_rng(seed)
    lo, la = np.radians(lon), np.radians(lat)
    out = np.empty((n, n), np.uint8)
    for b in range(0, n, block):
        e = slice(b, min(n, b + block))
        dlon = lo[None, :] - lo[e, None]
        dlat = la[None, :] - la[e, None]
        a = np.sin(dlat / 2) ** 2 + np.cos(la[e, None]) * np.cos(la[None, :]) * np.sin(dlon / 2) ** 2
        km = 2 * np.arcsin(np.sqrt(a)) * 6371.0
        base = np.minimum(40, np.floor(5.6 * km)).astype(np.int32)
        u = rng.random(base.shape)
        base += (u < 0.01)
        tail = rng.integers(41, 80, base.shape)
        base = np.where(rng.random(base.shape) < 0.0018, tail, base)
        out[e] = base.astype(np.uint8)
    out[np.arange(n), np.arange(n)] = 0
    return out


def synthetic_grid_city(side_m=800, service_m=800, neighbor_can_server=False, n_nodes=None,
                        bound=DEFAULT_BOUND, node_seed=7, cost_seed=11):
    """192-grid (side 800 m, 4139 nodes) or 768-grid (side 400 m, 16556 nodes)
    synthetic city of SURVEY 8d; other sizes scale the 21.6 nodes/cell density."""
    gw, gh, depth = grid_scale(bound, side_m, service_m)
    if n_nodes is None:
        n_nodes = int(round(4139 * (gw * gh) / 192))
    rng = np.random.default_rng(node_seed)
    lon = np.round(rng.uniform(bound[0], bound[1], n_nodes), 7)
    lat = np.round(rng.uniform(bound[2], bound[3], n_nodes), 7)
    w, e, s, n = bound
    cell = (np.minimum(gh - 1, ((lat - s) / ((n - s) / gh)).astype(np.int64)) * gw
            + np.minimum(gw - 1, ((lon - w) / ((e - w) / gw)).astype(np.int64)))
    nb_off, nb_idx = grid_neighbors(gw, gh)
    cost = synthetic_cost_table(lon, lat, cost_seed)
    nodes = [np.nonzero(cell == c)[0] for c in range(gw * gh)]
    city = City(cost, cell, nb_off, nb_idx, depth_limit=depth, neighbor_can_server=neighbor_can_server,
                cluster_nodes=nodes)
    city.lon, city.lat, city.grid = lon, lat, (gw, gh)
    return city


def alias_tables(p):
    """Walker / Vose alias tables of the pmf p, quantised to integers: thr u32 (accept i when u1 < thr[i]) and alias u16."""
    n = len(p)
    q = np.asarray(p, np.float64) * n
    alias = np.arange(n, dtype=np.int64)
    prob = np.ones(n, np.float64)
    small = [i for i in range(n) if q[i] < 1.0]
    large = [i for i in range(n) if q[i] >= 1.0]
    while small and large:
        s_, l_ = small.pop(), large.pop()
        prob[s_] = q[s_]
        alias[s_] = l_
        q[l_] = (q[l_] + q[s_]) - 1.0
        (small if q[l_] < 1.0 else large).append(l_)
    thr = np.minimum(np.floor(prob * 4294967296.0), 4294967295.0).astype(np.uint64)
    return thr.astype(np.uint32), alias.astype(np.uint16)


class DemandTables:
    """Integer sampling tables for the on-device generator.

    slot_cdf[s][j] = floor(P(Poisson(mean_s) <= slot_base[s] + j) * 2^32)
    (last entry 0xFFFFFFFF).  Ranks follow Zipf(0.83) and are drawn with Walker's alias method from two 32-bit
    uniforms (u0, u1): i = (u0 * n_rank) >> 32; rank = i if u1 < zipf_thr[i] else zipf_alias[i]  -- two independent
    table reads instead of a 12-step dependent CDF search.  perm_pick / perm_drop map rank -> node.
    (zipf_cdf, the cumulative thresholds of the same law, is kept for reference / tests.)"""

    def __init__(self, city, orders_per_day=200_000, zipf_exponent=0.83, perm_seed=5, cdf_len=1024,
                 slot_counts=SHIPPED_SLOT_COUNTS):
        from scipy.stats import poisson
        means = np.asarray(slot_counts, np.float64) * (orders_per_day / float(np.sum(slot_counts)))
        self.n_slots = len(means)
        self.cdf_len = cdf_len
        self.means = means
        base = np.maximum(0, np.floor(means - 8 * np.sqrt(means)).astype(np.int64) - 4)
        j = np.arange(cdf_len)[None, :]
        cdf = poisson.cdf(base[:, None] + j, means[:, None])
        t = np.minimum(np.floor(cdf * 4294967296.0), 4294967295.0).astype(np.uint64)
        t[:, -1] = 0xFFFFFFFF
        self.slot_cdf = t.astype(np.uint32)
        self.slot_base = base.astype(np.int32)
        valid = city.valid_nodes()
        self.n_rank = len(valid)
        p = np.arange(1, self.n_rank + 1, dtype=np.float64) ** (-zipf_exponent)
        c = np.cumsum(p) / p.sum()
        z = np.minimum(np.floor(c * 4294967296.0), 4294967295.0).astype(np.uint64)
        z[-1] = 0xFFFFFFFF
        self.zipf_cdf = z.astype(np.uint32)
        self.zipf_thr, self.zipf_alias = alias_tables(p / p.sum())
        rng = np.random.default_rng(perm_seed)
        self.perm_pick = valid[rng.permutation(self.n_rank)].astype(np.uint16)
        self.perm_drop = valid[rng.permutation(self.n_rank)].astype(np.uint16)
        self.mean_total = float(means.sum())
        self.max_orders = int(self.mean_total + 8 * np.sqrt(self.mean_total)) + 64
        self.max_orders_per_tick = int((self.slot_base + cdf_len).max())
        self.ticks = self.n_slots + 4       # = floor(m_last/10) + 5 with m_last in the last slot

    def to_device(self, device):
        import torch
        return {k: torch.from_numpy(getattr(self, k)).to(device)
                for k in ("slot_cdf", "slot_base", "zipf_thr", "zipf_alias", "perm_pick", "perm_drop")}
