"""NumPy restatement of the on-device synthetic generator (Philox4x32-10;
inverse-CDF sampling of the Poisson slot counts, alias sampling of the Zipf ranks).  TEST INFRASTRUCTURE ONLY: used by tests/ and bench.py's
CPU legs to give the oracle the same order streams / placements the GPU
generates.  Philox4x32-10 is the published Random123 algorithm (Salmon et al.,
SC'11); known-answer vectors from the Random123 distribution are checked in
tests/test_synthetic.py."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over the counter words (uint64 arrays holding 32-bit values)."""
    c0, c1, c2, c3 = [np.asarray(c, np.uint64) & MASK for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def cdf_search(cdf, u):
    """smallest j with u < cdf[j] (last threshold catches everything)."""
    j = np.searchsorted(cdf, u, side="right")
    return np.minimum(j, len(cdf) - 1)


def alias_draw(tables, u0, u1):
    """rank = i if u1 < thr[i] else alias[i], i = (u0 * n_rank) >> 32  (Walker's alias method, integer arithmetic)."""
    i = ((np.asarray(u0, np.uint64) * np.uint64(tables.n_rank)) >> np.uint64(32)).astype(np.int64)
    return np.where(np.asarray(u1, np.uint64) < tables.zipf_thr[i].astype(np.uint64), i, tables.zipf_alias[i].astype(np.int64))


def replica_orders(tables, seed, g):
    """(minute, pickup, delivery) of global replica g, in generation order."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    g0, g1 = g & 0xFFFFFFFF, (g >> 32) & 0xFFFFFFFF
    s = np.arange(tables.n_slots, dtype=np.uint64)
    x0, _, _, _ = philox4x32_10(s, 0, g0, g1, k0, k1)
    counts = np.array([tables.slot_base[i] + cdf_search(tables.slot_cdf[i], x0[i]) for i in range(tables.n_slots)], np.int64)
    minute, pick, drop = [], [], []
    for sl in range(tables.n_slots):
        n = int(counts[sl])
        if n == 0:
            continue
        i = np.arange(n, dtype=np.uint64)
        a, b, c, d = philox4x32_10(i, sl | (1 << 16), g0, g1, k0, k1)
        pick.append(tables.perm_pick[alias_draw(tables, a, b)])
        drop.append(tables.perm_drop[alias_draw(tables, c, d)])
        minute.append(np.full(n, 10 * sl, np.int32))
    return (np.concatenate(minute), np.concatenate(pick).astype(np.int32), np.concatenate(drop).astype(np.int32))


def replica_placement(valid_nodes, n_vehicles, seed, g):
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    g0, g1 = g & 0xFFFFFFFF, (g >> 32) & 0xFFFFFFFF
    v = np.arange(n_vehicles, dtype=np.uint64)
    x0, _, _, _ = philox4x32_10(v, 2 << 16, g0, g1, k0, k1)
    return np.asarray(valid_nodes)[((x0 * np.uint64(len(valid_nodes))) >> np.uint64(32)).astype(np.int64)].astype(np.int32)


def random_policy_moves(city, idle_vehicles, veh_cluster, tick, seed, g, prob):
    """Host restatement of policy_random_kernel: (vehicle, node) moves of global replica g at `tick`, in vehicle
    index order.  idle_vehicles: ascending vehicle indices that are idle after the match; veh_cluster[v]."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    g0, g1 = g & 0xFFFFFFFF, (g >> 32) & 0xFFFFFFFF
    q32 = min(0xFFFFFFFF, int(prob * 4294967296.0))
    v = np.asarray(idle_vehicles, dtype=np.uint64)
    if len(v) == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    x0, x1, x2, _ = philox4x32_10(v, (3 << 16) | tick, g0, g1, k0, k1)
    veh, node = [], []
    for i in range(len(v)):
        if int(x0[i]) >= q32:
            continue
        c = int(veh_cluster[int(v[i])])
        nb = city.neighbors(c)
        if len(nb) == 0:
            continue
        tc = int(nb[(int(x1[i]) * len(nb)) >> 32])
        nodes = city.cluster_nodes[tc]
        if len(nodes) == 0:
            continue
        veh.append(int(v[i])); node.append(int(nodes[(int(x2[i]) * len(nodes)) >> 32]))
    return np.array(veh, np.int32), np.array(node, np.int32)
