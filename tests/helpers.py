"""Shared parity-test plumbing: drive the CUDA engine and the CPU oracle in
lock-step on the same inputs and compare every observable bit-for-bit."""
import numpy as np

from oracle.oracle import Oracle, PARITY_THRESHOLD
from oracle.ref_harness import dispatch_decision


def make_oracle(city, V, minute, pickup, delivery, period=10, threshold=PARITY_THRESHOLD):
    return Oracle(city.cost_u8, city.node2cluster.astype(np.int32), city.nb_off.astype(np.int32),
                  city.nb_idx.astype(np.int32), minute, pickup, delivery, V, period,
                  city.depth_limit, city.neighbor_can_server, threshold)


def random_orders(city, n, rng, n_minutes=200, hot=0.6):
    """Sorted synthetic order stream with hot spots (ties + contention)."""
    valid = city.valid_nodes().astype(np.int32)
    minute = np.sort(rng.integers(0, n_minutes, n)).astype(np.int32)
    hotset = rng.choice(valid, size=max(1, len(valid) // 20), replace=False)
    pick = np.where(rng.random(n) < hot, rng.choice(hotset, n), rng.choice(valid, n)).astype(np.int32)
    drop = np.where(rng.random(n) < hot, rng.choice(hotset, n), rng.choice(valid, n)).astype(np.int32)
    return minute, pick, drop


def policy_moves(oracle, city, tick):
    """The shared deterministic dispatch policy (oracle/ref_harness.py),
    evaluated on the oracle's idle lists: cluster-ID order, list order."""
    veh, node = [], []
    for c in range(city.n_clusters):
        nb = [int(x) for x in city.neighbors(c) if len(city.cluster_nodes[int(x)])]
        for v in oracle.idle_list(c):
            k = dispatch_decision(tick, int(v), len(nb))
            if k < 0:
                continue
            nodes = city.cluster_nodes[nb[k]]
            veh.append(int(v))
            node.append(int(nodes[(tick + int(v)) % len(nodes)]))
    return np.array(veh, np.int32), np.array(node, np.int32)


def lockstep(engine, oracles, loc0, dispatch=False, check_lists_every=1, ticks=None, fused=False):
    """oracles: one per replica.  Returns the number of ticks compared.
    fused=True drives the engine through the one-launch-per-tick path
    (vds_tick) instead of update / match / supply_expect."""
    city = engine.city
    R = engine.R
    engine.reset(loc0)
    for r, o in enumerate(oracles):
        o.reset(loc0 if np.ndim(loc0) == 1 else loc0[r])
    T = engine.T if ticks is None else ticks
    for k in range(T):
        assert not oracles[0].done()
        if not fused:
            engine.update(k)
        for o in oracles:
            o.update()
        if not fused and check_lists_every and k % check_lists_every == 0:
            for r in (0, R - 1):
                got = engine.idle_lists(r)
                exp = oracles[r].idle_lists()
                for c in range(city.n_clusters):
                    assert np.array_equal(got[c], exp[c]), f"idle list order tick {k} replica {r} cluster {c}"
        if fused:
            engine.tick(k)
        else:
            engine.match(k)
            engine.supply_expect(k)
        for o in oracles:
            o.match(); o.supply_expect(); o.snapshot_pre_dispatch()
        if fused and check_lists_every and k % check_lists_every == 0:
            for r in (0, R - 1):
                got = engine.idle_lists(r)
                exp = oracles[r].idle_lists()
                for c in range(city.n_clusters):
                    assert np.array_equal(got[c], exp[c]), f"idle list order after match, tick {k} replica {r} cluster {c}"
        t = {n: engine.tensors[n].cpu().numpy() for n in ("per_match", "per_dispatch", "supply", "n_orders", "idle_live")}
        for r, o in enumerate(oracles):
            assert np.array_equal(t["per_match"][r], o.per_match()), f"per_match tick {k} replica {r}"
            assert np.array_equal(t["n_orders"][r], o.n_orders()), f"n_orders tick {k} replica {r}"
            assert np.array_equal(t["per_dispatch"][r], o.per_dispatch()), f"per_dispatch tick {k} replica {r}"
            assert np.array_equal(t["supply"][r], o.supply()), f"supply tick {k} replica {r}"
        if dispatch:
            off, mv, mn = [0], [], []
            for r, o in enumerate(oracles):
                v, n = policy_moves(o, city, k)
                assert o.dispatch(v, n) == len(v)
                mv.append(v); mn.append(n); off.append(off[-1] + len(v))
            engine.dispatch(k, np.array(off, np.int32), np.concatenate(mv), np.concatenate(mn))
        for o in oracles:
            o.end_tick()
        later = engine.tensors["idle_live"].cpu().numpy()
        for r, o in enumerate(oracles):
            assert np.array_equal(later[r], o.later_dispatch()), f"later_dispatch tick {k} replica {r}"
    if ticks is None:
        assert oracles[0].done()
    st = engine.stats().cpu().numpy()
    for r, o in enumerate(oracles):
        veh, wait, delta = engine.order_results(r)
        assert np.array_equal(veh, o.order_vehicle()), f"matched vehicle ids replica {r}"
        assert np.array_equal(wait, o.order_wait()), f"wait replica {r}"
        os_ = o.stats()
        assert tuple(st[r][:9]) == tuple(os_[:9]), f"stats replica {r}: {st[r]} vs {os_}"
    return T


def rollout_vs_oracle(engine, oracles, loc0, windows=None):
    """Fused replica-resident rollout (engine built with trace=True) vs the
    oracle: per-tick per-cluster trace, idle-list order at every window
    boundary, per-order results and counters.  windows: list of tick counts
    (default one window = the whole episode)."""
    city, R, T, C = engine.city, engine.R, engine.T, engine.city.n_clusters
    assert engine.fused and engine.trace is not None
    engine.reset(loc0)
    for r, o in enumerate(oracles):
        o.reset(loc0 if np.ndim(loc0) == 1 else loc0[r])
    exp = np.zeros((R, T, 4, C), np.int32)
    k = 0
    for wlen in (windows or [T]):
        engine.rollout(k, wlen)
        for kk in range(k, k + wlen):
            for r, o in enumerate(oracles):
                o.update(); o.match(); o.supply_expect(); o.snapshot_pre_dispatch()
                exp[r, kk, 0], exp[r, kk, 1] = o.per_match(), o.per_dispatch()
                exp[r, kk, 2], exp[r, kk, 3] = o.supply(), o.n_orders()
                o.end_tick()
        k += wlen
        t = {n: engine.tensors[n].cpu().numpy() for n in ("per_match", "per_dispatch", "supply", "n_orders", "idle_live")}
        for r, o in enumerate(oracles):
            assert np.array_equal(t["per_match"][r], exp[r, k - 1, 0]), f"per_match after window ending {k}"
            assert np.array_equal(t["per_dispatch"][r], exp[r, k - 1, 1])
            assert np.array_equal(t["supply"][r], exp[r, k - 1, 2])
            assert np.array_equal(t["n_orders"][r], exp[r, k - 1, 3])
            assert np.array_equal(t["idle_live"][r], o.later_dispatch())
        for r in (0, R - 1):
            got, want = engine.idle_lists(r), oracles[r].idle_lists()
            for c in range(C):
                assert np.array_equal(got[c], want[c]), f"idle list order after tick {k - 1} replica {r} cluster {c}"
            V = engine.V
            assert np.array_equal(engine.tensors["veh_loc"][r, :V].cpu().numpy().astype(np.int32), oracles[r].veh_loc())
            dest = engine.tensors["veh_dest"][r, :V].cpu().numpy().astype(np.int32)
            assert np.array_equal(np.where(dest == 0xFFFF, -1, dest), oracles[r].veh_dest())
    assert k == T and oracles[0].done()
    tr = engine.trace.cpu().numpy()
    for r in range(R):
        for f, name in enumerate(("per_match", "per_dispatch", "supply", "n_orders")):
            bad = np.argwhere(tr[r, :, f] != exp[r, :, f])
            assert len(bad) == 0, f"trace {name} replica {r}: first mismatch (tick, cluster) = {bad[0]}"
    st = engine.stats().cpu().numpy()
    for r, o in enumerate(oracles):
        veh, wait, delta = engine.order_results(r)
        assert np.array_equal(veh, o.order_vehicle()), f"matched vehicle ids replica {r}"
        assert np.array_equal(wait, o.order_wait()), f"wait replica {r}"
        os_ = o.stats()
        if city.neighbor_can_server and city.depth_limit > 0:
            # the speculative search path counts the candidates IT examined (>= the reference's lookups)
            assert tuple(st[r][:6]) == tuple(os_[:6]) and tuple(st[r][7:9]) == tuple(os_[7:9]), f"stats replica {r}: {st[r]} vs {os_}"
            assert st[r][6] >= os_[6]
        else:
            assert tuple(st[r][:9]) == tuple(os_[:9]), f"stats replica {r}: {st[r]} vs {os_}"
    return T


def engine_replica_orders(eng, r, n_slots):
    """(minute, pickup, delivery) of replica r's device-resident order stream
    (synthetic streams: slot s == tick s+1 == minute 10*s)."""
    ro = 0 if eng.OR == 1 else r
    n = int(eng.n_orders_total[ro].item())
    pd = eng.order_pd[ro, :n].cpu().numpy().astype(np.uint32)
    toff = eng.tick_off[ro].cpu().numpy()
    tick = np.searchsorted(toff, np.arange(n), side="right") - 1
    tick = np.clip(tick, 1, n_slots)
    minute = (10 * (tick - 1)).astype(np.int32)
    return minute, (pd & 0xFFFF).astype(np.int32), (pd >> 16).astype(np.int32)


# ------------------------------------------------------------ golden fixtures
import os as _os

GOLDEN = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")
SMALL_CASES = ("grid_d0", "grid_d2", "kmeans_d1", "grid_d0_dispatch", "grid_d1_dispatch", "grid_d1_focus")
REAL_CASES = ("kmeans", "grid6000", "grid5000d3", "grid400v10000")


def load_golden(name, real=False):
    """inputs + traces recorded from the unmodified reference (make_golden.py)."""
    if real:
        p = _os.path.join(GOLDEN, "_real", f"real_{name}.npz")
        if not _os.path.exists(p):
            return None
        z = dict(np.load(p))
        z["in_cost_u8"] = np.load(_os.path.join(GOLDEN, "_real", "real_cost.npz"))["cost_u8"]
        return z
    return dict(np.load(_os.path.join(GOLDEN, f"small_{name}.npz")))


def golden_city(z):
    from vehicles_dispatch_simulator_b200.engine import City
    p, depth, ncs, C, V, disp = [int(x) for x in z["in_params"]]
    nodes = [z["in_cl_nodes"][z["in_cl_nodes_off"][c]:z["in_cl_nodes_off"][c + 1]] for c in range(C)]
    return City(z["in_cost_u8"], z["in_node2cluster"], z["in_nb_off"], z["in_nb_idx"], depth_limit=depth,
                neighbor_can_server=bool(ncs), cluster_nodes=nodes), V, p


def golden_idle_lists(z, k, C):
    off, flat = z["tr_idle_off"], z["tr_idle_flat"]
    return [flat[off[k * C + c]:off[k * C + c + 1]] for c in range(C)]


def check_oracle_against_golden(z):
    """CPU oracle vs the reference's recorded trace, every observable."""
    city, V, p = golden_city(z)
    C = city.n_clusters
    o = make_oracle(city, V, z["in_order_minute"], z["in_order_pickup"], z["in_order_delivery"], p)
    o.reset(z["in_veh_loc0"])
    mo, mv = z["tr_moves_off"], z["tr_moves"]
    k = 0
    while not o.done():
        o.update()
        exp = golden_idle_lists(z, k, C)
        for c in range(C):
            assert np.array_equal(o.idle_list(c), exp[c]), f"idle list order tick {k} cluster {c}"
        o.match()
        assert np.array_equal(o.per_match(), z["tr_per_match"][k])
        assert np.array_equal(o.n_orders(), z["tr_n_orders"][k])
        o.supply_expect(); o.snapshot_pre_dispatch()
        assert np.array_equal(o.supply(), z["tr_supply"][k]), f"supply tick {k}"
        assert np.array_equal(o.per_dispatch(), z["tr_per_dispatch"][k])
        m = mv[mo[k]:mo[k + 1]]
        if len(m):
            assert o.dispatch(m[:, 0], m[:, 1]) == len(m)
        o.end_tick()
        assert np.array_equal(o.later_dispatch(), z["tr_later_dispatch"][k])
        assert tuple(o.stats()[:3]) == tuple(z["tr_counters"][k])
        k += 1
    assert k == int(z["tr_final"][6])
    assert np.array_equal(o.order_vehicle(), z["tr_order_vehicle"])
    assert np.array_equal(o.order_wait(), z["tr_order_wait"])
    assert tuple(o.stats()[:6]) == tuple(z["tr_final"][:6])
    return k


def check_engine_against_golden(engine, z, replicas=(0,), lists_every=1):
    """CUDA engine (already bound to the golden inputs) vs the reference's trace."""
    city = engine.city
    C = city.n_clusters
    T = int(z["tr_final"][6])
    assert engine.T == T
    engine.reset(z["in_veh_loc0"])
    mo, mv = z["tr_moves_off"], z["tr_moves"]
    for k in range(T):
        engine.update(k)
        if lists_every and k % lists_every == 0:
            exp = golden_idle_lists(z, k, C)
            for r in replicas:
                got = engine.idle_lists(r)
                for c in range(C):
                    assert np.array_equal(got[c], exp[c]), f"idle list order tick {k} cluster {c}"
        engine.match(k)
        engine.supply_expect(k)
        t = {n: engine.tensors[n].cpu().numpy() for n in ("per_match", "per_dispatch", "supply", "n_orders")}
        for r in replicas:
            assert np.array_equal(t["per_match"][r], z["tr_per_match"][k]), f"per_match tick {k}"
            assert np.array_equal(t["n_orders"][r], z["tr_n_orders"][k]), f"n_orders tick {k}"
            assert np.array_equal(t["supply"][r], z["tr_supply"][k]), f"supply tick {k}"
            assert np.array_equal(t["per_dispatch"][r], z["tr_per_dispatch"][k]), f"per_dispatch tick {k}"
        m = mv[mo[k]:mo[k + 1]]
        if len(m):
            off = np.arange(engine.R + 1, dtype=np.int32) * len(m)
            engine.dispatch(k, off, np.tile(m[:, 0], engine.R), np.tile(m[:, 1], engine.R))
        later = engine.tensors["idle_live"].cpu().numpy()
        for r in replicas:
            assert np.array_equal(later[r], z["tr_later_dispatch"][k]), f"later_dispatch tick {k}"
    st = engine.stats().cpu().numpy()
    for r in replicas:
        veh, wait, _ = engine.order_results(r)
        assert np.array_equal(veh, z["tr_order_vehicle"]), "matched vehicle ids"
        assert np.array_equal(wait, z["tr_order_wait"]), "wait"
        assert tuple(st[r][:6]) == tuple(z["tr_final"][:6]), f"{st[r]} vs {z['tr_final']}"
    return T
