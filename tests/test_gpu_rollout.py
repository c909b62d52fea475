"""-m gpu parity tests of the replica-resident fused rollout kernel
(csrc/rollout.cuh) through the C ABI: vs the CPU oracle (per-tick trace, idle
list order, per-order results, counters), vs the per-phase kernels (HBM state
bit-for-bit), and on the golden fixtures recorded from the reference."""
import numpy as np
import pytest

from tests.helpers import (SMALL_CASES, golden_city, load_golden, lockstep, make_oracle, random_orders,
                           rollout_vs_oracle)

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["node_queues", "slot_lists"])
def _fused_search(monkeypatch, request):
    """the speculative replica-resident search variant is opt-in (read by vds_create); these tests cover it.  Every
    test runs twice: with the node-queue rollout kernel (csrc/rollout_nq.cuh, opt-in)
    and with the per-cluster slot-list kernel (csrc/rollout.cuh) that is the default (the default; VDS_NQ=1 opts into the former)."""
    monkeypatch.setenv("VDS_FUSED_SEARCH", "1")
    if request.param == "node_queues":
        monkeypatch.setenv("VDS_NQ", "1")


def _city(side=800, service=800, ncs=False, n_nodes=700):
    from vehicles_dispatch_simulator_b200.synthetic import synthetic_grid_city
    return synthetic_grid_city(side_m=side, service_m=service, neighbor_can_server=ncs, n_nodes=n_nodes)


def _engine(city, V, minute, pickup, delivery, R=1, period=10, **kw):
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine, tick_offsets
    off, T = tick_offsets(minute, period)
    e = DispatchEngine(city, V, replicas=R, ticks=T, period=period, max_orders=len(minute),
                       max_orders_per_tick=max(64, int(np.diff(off).max())), **kw)
    e.bind_shared_orders(minute, pickup, delivery)
    return e


@pytest.mark.parametrize("V,n_orders,windows", [
    (150, 3000, None),                 # contended: most orders rejected
    (600, 6000, [1, 1, 3, 0]),         # window boundaries incl. single fused ticks
    (2500, 4000, None),                # Vp > 2048, idle lists > 32 (general scan path), 256-thread CTAs
    (9000, 5000, None),                # one replica = 126 KB of shared memory: one per SM, 1024-thread CTAs
    (2500, 3000, [2, 4, 5, 0]),        # 256-thread CTAs: windows on both sides of VDS_TMA_MAX_TICKS (TMA / register-path variant)
    (9000, 3000, [3, 0]),              # 1024-thread CTAs: a TMA window, then the register-path variant
])
def test_fused_rollout_vs_oracle(cuda_device, V, n_orders, windows):
    rng = np.random.default_rng(V)
    city = _city()
    minute, pick, drop = random_orders(city, n_orders, rng)
    loc0 = rng.choice(city.valid_nodes(), (2, V)).astype(np.int32)
    e = _engine(city, V, minute, pick, drop, R=2, trace=True)
    if windows:
        windows = list(windows)
        windows[-1] = e.T - sum(windows)
    rollout_vs_oracle(e, [make_oracle(city, V, minute, pick, drop) for _ in range(2)], loc0, windows)


@pytest.mark.parametrize("service,V,n_orders,windows", [
    (1600, 200, 5000, None),           # depth 1, starved: most orders search and reject
    (2800, 400, 5000, [2, 1, 0]),      # depth 3, window boundaries
    (2800, 1500, 5000, None),          # depth 3, plenty of idle vehicles: own-cluster matches dominate
    (2800, 6000, 8000, None),          # depth 3, 256-thread CTAs (84 KB of shared memory per replica)
])
def test_fused_search_rollout_vs_oracle(cuda_device, service, V, n_orders, windows):
    """speculative in-order neighbour search (csrc/rollout.cuh, SEARCH variant) vs the oracle's recursive DFS"""
    rng = np.random.default_rng(service + V)
    city = _city(service=service, ncs=True)
    minute, pick, drop = random_orders(city, n_orders, rng)
    loc0 = rng.choice(city.valid_nodes(), (2, V)).astype(np.int32)
    e = _engine(city, V, minute, pick, drop, R=2, trace=True)
    assert e.fused
    if windows:
        windows = list(windows)
        windows[-1] = e.T - sum(windows)
    rollout_vs_oracle(e, [make_oracle(city, V, minute, pick, drop) for _ in range(2)], loc0, windows)


def test_fused_search_contention(cuda_device):
    """all pickups in one cluster with NO idle vehicle: every order searches the same neighbour list, so almost
    every speculative pick inside a chunk collides with the previous order's (the re-evaluation path)."""
    rng = np.random.default_rng(4)
    city = _city(service=1600, ncs=True)
    V = 500
    sizes = [len(n) for n in city.cluster_nodes]
    hot = int(np.argmax(sizes))
    minute, _, drop = random_orders(city, 3000, rng)
    pick = rng.choice(city.cluster_nodes[hot], len(minute)).astype(np.int32)
    others = np.concatenate([city.cluster_nodes[int(c)] for c in city.neighbors(hot)])
    loc0 = rng.choice(others, V).astype(np.int32)
    e = _engine(city, V, minute, pick, drop, trace=True)
    rollout_vs_oracle(e, [make_oracle(city, V, minute, pick, drop)], loc0)


def test_fused_search_threshold(cuda_device):
    rng = np.random.default_rng(6)
    city = _city(service=2800, ncs=True)
    V = 400
    minute, pick, drop = random_orders(city, 4000, rng)
    loc0 = rng.choice(city.valid_nodes(), V).astype(np.int32)
    for thr in (10, 4):
        e = _engine(city, V, minute, pick, drop, reject_threshold=thr, trace=True)
        rollout_vs_oracle(e, [make_oracle(city, V, minute, pick, drop, threshold=thr)], loc0)
        e.close()


def test_fused_search_equals_per_phase_state(cuda_device):
    rng = np.random.default_rng(22)
    city = _city(service=2800, ncs=True)
    V, R = 777, 3
    minute, pick, drop = random_orders(city, 6000, rng)
    loc0 = rng.choice(city.valid_nodes(), (R, V)).astype(np.int32)
    a = _engine(city, V, minute, pick, drop, R=R)
    b = _engine(city, V, minute, pick, drop, R=R)
    a.reset(loc0); b.reset(loc0)
    for k0, n in ((0, 5), (5, 1), (6, a.T - 6)):
        a.rollout(k0, n)
        for k in range(k0, k0 + n):
            b.update(k); b.match(k); b.supply_expect(k)
        for name in ("veh_loc", "veh_cluster", "veh_arrive", "veh_dest", "veh_key", "order_res",
                     "per_match", "per_dispatch", "idle_live", "supply", "n_orders"):
            assert bool((a.tensors[name] == b.tensors[name]).all()), f"{name} differs after tick {k0 + n - 1}"
    sa, sb = a.stats(), b.stats()
    keep = [0, 1, 2, 3, 4, 5, 7, 8, 9]
    assert bool((sa[:, keep] == sb[:, keep]).all())


@pytest.mark.parametrize("side,ncs,service", [(400, False, 400), (400, True, 1000), (1500, False, 1500)])
def test_fused_rollout_other_grids(cuda_device, side, ncs, service):
    """768-cell grid (C > 256, several classification rounds per thread, empty cells) and a coarse 24-cell grid
    (C not a multiple of 4, idle lists of hundreds of vehicles)"""
    from vehicles_dispatch_simulator_b200.synthetic import synthetic_grid_city
    rng = np.random.default_rng(side)
    city = synthetic_grid_city(side_m=side, service_m=service, neighbor_can_server=ncs, n_nodes=900)
    V = 1200
    minute, pick, drop = random_orders(city, 6000, rng)
    loc0 = rng.choice(city.valid_nodes(), (2, V)).astype(np.int32)
    e = _engine(city, V, minute, pick, drop, R=2, trace=True)
    rollout_vs_oracle(e, [make_oracle(city, V, minute, pick, drop) for _ in range(2)], loc0, [3, 1, e.T - 4])


def test_fused_rollout_single_hot_cluster(cuda_device):
    """every vehicle and every pickup in ONE cluster: idle list of V entries, long sequential chain, ties."""
    rng = np.random.default_rng(2)
    city = _city()
    V = 700
    nodes = city.cluster_nodes[int(np.argmax([len(n) for n in city.cluster_nodes]))]
    minute, _, drop = random_orders(city, 3000, rng)
    pick = rng.choice(nodes, len(minute)).astype(np.int32)
    loc0 = rng.choice(nodes, V).astype(np.int32)
    e = _engine(city, V, minute, pick, drop, trace=True)
    rollout_vs_oracle(e, [make_oracle(city, V, minute, pick, drop)], loc0)


def test_fused_rollout_threshold(cuda_device):
    """reject-on-timeout with the intended 10-minute window: a rejected order must not consume a vehicle."""
    rng = np.random.default_rng(3)
    city = _city()
    V = 300
    minute, pick, drop = random_orders(city, 3000, rng)
    loc0 = rng.choice(city.valid_nodes(), V).astype(np.int32)
    for thr in (10, 3, 0):
        e = _engine(city, V, minute, pick, drop, reject_threshold=thr, trace=True)
        rollout_vs_oracle(e, [make_oracle(city, V, minute, pick, drop, threshold=thr)], loc0)
        e.close()


def test_fused_equals_per_phase_state(cuda_device):
    """HBM vehicle table + results after the fused rollout == after the per-phase kernels, bit for bit."""
    rng = np.random.default_rng(21)
    city = _city()
    V, R = 777, 4                       # V not a multiple of 8: padding lanes
    minute, pick, drop = random_orders(city, 6000, rng)
    loc0 = rng.choice(city.valid_nodes(), (R, V)).astype(np.int32)
    a = _engine(city, V, minute, pick, drop, R=R)
    b = _engine(city, V, minute, pick, drop, R=R)
    a.reset(loc0); b.reset(loc0)
    for k0, n in ((0, 7), (7, 1), (8, a.T - 8)):
        a.rollout(k0, n)
        for k in range(k0, k0 + n):
            b.update(k); b.match(k); b.supply_expect(k)
        for name in ("veh_loc", "veh_cluster", "veh_arrive", "veh_dest", "veh_key", "order_res",
                     "per_match", "per_dispatch", "idle_live", "supply", "n_orders"):
            assert bool((a.tensors[name] == b.tensors[name]).all()), f"{name} differs after tick {k0 + n - 1}"
    assert bool((a.stats() == b.stats()).all())


@pytest.mark.parametrize("ncs", [False])
def test_fused_tick_with_dispatch_lockstep(cuda_device, ncs):
    """the RL loop: one fused tick, then the dispatch primitive, per time slot."""
    rng = np.random.default_rng(11)
    city = _city()
    V = 400
    minute, pick, drop = random_orders(city, 4000, rng)
    loc0 = rng.choice(city.valid_nodes(), (2, V)).astype(np.int32)
    e = _engine(city, V, minute, pick, drop, R=2)
    assert e.fused
    lockstep(e, [make_oracle(city, V, minute, pick, drop) for _ in range(2)], loc0, dispatch=True, fused=True)


@pytest.mark.parametrize("name", [n for n in SMALL_CASES if "dispatch" not in n])
def test_fused_rollout_golden(cuda_device, name):
    """fused rollout vs the trace recorded from the unmodified Python reference."""
    z = load_golden(name)
    city, V, p = golden_city(z)
    e = _engine(city, V, z["in_order_minute"], z["in_order_pickup"], z["in_order_delivery"], R=2, period=p, trace=True)
    assert e.fused
    e.reset(z["in_veh_loc0"]); e.rollout()
    T = int(z["tr_final"][6])
    tr = e.trace.cpu().numpy()
    for r in range(2):
        assert np.array_equal(tr[r, :, 0], z["tr_per_match"][:T])
        assert np.array_equal(tr[r, :, 1], z["tr_per_dispatch"][:T])
        assert np.array_equal(tr[r, :, 2], z["tr_supply"][:T])
        assert np.array_equal(tr[r, :, 3], z["tr_n_orders"][:T])
        veh, wait, _ = e.order_results(r)
        assert np.array_equal(veh, z["tr_order_vehicle"]) and np.array_equal(wait, z["tr_order_wait"])
    st = e.stats().cpu().numpy()
    assert tuple(st[0][:6]) == tuple(z["tr_final"][:6])


@pytest.mark.parametrize("case", ["kmeans", "grid6000", "grid5000d3", "grid400v10000"])
def test_fused_rollout_real_day(cuda_device, case):
    """Shipped 2016-11-01 day (Kmeans-192 / 2000 vehicles = BASELINE configs[0]; Grid / 6000; Grid / 5000 / depth 3;
    Grid SideLengthMeter=400 / 10000 vehicles = the real-city parity variant of configs[4]) if present."""
    z = load_golden(case, real=True)
    if z is None:
        pytest.skip("tests/golden/_real not present (generated by make_golden.py --real)")
    city, V, p = golden_city(z)
    e = _engine(city, V, z["in_order_minute"], z["in_order_pickup"], z["in_order_delivery"], R=2, period=p, trace=True)
    e.reset(z["in_veh_loc0"]); e.rollout()
    T = int(z["tr_final"][6])
    tr = e.trace.cpu().numpy()
    assert np.array_equal(tr[1, :, 0], z["tr_per_match"][:T]) and np.array_equal(tr[1, :, 2], z["tr_supply"][:T])
    veh, wait, _ = e.order_results(1)
    assert np.array_equal(veh, z["tr_order_vehicle"]) and np.array_equal(wait, z["tr_order_wait"])
    assert tuple(e.stats().cpu().numpy()[1][:6]) == tuple(z["tr_final"][:6])
