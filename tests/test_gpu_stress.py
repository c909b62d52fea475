"""-m gpu: randomised configurations of the replica-resident rollout kernel against the oracle -- vehicle counts
around the packing boundaries (Vp multiple of 8, 2048-slot and CTA-width switches), skewed pickups (very long idle
lists in a few clusters, single-vehicle clusters everywhere else), timeout thresholds, random window boundaries,
distinct placements per replica.  VDS_STRESS_CASES widens the sweep (default 24 cases + the regressions, ~5 s).
Case 211 (coarse grid, timeout 3, > 32 orders in a cluster with <= 32 idle vehicles) once dead-locked the kernel: a
warp shuffle sat behind a per-lane condition at the 32-order chunk boundary."""
import os

import numpy as np
import pytest

from tests.helpers import make_oracle, random_orders, rollout_vs_oracle

pytestmark = pytest.mark.gpu

N_CASES = int(os.environ.get("VDS_STRESS_CASES", "24"))
CASES = sorted(set(range(N_CASES)) | {211})


@pytest.fixture(autouse=True, params=["node_queues", "slot_lists"])
def _rollout_kernel(monkeypatch, request):
    """both replica-resident kernels: rollout_nq_kernel (VDS_NQ=1) and rollout_local_kernel (default)"""
    if request.param == "node_queues":
        monkeypatch.setenv("VDS_NQ", "1")


@pytest.mark.timeout(120)
@pytest.mark.parametrize("case", CASES)
def test_random_rollout_configuration(cuda_device, case):
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine, tick_offsets
    from vehicles_dispatch_simulator_b200.synthetic import synthetic_grid_city
    rng = np.random.default_rng(1000 + case)
    side = int(rng.choice([400, 800, 800, 1500]))
    city = synthetic_grid_city(side_m=side, service_m=side, neighbor_can_server=False, n_nodes=int(rng.choice([300, 700, 1200])))
    V = int(rng.choice([7, 50, 333, 1999, 2047, 2049, 2900, 5600]))
    n_orders = int(rng.integers(1500, 9000))
    minute, pick, drop = random_orders(city, n_orders, rng, hot=float(rng.choice([0.3, 0.6, 0.9])))
    if case % 2:                       # pile the deliveries up in three clusters: idle lists of hundreds of vehicles
        hot_nodes = np.concatenate([city.cluster_nodes[c] for c in rng.choice(city.n_clusters, 3, replace=False)
                                    if len(city.cluster_nodes[c])] or [city.valid_nodes()[:1]])
        m = rng.random(len(drop)) < 0.7
        drop = np.where(m, rng.choice(hot_nodes, len(drop)), drop).astype(np.int32)
    thr = [600_000_000_000, 600_000_000_000, 10, 3][case % 4]
    R = 3
    off, T = tick_offsets(minute, 10)
    e = DispatchEngine(city, V, replicas=R, ticks=T, max_orders=len(minute),
                       max_orders_per_tick=max(64, int(np.diff(off).max())), reject_threshold=thr, trace=True)
    e.bind_shared_orders(minute, pick, drop)
    loc0 = rng.choice(city.valid_nodes(), (R, V)).astype(np.int32)
    cuts = sorted(set(int(x) for x in rng.integers(1, T, 3)))
    windows = [b - a for a, b in zip([0] + cuts, cuts + [T])]
    rollout_vs_oracle(e, [make_oracle(city, V, minute, pick, drop, threshold=thr) for _ in range(R)], loc0, windows)
    e.close()


N_SEARCH = int(os.environ.get("VDS_STRESS_SEARCH_CASES", "8"))


@pytest.mark.timeout(180)
@pytest.mark.parametrize("case", range(N_SEARCH))
def test_random_search_configuration(cuda_device, case):
    """match_search_kernel (neighbour search, per-phase path) against the oracle's recursive DFS: random depth, vehicle
    count, skew and threshold; results, counters (incl. lookups) and final idle lists."""
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine, tick_offsets
    from vehicles_dispatch_simulator_b200.synthetic import synthetic_grid_city
    rng = np.random.default_rng(5000 + case)
    side = int(rng.choice([400, 800, 800]))
    service = side * int(rng.choice([2, 3, 4]))                      # depth 1..3
    city = synthetic_grid_city(side_m=side, service_m=service, neighbor_can_server=True, n_nodes=int(rng.choice([300, 700])))
    V = int(rng.choice([9, 60, 400, 1500, 3100]))
    minute, pick, drop = random_orders(city, int(rng.integers(1500, 7000)), rng, hot=float(rng.choice([0.3, 0.6, 0.9])))
    thr = [600_000_000_000, 600_000_000_000, 10, 4][case % 4]
    R = 3
    off, T = tick_offsets(minute, 10)
    e = DispatchEngine(city, V, replicas=R, ticks=T, max_orders=len(minute),
                       max_orders_per_tick=max(64, int(np.diff(off).max())), reject_threshold=thr)
    e.bind_shared_orders(minute, pick, drop)
    loc0 = rng.choice(city.valid_nodes(), (R, V)).astype(np.int32)
    oracles = [make_oracle(city, V, minute, pick, drop, threshold=thr) for _ in range(R)]
    e.reset(loc0)
    e.rollout(0, T)                                                  # update / match_search / supply per tick
    st = e.stats().cpu().numpy()
    for r, o in enumerate(oracles):
        o.reset(loc0[r]); o.run()
        veh, wait, _ = e.order_results(r)
        assert np.array_equal(veh, o.order_vehicle()), f"matched vehicle ids replica {r}"
        assert np.array_equal(wait, o.order_wait())
        assert tuple(st[r][:9]) == tuple(o.stats()[:9]), f"{st[r]} vs {o.stats()}"
    e.close()


N_POLICY = int(os.environ.get("VDS_STRESS_POLICY_CASES", "6"))


@pytest.mark.timeout(180)
@pytest.mark.parametrize("case", range(N_POLICY))
def test_random_fused_policy_configuration(cuda_device, case):
    """vds_rollout_policy_random (hook fused into the rollout kernel) == tick + policy kernel + dispatch primitive,
    bit for bit, on random vehicle counts / probabilities / thresholds / window boundaries."""
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine, tick_offsets
    from vehicles_dispatch_simulator_b200.synthetic import synthetic_grid_city
    rng = np.random.default_rng(9000 + case)
    side = int(rng.choice([400, 800, 1500]))
    city = synthetic_grid_city(side_m=side, service_m=side, neighbor_can_server=False, n_nodes=int(rng.choice([300, 700])))
    V = int(rng.choice([5, 40, 333, 2040, 2900, 5700, 9100]))
    prob = float(rng.choice([0.02, 0.2, 0.7, 1.0]))
    thr = [600_000_000_000, 10, 600_000_000_000, 3][case % 4]
    minute, pick, drop = random_orders(city, int(rng.integers(1500, 7000)), rng, hot=float(rng.choice([0.3, 0.9])))
    off, T = tick_offsets(minute, 10)
    R, seed, first = 2, 31 + case, 100 * case
    loc0 = rng.choice(city.valid_nodes(), (R, V)).astype(np.int32)

    def make():
        e = DispatchEngine(city, V, replicas=R, ticks=T, max_orders=len(minute),
                           max_orders_per_tick=max(64, int(np.diff(off).max())), reject_threshold=thr)
        e.bind_shared_orders(minute, pick, drop)
        e.reset(loc0)
        return e

    a, b = make(), make()
    for k in range(T):
        a.tick(k)
        a.policy_random_dispatch(k, seed=seed, first_replica=first, prob=prob)
    cuts = sorted(set(int(x) for x in rng.integers(1, T, 2)))
    k = 0
    for n in [q - p for p, q in zip([0] + cuts, cuts + [T])]:
        b.rollout_policy_random(k, n, seed=seed, first_replica=first, prob=prob); k += n
    assert np.array_equal(a.stats().cpu().numpy(), b.stats().cpu().numpy())
    for name in ("veh_loc", "veh_cluster", "veh_arrive", "veh_dest", "veh_key", "per_match", "per_dispatch", "idle_live",
                 "supply", "n_orders", "disp_seq"):
        assert np.array_equal(a.tensors[name].cpu().numpy(), b.tensors[name].cpu().numpy()), name
    n = len(minute) - 1
    assert np.array_equal(a.tensors["order_res"][:, :n].cpu().numpy(), b.tensors["order_res"][:, :n].cpu().numpy())
    a.close(); b.close()
