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
