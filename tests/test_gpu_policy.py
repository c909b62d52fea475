"""-m gpu: the device-resident random DispatchFunction hook (BASELINE config 4) -- fused tick + policy kernel +
dispatch primitive per time slot -- vs the oracle driven by the host restatement of the same policy."""
import numpy as np
import pytest

from oracle.synth_ref import random_policy_moves
from tests.helpers import make_oracle, random_orders

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ncs,service,prob", [(False, 800, 0.3), (True, 1600, 0.15)])
def test_device_random_policy_lockstep(cuda_device, ncs, service, prob):
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine, tick_offsets
    from vehicles_dispatch_simulator_b200.synthetic import synthetic_grid_city
    rng = np.random.default_rng(17)
    city = synthetic_grid_city(side_m=800, service_m=service, neighbor_can_server=ncs, n_nodes=700)
    V, R, seed, first = 500, 3, 99, 40
    minute, pick, drop = random_orders(city, 5000, rng)
    off, T = tick_offsets(minute, 10)
    e = DispatchEngine(city, V, replicas=R, ticks=T, max_orders=len(minute), max_orders_per_tick=int(np.diff(off).max()))
    e.bind_shared_orders(minute, pick, drop)
    loc0 = rng.choice(city.valid_nodes(), (R, V)).astype(np.int32)
    oracles = [make_oracle(city, V, minute, pick, drop) for _ in range(R)]
    e.reset(loc0)
    for r, o in enumerate(oracles):
        o.reset(loc0[r])
    moved = 0
    for k in range(T):
        e.tick(k)
        p = e.policy_random_dispatch(k, seed=seed, first_replica=first, prob=prob)
        cnt = p["cnt"].cpu().numpy()
        for r, o in enumerate(oracles):
            o.update(); o.match(); o.supply_expect(); o.snapshot_pre_dispatch()
            idle = np.sort(np.concatenate([np.asarray(l, np.int64) for l in o.idle_lists()] + [np.zeros(0, np.int64)]))
            veh, node = random_policy_moves(city, idle, o.veh_cluster(), k, seed, first + r, prob)
            assert cnt[r] == len(veh), f"tick {k} replica {r}: {cnt[r]} moves on device, {len(veh)} on host"
            assert np.array_equal(p["veh"][r, :cnt[r]].cpu().numpy(), veh)
            assert np.array_equal(p["node"][r, :cnt[r]].cpu().numpy(), node)
            assert o.dispatch(veh, node) == len(veh)
            moved += len(veh)
            o.end_tick()
        later = e.tensors["idle_live"].cpu().numpy()
        for r, o in enumerate(oracles):
            assert np.array_equal(later[r], o.later_dispatch()), f"later_dispatch tick {k} replica {r}"
        if k % 4 == 0:
            for r in (0, R - 1):
                got, want = e.idle_lists(r), oracles[r].idle_lists()
                for c in range(city.n_clusters):
                    assert np.array_equal(got[c], want[c]), f"idle list order tick {k} replica {r} cluster {c}"
    assert moved > 100
    st = e.stats().cpu().numpy()
    for r, o in enumerate(oracles):
        veh, wait, _ = e.order_results(r)
        assert np.array_equal(veh, o.order_vehicle()) and np.array_equal(wait, o.order_wait())
        assert tuple(st[r][:6]) == tuple(o.stats()[:6]), f"{st[r]} vs {o.stats()}"
    # replicas differ (keyed by global id) and the split does not matter
    assert not np.array_equal(st[0][:6], st[1][:6])


@pytest.mark.parametrize("V,prob,windows", [(500, 0.3, None), (500, 0.05, [1, 2, 5, 0]), (2600, 0.2, None), (9000, 0.1, [3, 0])])
def test_fused_policy_rollout_equals_per_tick(cuda_device, V, prob, windows):
    """vds_rollout_policy_random (the hook fused into the replica-resident kernel: ONE launch per window) leaves
    bit-identical state, results and counters to the per-tick path tick -> policy kernel -> dispatch primitive, which
    the test above pins to the oracle.  128-, 256- and 512-thread CTA variants; window boundaries."""
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine, tick_offsets
    from vehicles_dispatch_simulator_b200.synthetic import synthetic_grid_city
    rng = np.random.default_rng(V)
    city = synthetic_grid_city(side_m=800, service_m=800, neighbor_can_server=False, n_nodes=700)
    R, seed, first = 3, 7, 1000
    minute, pick, drop = random_orders(city, 6000, rng)
    off, T = tick_offsets(minute, 10)
    loc0 = rng.choice(city.valid_nodes(), (R, V)).astype(np.int32)

    def make():
        e = DispatchEngine(city, V, replicas=R, ticks=T, max_orders=len(minute), max_orders_per_tick=int(np.diff(off).max()))
        e.bind_shared_orders(minute, pick, drop)
        e.reset(loc0)
        return e

    a, b = make(), make()
    for k in range(T):
        a.tick(k)
        a.policy_random_dispatch(k, seed=seed, first_replica=first, prob=prob)
    if windows:
        windows = list(windows); windows[-1] = T - sum(windows)
        k = 0
        for n in windows:
            b.rollout_policy_random(k, n, seed=seed, first_replica=first, prob=prob); k += n
    else:
        b.rollout_policy_random(0, T, seed=seed, first_replica=first, prob=prob)
    sa, sb = a.stats().cpu().numpy(), b.stats().cpu().numpy()
    assert sa[:, 4].min() > 50                                          # the policy did move vehicles
    assert np.array_equal(sa, sb), f"{sa} vs {sb}"
    for name in ("veh_loc", "veh_cluster", "veh_arrive", "veh_dest", "veh_key", "per_match", "per_dispatch", "idle_live",
                 "supply", "n_orders", "disp_seq"):
        assert np.array_equal(a.tensors[name].cpu().numpy(), b.tensors[name].cpu().numpy()), name
    n = len(minute) - 1
    assert np.array_equal(a.tensors["order_res"][:, :n].cpu().numpy(), b.tensors["order_res"][:, :n].cpu().numpy())
    a.close(); b.close()


def test_policy_rollout_with_neighbour_search_runs_per_tick(cuda_device):
    """with NeighborCanServer the hook cannot be fused (orders of a replica are sequential across clusters):
    rollout_policy_random runs the per-tick launches itself and gives the same state as the explicit loop."""
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine, tick_offsets
    from vehicles_dispatch_simulator_b200.synthetic import synthetic_grid_city
    rng = np.random.default_rng(5)
    city = synthetic_grid_city(side_m=800, service_m=1600, neighbor_can_server=True, n_nodes=500)
    V, R = 300, 2
    minute, pick, drop = random_orders(city, 3000, rng)
    off, T = tick_offsets(minute, 10)
    loc0 = rng.choice(city.valid_nodes(), (R, V)).astype(np.int32)
    engines = []
    for _ in range(2):
        e = DispatchEngine(city, V, replicas=R, ticks=T, max_orders=len(minute), max_orders_per_tick=int(np.diff(off).max()))
        e.bind_shared_orders(minute, pick, drop)
        e.reset(loc0)
        engines.append(e)
    a, b = engines
    for k in range(T):
        a.tick(k)
        a.policy_random_dispatch(k, seed=3, first_replica=0, prob=0.2)
    b.rollout_policy_random(0, T, seed=3, first_replica=0, prob=0.2)
    assert np.array_equal(a.stats().cpu().numpy(), b.stats().cpu().numpy())
    for name in ("veh_loc", "veh_cluster", "veh_arrive", "veh_dest", "veh_key", "idle_live"):
        assert np.array_equal(a.tensors[name].cpu().numpy(), b.tensors[name].cpu().numpy()), name
    a.close(); b.close()
