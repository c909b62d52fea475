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
