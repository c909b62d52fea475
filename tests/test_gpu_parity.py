"""-m gpu parity tests: CUDA path (through the C ABI) vs the CPU oracle on the
same seeded inputs, bit-exact (integer work: matched vehicle ids, reject set,
waits, per-cluster idle counts, SupplyExpect, idle-list ORDER, counters)."""
import numpy as np
import pytest

from tests.helpers import lockstep, make_oracle, random_orders

pytestmark = pytest.mark.gpu


def _city(side=800, service=800, ncs=False, n_nodes=700):
    from vehicles_dispatch_simulator_b200.synthetic import synthetic_grid_city
    return synthetic_grid_city(side_m=side, service_m=service, neighbor_can_server=ncs, n_nodes=n_nodes)


def _engine(city, V, minute, pickup, delivery, R=1, **kw):
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine, tick_offsets
    off, T = tick_offsets(minute, 10)
    e = DispatchEngine(city, V, replicas=R, ticks=T, max_orders=len(minute),
                       max_orders_per_tick=max(64, int(np.diff(off).max())), **kw)
    e.bind_shared_orders(minute, pickup, delivery)
    return e


@pytest.mark.parametrize("V,n_orders", [(150, 3000), (600, 6000), (2500, 4000)])
def test_local_match_lockstep(cuda_device, V, n_orders):
    rng = np.random.default_rng(V)
    city = _city()
    minute, pick, drop = random_orders(city, n_orders, rng)
    loc0 = rng.choice(city.valid_nodes(), V).astype(np.int32)
    e = _engine(city, V, minute, pick, drop)
    lockstep(e, [make_oracle(city, V, minute, pick, drop)], loc0)


@pytest.mark.parametrize("service,V", [(1600, 200), (2800, 400), (2800, 1500)])
def test_neighbor_search_lockstep(cuda_device, service, V):
    rng = np.random.default_rng(service + V)
    city = _city(service=service, ncs=True)
    assert city.depth_limit == (1 if service == 1600 else 3)
    minute, pick, drop = random_orders(city, 5000, rng)
    loc0 = rng.choice(city.valid_nodes(), V).astype(np.int32)
    e = _engine(city, V, minute, pick, drop)
    lockstep(e, [make_oracle(city, V, minute, pick, drop)], loc0)


def test_replicas_distinct_placements(cuda_device):
    rng = np.random.default_rng(5)
    city = _city()
    R, V = 6, 300
    minute, pick, drop = random_orders(city, 4000, rng)
    loc0 = rng.choice(city.valid_nodes(), (R, V)).astype(np.int32)
    e = _engine(city, V, minute, pick, drop, R=R)
    lockstep(e, [make_oracle(city, V, minute, pick, drop) for _ in range(R)], loc0, check_lists_every=5)


@pytest.mark.parametrize("ncs", [False, True])
def test_dispatch_primitive_lockstep(cuda_device, ncs):
    rng = np.random.default_rng(11)
    city = _city(service=1600 if ncs else 800, ncs=ncs)
    V = 400
    minute, pick, drop = random_orders(city, 4000, rng)
    loc0 = rng.choice(city.valid_nodes(), (2, V)).astype(np.int32)
    e = _engine(city, V, minute, pick, drop, R=2)
    lockstep(e, [make_oracle(city, V, minute, pick, drop) for _ in range(2)], loc0, dispatch=True)


def test_intended_timeout_threshold(cuda_device):
    """reject-on-timeout is real when the threshold is the intended 10 minutes."""
    rng = np.random.default_rng(3)
    city = _city()
    V = 300
    minute, pick, drop = random_orders(city, 3000, rng)
    loc0 = rng.choice(city.valid_nodes(), V).astype(np.int32)
    e = _engine(city, V, minute, pick, drop, reject_threshold=10)
    o = make_oracle(city, V, minute, pick, drop, threshold=10)
    lockstep(e, [o], loc0)
    veh, wait, _ = e.order_results(0)
    assert wait.max() <= 10 and (veh == -1).sum() > 0


def test_rollout_equals_stepwise(cuda_device):
    rng = np.random.default_rng(8)
    city = _city()
    V = 500
    minute, pick, drop = random_orders(city, 5000, rng)
    loc0 = rng.choice(city.valid_nodes(), V).astype(np.int32)
    e = _engine(city, V, minute, pick, drop, R=3)
    o = make_oracle(city, V, minute, pick, drop)
    o.reset(loc0); o.run()
    e.reset(loc0); e.rollout()
    st = e.stats().cpu().numpy()
    for r in range(3):
        assert tuple(st[r][:9]) == tuple(o.stats()[:9])
        assert np.array_equal(e.order_results(r)[0], o.order_vehicle())
    # determinism: run again, compare bytes
    res1 = e.tensors["order_res"].clone()
    e.reset(loc0); e.rollout()
    assert bool((e.tensors["order_res"] == res1).all())
