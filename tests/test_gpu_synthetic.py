"""-m gpu: the on-device Philox order/placement generator against its NumPy
restatement (oracle/synth_ref.py), and the generated workload end to end
against the oracle (bit-exact), incl. independence from the replica split."""
import numpy as np
import pytest

from oracle import synth_ref
from tests.helpers import engine_replica_orders, lockstep, make_oracle

pytestmark = pytest.mark.gpu


def _setup(R, first, V=400, ncs=False, service=800, orders_per_day=30000, n_nodes=900):
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine
    from vehicles_dispatch_simulator_b200.synthetic import DemandTables, synthetic_grid_city
    city = synthetic_grid_city(service_m=service, neighbor_can_server=ncs, n_nodes=n_nodes)
    tables = DemandTables(city, orders_per_day=orders_per_day)
    e = DispatchEngine(city, V, replicas=R, ticks=tables.ticks, max_orders=tables.max_orders,
                       per_replica_orders=True, max_orders_per_tick=tables.max_orders_per_tick)
    e.generate_orders(tables, seed=77, first_replica=first)
    loc0 = e.generate_placement(seed=77, first_replica=first)
    return city, tables, e, loc0


def test_generator_matches_numpy_philox(cuda_device):
    city, tables, e, loc0 = _setup(R=5, first=3)
    for r in range(5):
        minute, pick, drop = synth_ref.replica_orders(tables, 77, 3 + r)
        m2, p2, d2 = engine_replica_orders(e, r, tables.n_slots)
        assert len(minute) == len(m2)
        assert np.array_equal(pick, p2) and np.array_equal(drop, d2)
        assert np.array_equal(minute[:-1], m2[:-1])           # last order: never consumed (Q8)
        exp = synth_ref.replica_placement(city.valid_nodes(), e.V, 77, 3 + r)
        assert np.array_equal(exp, loc0[r].cpu().numpy().astype(np.int32))
    # Didi-shaped: mean close to the requested day volume
    n = e.n_orders_total.cpu().numpy()
    assert abs(n.mean() - tables.mean_total) < 6 * np.sqrt(tables.mean_total)


@pytest.mark.parametrize("ncs,service", [(False, 800), (True, 2800)])
def test_generated_workload_bit_exact(cuda_device, ncs, service):
    R = 4
    city, tables, e, loc0 = _setup(R=R, first=0, ncs=ncs, service=service)
    oracles = []
    for r in range(R):
        minute, pick, drop = engine_replica_orders(e, r, tables.n_slots)
        oracles.append(make_oracle(city, e.V, minute, pick, drop))
    lockstep(e, oracles, loc0.cpu().numpy().astype(np.int32), check_lists_every=7)


def test_split_independence(cuda_device):
    """replica g's results do not depend on which rank/offset hosts it."""
    _, _, ea, la = _setup(R=6, first=0)
    ea.reset(la); ea.rollout()
    sa = ea.stats().cpu().numpy()
    _, _, eb, lb = _setup(R=3, first=3)
    eb.reset(lb); eb.rollout()
    sb = eb.stats().cpu().numpy()
    assert np.array_equal(sa[3:6], sb)
    na = ea.n_orders_total.cpu().numpy()
    for r in range(3):
        assert np.array_equal(ea.tensors["order_res"][3 + r, :na[3 + r]].cpu().numpy(),
                              eb.tensors["order_res"][r, :na[3 + r]].cpu().numpy())


def test_fused_generate_prepare_equals_two_kernels(cuda_device):
    """vds_generate_prepared_orders (generator + pricing + per-(tick, cluster) grouping in one kernel, LUT-accelerated CDF
    search) writes bit-identical streams and derived arrays to vds_generate_orders + vds_prepare_orders."""
    import torch
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine
    from vehicles_dispatch_simulator_b200.synthetic import DemandTables, synthetic_grid_city
    city = synthetic_grid_city(n_nodes=900)
    tables = DemandTables(city, orders_per_day=40_000)
    mk = lambda: DispatchEngine(city, 300, replicas=5, ticks=tables.ticks, max_orders=tables.max_orders, per_replica_orders=True,
                                max_orders_per_tick=tables.max_orders_per_tick)
    a, b = mk(), mk()
    a.generate_orders(tables, seed=77, first_replica=9)
    b.generate_orders(tables, seed=77, first_replica=9, fused=True)
    n = a.n_orders_total.cpu().numpy()
    assert np.array_equal(n, b.n_orders_total.cpu().numpy())
    assert torch.equal(a.tick_off, b.tick_off) and torch.equal(a.value_total, b.value_total)
    assert torch.equal(a.tick_value, b.tick_value) and torch.equal(a.cluster_off, b.cluster_off)
    for r in range(5):
        for name in ("order_pd", "order_value"):
            assert torch.equal(getattr(a, name)[r, :n[r]], getattr(b, name)[r, :n[r]]), (name, r)
        for name in ("sorted_pd", "sorted_idx"):
            assert torch.equal(getattr(a, name)[r, :n[r] - 1], getattr(b, name)[r, :n[r] - 1]), (name, r)
    loc0 = a.generate_placement(seed=77, first_replica=9)
    for e in (a, b):
        e.reset(loc0); e.rollout(0, e.T)
    assert torch.equal(a.stats(), b.stats()) and torch.equal(a.tensors["order_res"], b.tensors["order_res"])
    a.close(); b.close()
