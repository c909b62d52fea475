"""-m gpu: BASELINE.json's full-size workloads (bench.py's configs), checked through properties that do not
depend on the size -- conservation laws, determinism, replica-split independence -- plus bit-exact parity of a
sample of replicas against the oracle (the oracle needs ~10 ms per replica-day, so a handful is affordable)."""
import hashlib

import numpy as np
import pytest

import bench
from tests.helpers import engine_replica_orders, make_oracle

pytestmark = pytest.mark.gpu


def _digest(t):
    return hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()


def _check_conservation(eng, st):
    """simulator.py:917, 944-969: every consumed order is matched or rejected; a vehicle is idle or en route."""
    n_tot = eng.n_orders_total.cpu().numpy().astype(np.int64)
    assert np.array_equal(st[:, 0], n_tot - 1)                         # OrderNum = N - 1 (Q8)
    assert np.array_equal(st[:, 0], st[:, 1] + st[:, 8])               # orders = rejects + matches
    assert (st[:, 7] <= st[:, 8]).all()                                # arrivals <= matches (no dispatches here)
    assert (st[:, 9] == eng.T).all()
    arrive = eng.tensors["veh_arrive"][:, :eng.V].cpu().numpy()
    idle = (arrive == 0xFFFF).sum(1)
    assert np.array_equal(idle, eng.tensors["idle_live"].cpu().numpy().sum(1))   # per-cluster counts add up
    assert np.array_equal(eng.V - idle, st[:, 8] - st[:, 7])           # en route = matched and not yet arrived
    # a matched order's result word names a vehicle < V and a wait <= 255; rejects are 0xFFFF
    for r in (0, eng.R - 1):
        veh, wait, delta = eng.order_results(r)
        m = veh >= 0
        assert int(m.sum()) == int(st[r, 8]) and (veh[m] < eng.V).all()
        assert int(wait[m].sum()) == int(st[r, 2])


def _sample_parity(city, tables, eng, loc0, replicas):
    for r in replicas:
        minute, pick, drop = engine_replica_orders(eng, r, tables.n_slots)
        o = make_oracle(city, eng.V, minute, pick, drop)
        o.reset(loc0[r].cpu().numpy().astype(np.int32))
        o.run()
        veh, wait, delta = eng.order_results(r)
        assert np.array_equal(veh, o.order_vehicle()), f"matched vehicle ids, replica {r}"
        assert np.array_equal(wait, o.order_wait()), f"wait, replica {r}"
        assert tuple(eng.stats()[r].tolist()[:9]) == tuple(o.stats()[:9]), f"stats, replica {r}"


def test_config2_full_size(cuda_device):
    """1024 replicas x 192-grid x 2000 vehicles, depth 0 (BASELINE configs[1]; the bench's default workload)."""
    w = bench.WORKLOADS["config2"]
    city, tables, eng, loc0 = bench.build_workload(w, w["replicas"], 0, 0)
    eng.reset(loc0); eng.rollout(0, eng.T)
    st = eng.stats().cpu().numpy()
    d1 = (_digest(eng.tensors["order_res"]), _digest(eng.tensors["veh_loc"]), _digest(eng.tensors["veh_key"]))
    _check_conservation(eng, st)
    _sample_parity(city, tables, eng, loc0, (0, 511, 1023))
    # determinism: the same episode again, bit for bit (atomics only order unordered slots; keys decide)
    eng.reset(loc0); eng.rollout(0, eng.T)
    assert np.array_equal(st, eng.stats().cpu().numpy())
    assert d1 == (_digest(eng.tensors["order_res"]), _digest(eng.tensors["veh_loc"]), _digest(eng.tensors["veh_key"]))
    # windows: 148 single fused ticks == one fused episode
    eng.reset(loc0)
    for k in range(eng.T):
        eng.tick(k)
    assert np.array_equal(st, eng.stats().cpu().numpy())
    assert d1[0] == _digest(eng.tensors["order_res"])
    eng.close()


def test_config3_sample(cuda_device):
    """192-grid x 5000 vehicles, neighbour-search depth 3 (BASELINE configs[2]) at 64 replicas: conservation,
    determinism, and two replicas bit-exact against the oracle's recursive DFS."""
    w = bench.WORKLOADS["config3"]
    city, tables, eng, loc0 = bench.build_workload(w, 64, 0, 0)
    eng.reset(loc0); eng.rollout(0, eng.T)
    st = eng.stats().cpu().numpy()
    d1 = _digest(eng.tensors["order_res"])
    _check_conservation(eng, st)
    _sample_parity(city, tables, eng, loc0, (0, 63))
    eng.reset(loc0); eng.rollout(0, eng.T)
    assert np.array_equal(st, eng.stats().cpu().numpy()) and d1 == _digest(eng.tensors["order_res"])
    # replica g does not depend on the offset that hosts it (8-GPU sharding)
    city2, tables2, eng2, loc2 = bench.build_workload(w, 8, 0, 56)
    eng2.reset(loc2); eng2.rollout(0, eng2.T)
    assert np.array_equal(st[56:64], eng2.stats().cpu().numpy())
    eng.close(); eng2.close()


def test_config5_sample(cuda_device):
    """768-grid x 10000 vehicles on the 16,556-node synthetic city (274 MB cost table, 1024-thread CTAs, generated
    per-replica streams; BASELINE configs[4]) at 8 replicas: conservation, determinism, two replicas bit-exact against
    the oracle, independence of the replica offset (8-GPU sharding)."""
    w = bench.WORKLOADS["config5"]
    city, tables, eng, loc0 = bench.build_workload(w, 8, 0, 0)
    assert city.n_clusters == 768 and city.n_nodes == 16556 and eng.V == 10000
    eng.reset(loc0); eng.rollout(0, eng.T)
    st = eng.stats().cpu().numpy()
    d1 = (_digest(eng.tensors["order_res"]), _digest(eng.tensors["veh_key"]))
    _check_conservation(eng, st)
    _sample_parity(city, tables, eng, loc0, (0, 7))
    eng.reset(loc0); eng.rollout(0, eng.T)
    assert np.array_equal(st, eng.stats().cpu().numpy())
    assert d1 == (_digest(eng.tensors["order_res"]), _digest(eng.tensors["veh_key"]))
    city2, tables2, eng2, loc2 = bench.build_workload(w, 2, 0, 6)
    eng2.reset(loc2); eng2.rollout(0, eng2.T)
    assert np.array_equal(st[6:8], eng2.stats().cpu().numpy())
    eng.close(); eng2.close()


def test_config4_sample(cuda_device):
    """BASELINE configs[3] at full per-GPU size (1024 replicas x 192-grid x 2000 vehicles, random-policy Dispatch hook
    fused into the rollout kernel): two replicas bit-exact against the oracle driven by the NumPy restatement of the
    policy; determinism; offset independence."""
    from oracle.synth_ref import random_policy_moves
    w = bench.WORKLOADS["config4"]
    city, tables, eng, loc0 = bench.build_workload(w, w["replicas"], 0, 0)
    prob = w["policy"]
    eng.reset(loc0)
    eng.rollout_policy_random(0, eng.T, seed=bench.SEED, first_replica=0, prob=prob)
    st = eng.stats().cpu().numpy()
    d1 = (_digest(eng.tensors["order_res"]), _digest(eng.tensors["veh_key"]))
    assert (st[:, 4] > 1000).all()                                         # the hook moved vehicles in every replica
    assert np.array_equal(st[:, 0], st[:, 1] + st[:, 8])
    for r in (0, 1023):
        minute, pick, drop = engine_replica_orders(eng, r, tables.n_slots)
        o = make_oracle(city, eng.V, minute, pick, drop)
        o.reset(loc0[r].cpu().numpy().astype(np.int32))
        for k in range(eng.T):
            o.update(); o.match(); o.supply_expect(); o.snapshot_pre_dispatch()
            idle = np.sort(np.concatenate([np.asarray(l, np.int64) for l in o.idle_lists()] + [np.zeros(0, np.int64)]))
            veh, node = random_policy_moves(city, idle, o.veh_cluster(), k, bench.SEED, r, prob)
            assert o.dispatch(veh, node) == len(veh)
            o.end_tick()
        veh, wait, _ = eng.order_results(r)
        assert np.array_equal(veh, o.order_vehicle()), f"matched vehicle ids, replica {r}"
        assert np.array_equal(wait, o.order_wait())
        assert tuple(st[r][:6]) == tuple(o.stats()[:6]), f"{st[r]} vs {o.stats()}"
        got, want = eng.idle_lists(r), o.idle_lists()
        for c in range(city.n_clusters):
            assert np.array_equal(got[c], want[c]), f"final idle list order, replica {r} cluster {c}"
    eng.reset(loc0)
    eng.rollout_policy_random(0, eng.T, seed=bench.SEED, first_replica=0, prob=prob)
    assert np.array_equal(st, eng.stats().cpu().numpy())
    assert d1 == (_digest(eng.tensors["order_res"]), _digest(eng.tensors["veh_key"]))
    city2, tables2, eng2, loc2 = bench.build_workload(w, 4, 0, 1020)
    eng2.reset(loc2)
    eng2.rollout_policy_random(0, eng2.T, seed=bench.SEED, first_replica=1020, prob=prob)
    assert np.array_equal(st[1020:1024], eng2.stats().cpu().numpy())
    eng.close(); eng2.close()
