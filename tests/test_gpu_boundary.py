"""-m gpu: round-2 boundary rows -- drop-in imports (Demo_simulation.py unmodified), post-episode object state,
FocusOnLocalRegion / Reload as a device-side stream swap + compaction, the cluster-graph builder kernel
(8f-3), deterministic dispatch with duplicate vehicles, the observation ring + time/weather features (8f-1)."""
import os
import random
import shutil
import subprocess
import sys
import time

import numpy as np
import pytest

from tests.golden.make_golden import SMALL
from tests.helpers import GOLDEN, golden_city, load_golden, make_oracle, random_orders

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL_DATA = os.path.join(GOLDEN, "small_city", "data")

# The reference's Demo_simulation.py:1-20, character for character (the script a user of the reference runs).
DEMO = '''from config.setting import *
from simulator.simulator import Simulation

if __name__ == "__main__":
    EXPSIM = Simulation(
                        ClusterMode = ClusterMode,
                        DemandPredictionMode = DemandPredictionMode,
                        DispatchMode = DispatchMode,
                        VehiclesNumber = VehiclesNumber,
                        TimePeriods = TIMESTEP,
                        LocalRegionBound = LocalRegionBound,
                        SideLengthMeter = SideLengthMeter,
                        VehiclesServiceMeter = VehiclesServiceMeter,
                        NeighborCanServer = NeighborCanServer,
                        FocusOnLocalRegion = FocusOnLocalRegion,
                        )

    EXPSIM.CreateAllInstantiate()
    EXPSIM.SimCity()
'''


def test_demo_simulation_runs_unmodified(cuda_device, tmp_path):
    """Demo_simulation.py of the reference, unchanged, with PYTHONPATH pointing at the alias packages and cwd holding
    ./data (the reference's cwd-relative layout): the drop-in claim of INTEGRATION.md path A."""
    (tmp_path / "Demo_simulation.py").write_text(DEMO)
    os.symlink(SMALL_DATA, tmp_path / "data")
    env = dict(os.environ, TZ="UTC",
               PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "vehicles_dispatch_simulator_b200", "compat"), ROOT]))
    r = subprocess.run([sys.executable, "Demo_simulation.py"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    assert "Experiment over" in out and "Number of Grids: 192" in out and "Number of Vehicles: 300" in out
    assert "Number of Orders: 7000" in out and "Totally Order value:" in out
    rej = int(out.split("Number of Reject: ")[1].split()[0])
    assert 0 < rej < 7000


def _make(name, cls=None, data_dir=SMALL_DATA, **extra):
    from vehicles_dispatch_simulator_b200 import Simulation
    from vehicles_dispatch_simulator_b200.setting import LocalRegionBound, TIMESTEP
    os.environ["TZ"] = "UTC"
    time.tzset()
    kw = dict(SMALL[name])
    sim = (cls or Simulation)(
        ClusterMode=kw["ClusterMode"], DemandPredictionMode="None", DispatchMode="Simulation",
        VehiclesNumber=kw["VehiclesNumber"], TimePeriods=TIMESTEP, LocalRegionBound=kw.get("LocalRegionBound", LocalRegionBound),
        SideLengthMeter=800, VehiclesServiceMeter=kw.get("VehiclesServiceMeter", 800),
        NeighborCanServer=kw.get("NeighborCanServer", False), FocusOnLocalRegion=kw.get("FocusOnLocalRegion", False),
        data_dir=data_dir, **extra)
    random.seed(0)
    sim.CreateAllInstantiate()
    return sim


def test_focus_on_local_region_vs_reference(cuda_device):
    """FocusOnLocalRegion=True end to end: the order stream AS LOADED goes to the device and vds_load_orders compacts it
    (keep = both end points inside a cluster), bit-exact against the unmodified reference run with the same flag."""
    z = load_golden("grid_d1_focus")
    sim = _make("grid_d1_focus", replicas=2)
    e = sim.engine
    n = len(z["in_order_minute"])
    assert int(e.n_orders_total[0]) == n == len(sim.Orders) < len(sim._raw_stream[0])
    pd_ = e.order_pd[0, :n].cpu().numpy()
    assert np.array_equal(pd_ & 0xFFFF, z["in_order_pickup"]) and np.array_equal(pd_ >> 16, z["in_order_delivery"])
    from vehicles_dispatch_simulator_b200.engine import tick_offsets
    assert np.array_equal(e.tick_off[0].cpu().numpy(), tick_offsets(z["in_order_minute"], 10)[0])
    sim.SimCity()
    assert (sim.OrderNum, sim.RejectNum, sim.TotallyWaitTime, sim.SumOrderValue) == tuple(int(x) for x in z["tr_final"][:4])
    for r in range(2):
        veh, wait, _ = e.order_results(r)
        assert np.array_equal(veh, z["tr_order_vehicle"]) and np.array_equal(wait, z["tr_order_wait"])


def test_hook_free_simcity_leaves_objects_populated(cuda_device):
    """after a hook-free SimCity the Order / Cluster / Vehicle objects hold the end-of-episode
    state like the reference's do (post-episode analysis reads sim.Orders, sim.Clusters)."""
    z = load_golden("grid_d0")
    sim = _make("grid_d0")
    sim.SimCity()
    infos = [o.ArriveInfo for o in sim.Orders]
    assert infos[-1] is None and sum(i == "Reject" for i in infos) == int(z["tr_final"][1])
    assert sum(1 for i in infos if i and i.startswith("ArriveTime:")) > 0
    waits = np.array([-1 if o.PickupWaitTime is None else o.PickupWaitTime for o in sim.Orders])
    assert np.array_equal(waits, z["tr_order_wait"])
    later = z["tr_later_dispatch"][-1]
    assert [len(c.IdleVehicles) for c in sim.Clusters] == later.tolist()
    en_route = sum(len(c.VehiclesArrivetime) for c in sim.Clusters)
    assert en_route + int(later.sum()) == len(sim.Vehicles)
    veh = z["tr_order_vehicle"]
    for c in sim.Clusters:
        for v in c.VehiclesArrivetime:
            assert v.DeliveryPoint is not None and len(v.Orders) == 1 and veh[v.Orders[0].ID] == sim.Vehicles.index(v)
    # refresh_objects=False skips the host round trip (counters and tensors only)
    sim2 = _make("grid_d0", refresh_objects=False)
    sim2.SimCity()
    assert sim2.RejectNum == sim.RejectNum and all(o.ArriveInfo is None for o in sim2.Orders)


def test_reload_swaps_the_stream_on_the_device(cuda_device, tmp_path):
    """Reload (simulator.py:130-212): another day's file under data/test/ replaces the stream without a new engine;
    the episode equals a fresh Simulation built directly on that day."""
    import pandas as pd
    d = tmp_path / "data"
    shutil.copytree(SMALL_DATA, d)
    day1 = pd.read_csv(d / "order_20161101.csv")
    rng = np.random.default_rng(3)
    mn, mx = day1["Start_time"].min() // 60, day1["Start_time"].max() // 60
    keep = (rng.random(len(day1)) < 0.8) | (day1["Start_time"] // 60 == mn) | (day1["Start_time"] // 60 == mx)
    day2 = day1[keep].sample(frac=1.0, random_state=5).copy()            # fewer orders, shuffled file order, other hot spots
    day2["Start_time"] = day2["Start_time"] + 86400
    swap = rng.random(len(day2)) < 0.3
    day2.loc[swap, ["NodeS", "NodeE"]] = day2.loc[swap, ["NodeE", "NodeS"]].values
    os.makedirs(d / "test")
    day2.to_csv(d / "test" / "order_20161102.csv", index=False)
    d2 = tmp_path / "fresh" / "data"
    shutil.copytree(SMALL_DATA, d2)
    day2.to_csv(d2 / "order_20161101.csv", index=False)

    sim = _make("grid_d2", data_dir=str(d), replicas=2)
    sim.SimCity()
    eng = sim.engine
    first = (sim.RejectNum, sim.TotallyWaitTime)
    random.seed(7)
    sim.Reload("1102")
    assert sim.engine is eng                                              # same engine, same HBM buffers
    assert len(sim.Orders) == len(day2)
    sim.SimCity()
    fresh = _make("grid_d2", data_dir=str(d2), replicas=1)
    random.seed(7)
    fresh.Reset()
    fresh.SimCity()
    assert (sim.OrderNum, sim.RejectNum, sim.TotallyWaitTime, sim.SumOrderValue) == \
        (fresh.OrderNum, fresh.RejectNum, fresh.TotallyWaitTime, fresh.SumOrderValue)
    assert (sim.RejectNum, sim.TotallyWaitTime) != first
    a, b = sim.engine.order_results(1), fresh.engine.order_results(0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_load_orders_compaction_matches_numpy(cuda_device):
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine, tick_offsets
    from vehicles_dispatch_simulator_b200.synthetic import synthetic_grid_city
    rng = np.random.default_rng(2)
    city = synthetic_grid_city(n_nodes=600)
    city.node2cluster = city.node2cluster.copy()
    city.node2cluster[rng.choice(600, 150, replace=False)] = -1           # a quarter of the nodes lie outside the region
    n = 50_000
    minute = np.sort(rng.integers(17, 1440, n)).astype(np.int32)
    pick, drop = rng.integers(0, 600, n).astype(np.int32), rng.integers(0, 600, n).astype(np.int32)
    keep = (city.node2cluster[pick] >= 0) & (city.node2cluster[drop] >= 0)
    km = minute[keep] - minute[keep][0]
    off, T = tick_offsets(km, 10)
    e = DispatchEngine(city, 64, replicas=1, ticks=T, max_orders=n, max_orders_per_tick=int(np.diff(off).max()) + 8)
    assert e.load_orders(minute, pick, drop, drop_uncovered=True) == int(keep.sum())
    pd_ = e.order_pd[0, :keep.sum()].cpu().numpy()
    assert np.array_equal(pd_ & 0xFFFF, pick[keep]) and np.array_equal(pd_ >> 16, drop[keep])
    assert np.array_equal(e.tick_off[0].cpu().numpy(), off)
    from vehicles_dispatch_simulator_b200 import VdsError
    with pytest.raises(VdsError):
        e.load_orders(minute, pick, drop)                                 # uncovered nodes without the filter: loud
    e.close()


def test_cluster_graph_builder_kernel(cuda_device, tmp_path):
    """CreateCluster's O(C^2 n^2) mean-road-cost loop (simulator.py:594-631) as one kernel; the Neighbor
    lists it produces equal the ones the unmodified reference computed on the same CSVs (golden kmeans_d1)."""
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine
    z = load_golden("kmeans_d1")
    city, V, p = golden_city(z)
    e = DispatchEngine(city, 8, replicas=1, ticks=1, max_orders=1, max_orders_per_tick=1)
    got = e.cluster_cost_sums().cpu().numpy()
    A = city.cost_u8.astype(np.int64)
    for i in range(0, city.n_clusters, 7):
        for j in range(city.n_clusters):
            ni, nj = city.cluster_nodes[i], city.cluster_nodes[j]
            assert got[i, j] == (A[np.ix_(nj, ni)].sum() if len(ni) and len(nj) else 0)
    e.close()
    d = tmp_path / "data"
    shutil.copytree(SMALL_DATA, d)
    os.remove(d / "(104.011, 104.125, 30.618, 30.703)192KmeansClusteringNeighbor.csv")     # recompute on the GPU
    sim = _make("kmeans_d1", data_dir=str(d))
    assert np.array_equal(sim.city.nb_off, z["in_nb_off"]) and np.array_equal(sim.city.nb_idx, z["in_nb_idx"])
    a = open(d / "(104.011, 104.125, 30.618, 30.703)192KmeansClusteringNeighbor.csv").read()
    b = open(os.path.join(SMALL_DATA, "(104.011, 104.125, 30.618, 30.703)192KmeansClusteringNeighbor.csv")).read()
    assert a == b                                                         # the cache file the reference itself wrote


def test_dispatch_duplicate_and_invalid_moves_are_deterministic(cuda_device):
    """two moves naming the same vehicle -> the lowest move index wins (what the oracle's / a
    Python agent's sequential loop does), en-route and out-of-range entries are skipped, only applied moves are
    numbered.  Repeated 20 times: bit-identical state."""
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine, tick_offsets
    from vehicles_dispatch_simulator_b200.synthetic import synthetic_grid_city
    rng = np.random.default_rng(11)
    city = synthetic_grid_city(n_nodes=500)
    V, R = 400, 3
    minute, pick, drop = random_orders(city, 4000, rng, n_minutes=100)
    off, T = tick_offsets(minute, 10)
    e = DispatchEngine(city, V, replicas=R, ticks=T, max_orders=len(minute), max_orders_per_tick=int(np.diff(off).max()))
    e.bind_shared_orders(minute, pick, drop)
    loc0 = rng.choice(city.valid_nodes(), (R, V)).astype(np.int32)
    valid = city.valid_nodes().astype(np.int32)
    digests = set()
    for rep in range(20):
        oracles = [make_oracle(city, V, minute, pick, drop) for _ in range(R)]
        e.reset(loc0)
        for r, o in enumerate(oracles):
            o.reset(loc0[r])
        mrng = np.random.default_rng(5)
        for k in range(T):
            e.tick(k)
            offs, mv, mn = [0], [], []
            for r, o in enumerate(oracles):
                o.update(); o.match(); o.supply_expect(); o.snapshot_pre_dispatch()
                m = 300
                veh = mrng.integers(-2, V + 2, m).astype(np.int32)         # duplicates, en-route vehicles, out of range
                veh[m // 2:] = veh[:m - m // 2][::-1]                        # every vehicle named (at least) twice
                node = mrng.choice(valid, m).astype(np.int32)
                ok = (veh >= 0) & (veh < V)
                applied = o.dispatch(veh, node)                             # sequential loop: first mention wins
                assert 0 < applied < ok.sum()
                mv.append(veh); mn.append(node); offs.append(offs[-1] + m)
                o.end_tick()
            e.dispatch(k, np.array(offs, np.int32), np.concatenate(mv), np.concatenate(mn))
            later = e.tensors["idle_live"].cpu().numpy()
            for r, o in enumerate(oracles):
                assert np.array_equal(later[r], o.later_dispatch()), f"later_dispatch tick {k} replica {r}"
            if rep == 0 and k % 3 == 0:
                for r in range(R):
                    got, want = e.idle_lists(r), oracles[r].idle_lists()
                    for c in range(city.n_clusters):
                        assert np.array_equal(got[c], want[c]), f"idle list order tick {k} replica {r} cluster {c}"
        st = e.stats().cpu().numpy()
        for r, o in enumerate(oracles):
            assert tuple(st[r][:6]) == tuple(o.stats()[:6])
            if rep == 0:
                veh_, wait_, _ = e.order_results(r)
                assert np.array_equal(veh_, o.order_vehicle()) and np.array_equal(wait_, o.order_wait())
        digests.add(tuple(e.tensors[n].cpu().numpy().tobytes() for n in ("veh_key", "veh_arrive", "veh_dest", "veh_loc")))
    assert len(digests) == 1
    e.close()


@pytest.mark.parametrize("ncs,service", [(False, 800), (True, 1600)])
def test_observation_ring(cuda_device, ncs, service):
    """the compact per-tick state record.  Fused windows write it from shared memory; the per-phase path
    packs it with vds_observe; both equal the oracle's per-tick values (and the i32 trace buffer)."""
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine, tick_offsets
    from vehicles_dispatch_simulator_b200.synthetic import synthetic_grid_city
    rng = np.random.default_rng(23)
    city = synthetic_grid_city(service_m=service, neighbor_can_server=ncs, n_nodes=500)
    V, R = 350, 2
    minute, pick, drop = random_orders(city, 4000, rng, n_minutes=150)
    off, T = tick_offsets(minute, 10)
    loc0 = rng.choice(city.valid_nodes(), (R, V)).astype(np.int32)
    exp = np.zeros((R, T, 4, city.n_clusters), np.int64)
    for r in range(R):
        o = make_oracle(city, V, minute, pick, drop)
        o.reset(loc0[r])
        for k in range(T):
            o.update(); o.match(); o.supply_expect(); o.snapshot_pre_dispatch()
            exp[r, k] = [o.per_match(), o.n_orders(), o.supply(), o.per_dispatch()]
            o.end_tick()
    mk = lambda: DispatchEngine(city, V, replicas=R, ticks=T, max_orders=len(minute), max_orders_per_tick=int(np.diff(off).max()))
    # (a) one fused window, ring = whole episode
    e = mk(); e.bind_shared_orders(minute, pick, drop)
    obs = e.bind_observations()
    e.reset(loc0); e.rollout(0, T)
    assert np.array_equal(obs.cpu().numpy().astype(np.int64), exp)
    # (b) ring of 4 slots, windows of 3 ticks: slot (k mod 4) holds tick k
    obs = e.bind_observations(ring=4)
    e.reset(loc0)
    for k0 in range(0, T - T % 3, 3):
        e.rollout(k0, 3)
        got = obs.cpu().numpy().astype(np.int64)
        for k in range(k0, k0 + 3):
            assert np.array_equal(got[:, k % 4], exp[:, k]), f"tick {k}"
    # (c) per-phase calls + vds_observe
    obs = e.bind_observations(ring=2)
    e.reset(loc0)
    for k in range(T):
        e.update(k); e.match(k); e.supply_expect(k); e.observe(k)
        if k % 5 == 0:
            assert np.array_equal(obs[:, k % 2].cpu().numpy().astype(np.int64), exp[:, k]), f"tick {k}"
    e.unbind_observations()
    e.close()


def test_time_and_weather_features_kernel(cuda_device):
    """GetTimeAndWeather (simulator.py:842-866) per tick, on the device, vs the facade's host method on the small
    city's 2016-11-01 day (and a day late in the month)."""
    import pandas as pd
    from vehicles_dispatch_simulator_b200.objects import Order
    sim = _make("grid_d0")
    e = sim.engine
    tables = (sim.WeatherType, sim.MinimumTemperature, sim.MaximumTemperature, sim.WindDirection, sim.WindPower)
    for shift_days in (0, 22):
        t0 = sim.Orders[0].ReleasTime - sim.TimePeriods + pd.Timedelta(days=shift_days)
        t0_min = int((t0 - pd.Timestamp("1970-01-01")) // pd.Timedelta(minutes=1))
        got = e.time_features(t0_min, tables).cpu().numpy()
        for k in range(e.T):
            t = t0 + k * pd.Timedelta(minutes=10)
            if t.month != 11:
                assert np.isnan(got[k]).all()
                continue
            want = sim.GetTimeAndWeather(Order(0, t, 0, 0, None, None, None, None))
            assert np.allclose(got[k], np.array(want, np.float32), rtol=0, atol=1e-6), (k, got[k], want)
