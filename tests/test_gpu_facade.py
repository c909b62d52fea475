"""-m gpu: the reference-facing Simulation facade end to end on the committed
small-city CSV dataset, against traces recorded from the unmodified reference.
The agent below is written exactly as one would write it against the
reference's objects (list/dict mutation) -- the drop-in claim."""
import os
import random
import time

import numpy as np
import pytest

from oracle.ref_harness import dispatch_decision
from tests.golden.make_golden import SMALL
from tests.helpers import GOLDEN, REAL_CASES, check_engine_against_golden, golden_city, golden_idle_lists, load_golden

pytestmark = pytest.mark.gpu


def _make(cls, name, **extra):
    from vehicles_dispatch_simulator_b200.setting import LocalRegionBound, TIMESTEP
    os.environ["TZ"] = "UTC"
    time.tzset()
    kw = dict(SMALL[name])
    sim = cls(ClusterMode=kw["ClusterMode"], DemandPredictionMode="None", DispatchMode="Simulation",
              VehiclesNumber=kw["VehiclesNumber"], TimePeriods=TIMESTEP, LocalRegionBound=LocalRegionBound,
              SideLengthMeter=800, VehiclesServiceMeter=kw.get("VehiclesServiceMeter", 800),
              NeighborCanServer=kw.get("NeighborCanServer", False), FocusOnLocalRegion=False,
              data_dir=os.path.join(GOLDEN, "small_city", "data"), **extra)
    random.seed(0)
    sim.CreateAllInstantiate()
    return sim


@pytest.mark.parametrize("name", ["grid_d0", "grid_d2", "kmeans_d1"])
def test_hook_free_simcity(cuda_device, name):
    from vehicles_dispatch_simulator_b200.simulation import Simulation
    z = load_golden(name)
    sim = _make(Simulation, name, replicas=3)
    sim.SimCity()
    assert (sim.OrderNum, sim.RejectNum, sim.TotallyWaitTime, sim.SumOrderValue) == tuple(int(x) for x in z["tr_final"][:4])
    assert sim.step == 148
    for r in range(3):
        veh, wait, _ = sim.engine.order_results(r)
        assert np.array_equal(veh, z["tr_order_vehicle"]) and np.array_equal(wait, z["tr_order_wait"])
    # second episode after Reset with a new (seeded) placement still runs and differs
    random.seed(1)
    sim.Reset()
    sim.SimCity()
    assert sim.OrderNum == int(z["tr_final"][0])


def _agent_class():
    from vehicles_dispatch_simulator_b200.setting import MINUTES
    from vehicles_dispatch_simulator_b200.simulation import Simulation

    class Agent(Simulation):
        trace = None

        def RewardFunction(self):                    # observes the object graph like a reference agent would
            vid = self._vindex
            self.trace.append(dict(
                per_match=[c.PerMatchIdleVehicles for c in self.Clusters],
                n_orders=[len(c.Orders) for c in self.Clusters],
                idle_after_match=[[vid[id(v)] for v in c.IdleVehicles] for c in self.Clusters],
                counters=(self.OrderNum, self.RejectNum, self.TotallyWaitTime),
                arrive_keys=[[vid[id(v)] for v in c.VehiclesArrivetime] for c in self.Clusters]))

        def DispatchFunction(self):
            self.trace[-1]["supply"] = np.array(self.SupplyExpect).astype(np.int64).tolist()
            self.trace[-1]["per_dispatch"] = [c.PerDispatchIdleVehicles for c in self.Clusters]
            if not self.do_dispatch:
                return
            for c in self.Clusters:
                for v in list(c.IdleVehicles):
                    nb = [n for n in c.Neighbor if len(n.Nodes)]
                    k = dispatch_decision(self.step, self._vindex[id(v)], len(nb))
                    if k < 0:
                        continue
                    to = nb[k]
                    node = to.Nodes[(self.step + self._vindex[id(v)]) % len(to.Nodes)][0]
                    cost = self.RoadCost(v.LocationNode, node)
                    c.IdleVehicles.remove(v)
                    v.DeliveryPoint = node
                    to.VehiclesArrivetime[v] = self.RealExpTime + np.timedelta64(cost * MINUTES)
                    self.DispatchNum += 1
                    self.TotallyDispatchCost += cost
    return Agent


@pytest.mark.parametrize("name", ["grid_d0", "grid_d0_dispatch", "grid_d1_dispatch"])
def test_agent_hooks_and_dispatch(cuda_device, name):
    z = load_golden(name)
    Agent = _agent_class()
    sim = _make(Agent, name)
    sim.trace = []
    sim.do_dispatch = bool(z["in_params"][5])
    sim.SimCity()
    assert len(sim.trace) == 148
    for k, t in enumerate(sim.trace):
        assert t["per_match"] == z["tr_per_match"][k].tolist(), f"per_match tick {k}"
        assert t["n_orders"] == z["tr_n_orders"][k].tolist(), f"n_orders tick {k}"
        assert t["supply"] == z["tr_supply"][k].tolist(), f"supply tick {k}"
        assert t["per_dispatch"] == z["tr_per_dispatch"][k].tolist(), f"per_dispatch tick {k}"
        assert t["counters"] == tuple(int(x) for x in z["tr_counters"][k]), f"counters tick {k}"
    assert (sim.OrderNum, sim.RejectNum, sim.TotallyWaitTime, sim.SumOrderValue, sim.DispatchNum, sim.TotallyDispatchCost) \
        == tuple(int(x) for x in z["tr_final"][:6])
    veh, wait, _ = sim.engine.order_results(0)
    assert np.array_equal(veh, z["tr_order_vehicle"]) and np.array_equal(wait, z["tr_order_wait"])
    # Order objects carry the reference's ArriveInfo strings
    infos = [o.ArriveInfo for o in sim.Orders]
    assert infos[-1] is None                                   # last order never processed (Q8)
    assert sum(i == "Reject" for i in infos) == int(z["tr_final"][1])
    assert all(i is None or i == "Reject" or i == "Success" or i.startswith("ArriveTime:2016-11-0") for i in infos)
    assert sum(1 for i in infos if i and i.startswith("ArriveTime:")) > 0


@pytest.mark.parametrize("name", ["grid_d0", "grid_d2", "kmeans_d1", "grid_d0_dispatch", "grid_d1_dispatch"])
def test_engine_small_city_vs_reference_trace(cuda_device, name):
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine, tick_offsets
    z = load_golden(name)
    city, V, p = golden_city(z)
    off, T = tick_offsets(z["in_order_minute"], p)
    e = DispatchEngine(city, V, replicas=2, period=p, ticks=T, max_orders=len(z["in_order_minute"]),
                       max_orders_per_tick=int(np.diff(off).max()))
    e.bind_shared_orders(z["in_order_minute"], z["in_order_pickup"], z["in_order_delivery"])
    assert check_engine_against_golden(e, z, replicas=(0, 1)) == 148


@pytest.mark.parametrize("name", REAL_CASES)
def test_engine_real_day_vs_reference_trace(cuda_device, name):
    """BASELINE config 1 (Kmeans / 2000 vehicles) and the two other SURVEY 8c
    shapes on the shipped day: every tick, every cluster, every order."""
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine, tick_offsets
    z = load_golden(name, real=True)
    if z is None:
        pytest.skip("tests/golden/_real not present")
    city, V, p = golden_city(z)
    off, T = tick_offsets(z["in_order_minute"], p)
    e = DispatchEngine(city, V, replicas=2, period=p, ticks=T, max_orders=len(z["in_order_minute"]),
                       max_orders_per_tick=int(np.diff(off).max()))
    e.bind_shared_orders(z["in_order_minute"], z["in_order_pickup"], z["in_order_delivery"])
    assert check_engine_against_golden(e, z, replicas=(0, 1), lists_every=6) == 148
