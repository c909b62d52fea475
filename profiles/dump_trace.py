"""Developer tool: per-tick per-cluster (idle before match, idle after, supply, orders) of a bench workload
for a few replicas -> gpurun_out/trace_<workload>.npz"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import numpy as np
    import torch
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine
    from vehicles_dispatch_simulator_b200.synthetic import DemandTables, synthetic_grid_city
    wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    w = bench.WORKLOADS[wl]
    torch.cuda.set_device(0)
    city = synthetic_grid_city(side_m=w["side"], service_m=800, neighbor_can_server=False)
    tables = DemandTables(city)
    eng = DispatchEngine(city, w["vehicles"], replicas=R, ticks=tables.ticks, max_orders=tables.max_orders,
                         per_replica_orders=True, max_orders_per_tick=tables.max_orders_per_tick, device=0, trace=True)
    eng.generate_orders(tables, seed=bench.SEED, first_replica=0)
    loc0 = eng.generate_placement(seed=bench.SEED, first_replica=0)
    eng.reset(loc0)
    eng.rollout(0, eng.T)
    torch.cuda.synchronize()
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez_compressed(f"gpurun_out/trace_{wl}.npz", trace=eng.trace.cpu().numpy().astype(np.int16))
    print("ok", eng.stats()[0].tolist())


if __name__ == "__main__":
    main()
