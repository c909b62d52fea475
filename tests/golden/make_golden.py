"""Regenerates the golden fixtures by running the UNMODIFIED reference
(/root/reference, via oracle/ref_harness.py).  Run in the build container:

    python tests/golden/make_golden.py            # small city (committed)
    python tests/golden/make_golden.py --real     # + shipped-data traces (git-ignored, travels with gpurun)
    python tests/golden/make_golden.py --real --only grid400v10000   # (re)generate one real case, keep the others

Small city: a 420-node synthetic dataset written in the reference's own CSV
formats under tests/golden/small_city/data/ (committed, ~1.5 MB), run through
the reference in four configurations.  Real data: the three SURVEY 8c
configurations on the shipped 2016-11-01 day; the packed inputs + traces go to
tests/golden/_real/*.npz (17 MB cost table -> not committed) and their digests
to tests/golden/real_digest.json (committed).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H  # noqa: E402

BOUND = (104.011, 104.125, 30.618, 30.703)
BOUND_STR = "(104.011, 104.125, 30.618, 30.703)"

SMALL = {
    "grid_d0": dict(ClusterMode="Grid", VehiclesNumber=150),
    "grid_d2": dict(ClusterMode="Grid", VehiclesNumber=120, VehiclesServiceMeter=2000, NeighborCanServer=True),
    "kmeans_d1": dict(ClusterMode="KmeansClustering", VehiclesNumber=100, VehiclesServiceMeter=1200, NeighborCanServer=True),
    "grid_d0_dispatch": dict(ClusterMode="Grid", VehiclesNumber=150, dispatch=True),
    "grid_d1_dispatch": dict(ClusterMode="Grid", VehiclesNumber=150, VehiclesServiceMeter=1600, NeighborCanServer=True, dispatch=True),
    # FocusOnLocalRegion (simulator.py:356-370, 382-402): nodes and orders outside the reference's own sub-box
    # (config/setting.py:16) are dropped, the grid is laid over the sub-box (10 x 10 cells of 800 m)
    "grid_d1_focus": dict(ClusterMode="Grid", VehiclesNumber=120, VehiclesServiceMeter=1600, NeighborCanServer=True,
                          FocusOnLocalRegion=True, LocalRegionBound=(104.035, 104.105, 30.625, 30.695)),
}
REAL = {
    "kmeans": dict(ClusterMode="KmeansClustering", VehiclesNumber=2000),
    "grid6000": dict(ClusterMode="Grid", VehiclesNumber=6000),
    "grid5000d3": dict(ClusterMode="Grid", VehiclesNumber=5000, VehiclesServiceMeter=2800, NeighborCanServer=True),
    # SURVEY 8d "parity variant of the 768-grid" (BASELINE config 5's shape on the REAL city): 32 x 24 cells of
    # 400 m, 111 of them empty, V = 10000
    "grid400v10000": dict(ClusterMode="Grid", VehiclesNumber=10000, SideLengthMeter=400, VehiclesServiceMeter=400),
}


def write_small_city(d, n_nodes=420, n_orders=7000, n_drivers=300, seed=2016):
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(seed)
    # nodes strictly inside grid cells (the reference raises on a boundary hit)
    lon = np.round(rng.uniform(BOUND[0] + 1e-4, BOUND[1] - 1e-4, n_nodes), 7)
    lat = np.round(rng.uniform(BOUND[2] + 1e-4, BOUND[3] - 1e-4, n_nodes), 7)
    ids = rng.choice(np.arange(10 ** 9, 6 * 10 ** 9, 7), n_nodes, replace=False)
    with open(os.path.join(d, "Node.csv"), "w") as f:
        f.write("0,NodeID,WayID,Longitude,Latitude,RoadName,Gid,Distance\n")
        for i in range(n_nodes):
            f.write(f"{i},{ids[i]},{1000 + i},{lon[i]:.7f},{lat[i]:.7f},,{i},{rng.uniform(50, 300):.1f}\n")
    perm = rng.permutation(n_nodes)                      # NodeIDList order differs from Node.csv order
    with open(os.path.join(d, "NodeIDList.txt"), "w") as f:
        for i in perm:
            f.write(f"{ids[i]}\n")
    # asymmetric, fractional road costs in minutes, indexed like the shipped AccurateMap.csv
    plon, plat = lon[perm], lat[perm]
    km = np.hypot((plon[:, None] - plon[None, :]) * 96.0, (plat[:, None] - plat[None, :]) * 111.0)
    A = np.minimum(40.0, 5.6 * km) + rng.uniform(0, 0.999, km.shape)
    A = np.where(rng.random(km.shape) < 0.004, rng.uniform(41, 79.3, km.shape), A)
    A[np.arange(n_nodes), np.arange(n_nodes)] = 0.0
    A = np.round(A, 3)
    with open(os.path.join(d, "AccurateMap.csv"), "w") as f:
        for row in A:
            f.write(",".join(f"{x:g}" for x in row) + "\n")
    # one day of orders with rush hours and hot spots
    base = 1477958400                                     # 2016-11-01 00:00:00 UTC
    hour_w = np.array([2, 1, 1, 1, 1, 2, 4, 8, 10, 8, 7, 7, 8, 8, 7, 7, 8, 10, 10, 8, 6, 5, 4, 3], float)
    sec = np.sort(rng.choice(24, n_orders, p=hour_w / hour_w.sum()) * 3600 + rng.integers(0, 3600, n_orders))
    hot = rng.choice(n_nodes, 25, replace=False)
    pick = np.where(rng.random(n_orders) < 0.55, rng.choice(hot, n_orders), rng.integers(0, n_nodes, n_orders))
    drop = np.where(rng.random(n_orders) < 0.45, rng.choice(hot, n_orders), rng.integers(0, n_nodes, n_orders))
    shuffle = rng.permutation(n_orders)                   # file order is not time order
    with open(os.path.join(d, "order_20161101.csv"), "w") as f:
        f.write("ID,Start_time,End_time,PointS_Longitude,PointS_Latitude,PointE_Longitude,PointE_Latitude,NodeS,NodeE\n")
        for j in shuffle:
            f.write(f"{hashlib.md5(str(j).encode()).hexdigest()},{base + sec[j]},{base + sec[j] + 900},"
                    f"{lon[pick[j]]},{lat[pick[j]]},{lon[drop[j]]},{lat[drop[j]]},{ids[pick[j]]},{ids[drop[j]]}\n")
    with open(os.path.join(d, "Drivers1101.csv"), "w") as f:
        f.write("DiverID,Start_time,NodeS\n")
        for i in range(n_drivers):
            f.write(f"{i},2016/11/1 8:01,{ids[rng.integers(0, n_nodes)]}\n")
    # a 192-label "Kmeans" partition (coarse lon/lat quantisation, shuffled labels)
    gx = np.minimum(15, ((lon - BOUND[0]) / (BOUND[1] - BOUND[0]) * 16).astype(int))
    gy = np.minimum(11, ((lat - BOUND[2]) / (BOUND[3] - BOUND[2]) * 12).astype(int))
    relabel = rng.permutation(192)
    lab = relabel[(gy * 16 + gx + (gx % 3 == 0) * 16) % 192]
    with open(os.path.join(d, BOUND_STR + "192KmeansClusteringClusters.csv"), "w") as f:
        f.write("0\n")
        for v in lab:
            f.write(f"{v}\n")


def save(path, inp, tr):
    np.savez_compressed(path, **{"in_" + k: v for k, v in inp.items()}, **{"tr_" + k: v for k, v in tr.items()})


def digest(tr):
    h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    return {"final": [int(x) for x in tr["final"]],
            "order_vehicle_sha256": h(tr["order_vehicle"].astype(np.int32)),
            "order_wait_sha256": h(tr["order_wait"].astype(np.int32)),
            "per_match_sha256": h(tr["per_match"].astype(np.int32)),
            "per_dispatch_sha256": h(tr["per_dispatch"].astype(np.int32)),
            "supply_sha256": h(tr["supply"].astype(np.int32)),
            "ref_wall_s": float(tr["ref_wall_s"][0])}


def main():
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    small_dir = os.path.join(HERE, "small_city")
    if only is None:
        if "--small-only" not in sys.argv:
            write_small_city(os.path.join(small_dir, "data"))
            nb = os.path.join(small_dir, "data", BOUND_STR + "192KmeansClusteringNeighbor.csv")
            if os.path.exists(nb):
                os.remove(nb)                                    # let the reference compute it
        for name, kw in SMALL.items():
            if "--small-only" in sys.argv and name != sys.argv[sys.argv.index("--small-only") + 1]:
                continue
            inp, tr, _ = H.run_reference(small_dir, seed=0, **kw)
            save(os.path.join(HERE, f"small_{name}.npz"), inp, tr)
            print("small", name, tr["final"].tolist())
    if "--real" in sys.argv:
        wd = H.real_data_dir()
        os.makedirs(os.path.join(HERE, "_real"), exist_ok=True)
        dpath = os.path.join(HERE, "real_digest.json")
        dg = json.load(open(dpath)) if (only and os.path.exists(dpath)) else {}
        for name, kw in REAL.items():
            if only and name != only:
                continue
            cache = f"/tmp/vds_ref/real_{name}.npz"
            if os.path.exists(cache):
                z = np.load(cache)
                inp = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
                tr = {k[3:]: z[k] for k in z.files if k.startswith("tr_")}
            else:
                inp, tr, _ = H.run_reference(wd, seed=0, **kw)
            cost = inp.pop("cost_u8")                    # identical for all real configs: stored once
            np.savez_compressed(os.path.join(HERE, "_real", "real_cost.npz"), cost_u8=cost)
            save(os.path.join(HERE, "_real", f"real_{name}.npz"), inp, tr)
            dg[name] = digest(tr)
            print("real", name, dg[name]["final"])
        json.dump(dg, open(dpath, "w"), indent=1)


if __name__ == "__main__":
    main()
