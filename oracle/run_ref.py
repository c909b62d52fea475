"""Time ONE episode of the unmodified Python reference (oracle/_ref, see make_ref.py) on its own shipped day:
BASELINE configs[0] = Kmeans-192 clusters / 2000 vehicles / order_20161101.  Prints one JSON line:
{ticks, simcity_s, setup_s, final, golden_ok}.  `simcity_s` is the wall time of Simulation.SimCity() only -- the
reference's own timing convention; setup (CreateAllInstantiate, ~30 s) is reported beside it.
With --seed 0 the end-of-episode counters must equal the golden line of SURVEY 8c (RejectNum 141716 ...), which pins
the copy that travelled to the box.  TEST / MEASUREMENT INFRASTRUCTURE ONLY."""
import argparse
import io
import json
import os
import random
import sys
import time
import zipfile
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
GOLDEN_SEED0 = [209422, 141716, 102637, 1671284]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--tag", default="0")
    ap.add_argument("--cluster-mode", default="KmeansClustering")
    ap.add_argument("--vehicles", type=int, default=2000)
    a = ap.parse_args()
    os.environ["TZ"] = "UTC"
    time.tzset()
    work = os.path.join("/tmp", "vds_ref_run")
    data = os.path.join(work, "data")
    os.makedirs(data, exist_ok=True)
    src = os.path.join(REF, "data")
    for name in os.listdir(src):                       # cwd-relative ./data (preprocessing/readfiles.py:102-107)
        s = os.path.join(src, name)
        if name.endswith(".zip"):
            with zipfile.ZipFile(s) as z:
                for m in z.namelist():
                    t = os.path.join(data, m)
                    if not os.path.exists(t):
                        tmp = t + f".{os.getpid()}.tmp"
                        with z.open(m) as fi, open(tmp, "wb") as fo:
                            while True:
                                b = fi.read(1 << 22)
                                if not b:
                                    break
                                fo.write(b)
                        os.replace(tmp, t)
        else:
            t = os.path.join(data, name)
            if not os.path.exists(t):
                tmp = t + f".{os.getpid()}.tmp"
                with open(s, "rb") as fi, open(tmp, "wb") as fo:
                    fo.write(fi.read())
                os.replace(tmp, t)
    os.chdir(work)
    sys.path.insert(0, REF)
    buf = io.StringIO()
    with redirect_stdout(buf):
        import numpy
        import pandas
        from config import setting
        from simulator.simulator import Simulation
        random.seed(a.seed)
        t0 = time.perf_counter()
        sim = Simulation(ClusterMode=a.cluster_mode, DemandPredictionMode="None", DispatchMode="Simulation",
                         VehiclesNumber=a.vehicles, TimePeriods=setting.TIMESTEP,
                         LocalRegionBound=(104.011, 104.125, 30.618, 30.703), SideLengthMeter=800,
                         VehiclesServiceMeter=800, NeighborCanServer=False, FocusOnLocalRegion=False)
        sim.CreateAllInstantiate()
        t1 = time.perf_counter()
        sim.SimCity()
        t2 = time.perf_counter()
    value = int(sum(o.OrderValue for o in sim.Orders if o.ArriveInfo != "Reject"))
    final = [sim.OrderNum, sim.RejectNum, sim.TotallyWaitTime, value]
    golden_ok = None
    if a.seed == 0 and a.cluster_mode == "KmeansClustering" and a.vehicles == 2000:
        golden_ok = final == GOLDEN_SEED0
    print(json.dumps({"tag": a.tag, "ticks": int(sim.step), "simcity_s": t2 - t1, "setup_s": t1 - t0, "final": final,
                      "golden_ok": golden_ok,
                      "versions": {"python": sys.version.split()[0], "numpy": numpy.__version__, "pandas": pandas.__version__}}))


if __name__ == "__main__":
    main()
