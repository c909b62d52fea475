"""Developer tool: per-phase clock64 breakdown of the fused rollout kernel (needs a -DROLL_PROF build,
selected with VDS_LIB_PATH).  Prints the share of warp-time per phase."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

NAMES = ["update(1+2)", "bar", "scan(3)", "scatter+prefill(4+5)", "bar", "classify(6a)", "bar", "lane pass",
         "warp path(6b)", "bar after match", "tick tail", "tick head -> wl_cnt=0", "bar", "arrive scan", "bar", "arrivals dense"]
NAMES[0] = "park coff (my_off wait)"


def main():
    import torch
    from vehicles_dispatch_simulator_b200 import _native
    wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
    R = bench.WORKLOADS[wl]["replicas"]
    torch.cuda.set_device(0)
    city, tables, eng, loc0 = bench.build_workload(bench.WORKLOADS[wl], R, 0, 0)
    L = _native.lib()
    buf = (C.c_ulonglong * 16)()
    eng.reset(loc0); eng.rollout(0, eng.T)
    L.vds_debug_prof(buf, 1)
    eng.reset(loc0); eng.rollout(0, eng.T)
    L.vds_debug_prof(buf, 1)
    tot = sum(buf[:16])
    for i in range(16):
        print(f"{NAMES[i]:24s} {100.0 * buf[i] / tot:6.2f} %")


if __name__ == "__main__":
    main()
