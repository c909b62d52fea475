"""Profiling driver: one episode of a bench workload, tick by tick (used under
ncu; numbers printed under a profiler are never bench values)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
    R = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[wl]["replicas"]
    nt = int(sys.argv[3]) if len(sys.argv) > 3 else None
    torch.cuda.set_device(0)
    city, tables, eng, loc0 = bench.build_workload(bench.WORKLOADS[wl], R, 0, 0)
    eng.reset(loc0)
    eng.rollout(0, nt or eng.T)
    torch.cuda.synchronize()
    print("episode done", eng.stats()[0].tolist(), "launches", eng.launches)


if __name__ == "__main__":
    main()
