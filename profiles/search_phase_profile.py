"""Per-tick device time of the neighbour-search workload split by phase (update / match), CUDA events.
Usage: search_phase_profile.py [replicas] [stop_tick]   (VDS_SEARCH_NODES=0 selects the slot-list kernels).
With stop_tick the episode stops after that tick's update (so `ncu -k regex:match --launch-skip N` can hit it)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import numpy as np
    import torch
    w = bench.WORKLOADS["config3"]
    R = int(sys.argv[1]) if len(sys.argv) > 1 else w["replicas"]
    stop = int(sys.argv[2]) if len(sys.argv) > 2 else None
    torch.cuda.set_device(0)
    city, tables, eng, loc0 = bench.build_workload(w, R, 0, 0)
    T = eng.T if stop is None else stop + 1
    print("search_nodes_active", eng.search_nodes_active)
    for rep in range(2 if stop is None else 1):
        eng.reset(loc0)
        ev = []
        for k in range(T):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record(); eng.update(k); b.record(); eng.match(k); c.record()
            ev.append((a, b, c))
        torch.cuda.synchronize()
    up = np.array([a.elapsed_time(b) for a, b, c in ev]); ma = np.array([b.elapsed_time(c) for a, b, c in ev])
    print(f"update total {up.sum():.2f} ms, match total {ma.sum():.2f} ms")
    for k in range(0, T, 8):
        print(f"tick {k:3d}: update {up[k]:.3f} ms  match {ma[k]:.3f} ms")
    st = eng.stats().cpu().numpy()
    print("checksum", int(st[:, :4].sum()), int(st[:, 6].sum()))


if __name__ == "__main__":
    main()
