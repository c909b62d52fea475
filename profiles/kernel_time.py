"""Device time of one episode of a bench workload (CUDA events, mean of N after warm-up).
Usage: kernel_time.py [workload] [reps]   (VDS_LIB_PATH selects a kernel variant)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import numpy as np
    import torch
    wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    R = bench.WORKLOADS[wl]["replicas"]
    torch.cuda.set_device(0)
    city, tables, eng, loc0 = bench.build_workload(bench.WORKLOADS[wl], R, 0, 0)
    ts = []
    for i in range(reps + 3):
        eng.reset(loc0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.rollout(0, eng.T); b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    st = eng.stats().cpu().numpy()
    print(f"{os.environ.get('VDS_LIB_PATH', 'libvds.so')} {wl}: {np.mean(ts):.3f} ms/episode (min {np.min(ts):.3f}) "
          f"checksum {int(st[:, :4].sum())} {int(st[:, 6].sum())}")


if __name__ == "__main__":
    main()
