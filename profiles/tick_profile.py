"""Per-window device time of the fused rollout (CUDA events around vds_rollout windows),
with the per-window work counters.  Usage: tick_profile.py [workload] [window]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
    win = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    R = bench.WORKLOADS[wl]["replicas"]
    torch.cuda.set_device(0)
    city, tables, eng, loc0 = bench.build_workload(bench.WORKLOADS[wl], R, 0, 0)
    T = eng.T
    for rep in range(2):
        eng.reset(loc0)
        evs, sts = [], [eng.stats().cpu().numpy().astype(float).sum(0)]
        for k0 in range(0, T, win):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.rollout(k0, min(win, T - k0)); b.record()
            evs.append((k0, a, b))
            if rep == 1:
                sts.append(eng.stats().cpu().numpy().astype(float).sum(0))
        torch.cuda.synchronize()
    tot = 0.0
    print("k0  ms   orders/tick matches/tick lookups/tick arrivals/tick (per replica)")
    for i, (k0, a, b) in enumerate(evs):
        ms = a.elapsed_time(b); tot += ms
        d = (sts[i + 1] - sts[i]) / R / min(win, T - k0)
        print(f"{k0:3d} {ms:7.3f} {d[0]:8.1f} {d[8]:8.1f} {d[6]:9.1f} {d[7]:8.1f}")
    print("total ms", tot)


if __name__ == "__main__":
    main()
