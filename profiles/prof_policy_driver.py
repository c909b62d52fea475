"""Profiling driver: one episode of config4 through the fused policy rollout (used under ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    w = bench.WORKLOADS["config4"]
    torch.cuda.set_device(0)
    city, tables, eng, loc0 = bench.build_workload(w, w["replicas"], 0, 0)
    eng.reset(loc0)
    eng.rollout_policy_random(0, eng.T, seed=bench.SEED, first_replica=0, prob=w["policy"])
    torch.cuda.synchronize()
    print("episode done", eng.stats()[0].tolist(), "launches", eng.launches)


if __name__ == "__main__":
    main()
