"""Profiling driver: one fused generate+prepare pass over config2's streams (used under ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    torch.cuda.set_device(0)
    city, tables, eng, loc0 = bench.build_workload(bench.WORKLOADS["config2"], 1024, 0, 0)
    for _ in range(2):
        eng.generate_orders(tables, seed=bench.SEED, first_replica=0, check=False, fused=True)
    torch.cuda.synchronize()
    print("done", int(eng.n_orders_total[0]))


if __name__ == "__main__":
    main()
