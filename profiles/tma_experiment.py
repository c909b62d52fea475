"""TMA (cp.async.bulk) staging experiment: window-start load / window-end write-back of the replica's vehicle table in
rollout_local_kernel as bulk async copies (VDS_TMA=1) vs LDG->STS / LDS->STG loops (VDS_TMA=0).  Prints ms per episode
for (a) one 148-tick window, (b) 148 one-tick windows (the per-tick RL path, where the table moves every tick),
(c) config4 per-tick (tick + policy kernel + dispatch primitive).  Run once per setting; numbers are CUDA-event times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def timed(fn, reps=5):
    import torch
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def main():
    import torch
    torch.cuda.set_device(0)
    city, tables, eng, loc0 = bench.build_workload(bench.WORKLOADS["config2"], 1024, 0, 0)
    T = eng.T

    def episode():
        eng.reset(loc0); eng.rollout(0, T)

    def per_tick():
        eng.reset(loc0)
        for k in range(T):
            eng.tick(k)

    def per_tick_policy():
        eng.reset(loc0)
        for k in range(T):
            eng.tick(k)
            eng.policy_random_dispatch(k, seed=1234, first_replica=0, prob=0.05)

    out = {"VDS_TMA": os.environ.get("VDS_TMA", "default"), "kernel": eng.rollout_kernel_name}
    for name, fn in (("episode_window_ms", episode), ("148_one_tick_windows_ms", per_tick), ("config4_per_tick_ms", per_tick_policy)):
        out[name] = timed(fn)
    ref = eng.stats().cpu().numpy().tolist()[0]
    out["stats_r0"] = ref
    print(out)


if __name__ == "__main__":
    main()
