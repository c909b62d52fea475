"""Per-tick RL loop (tick -> device policy -> dispatch primitive, 444 launches per episode, BASELINE config 4) issued
eagerly from Python vs replayed as ONE captured CUDA graph (DispatchEngine.capture_graph).  CUDA-event ms per episode."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    torch.cuda.set_device(0)
    city, tables, eng, loc0 = bench.build_workload(bench.WORKLOADS["config4"], 1024, 0, 0)
    T = eng.T
    obs = eng.bind_observations(ring=2)

    def episode():
        eng.reset(loc0)
        for k in range(T):
            eng.tick(k)
            eng.policy_random_dispatch(k, seed=1234, first_replica=0, prob=0.05)
        eng.stats()

    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return min(ts), sum(ts) / len(ts)

    eager = timed(episode)
    s0 = eng._stats_out.clone()
    g = eng.capture_graph(episode)
    graph = timed(g.replay)
    assert torch.equal(s0, eng._stats_out)
    import time
    t0 = time.perf_counter(); episode(); t_host = time.perf_counter() - t0; torch.cuda.synchronize()
    print({"eager_ms": eager, "graph_ms": graph, "host_issue_ms_eager": 1e3 * t_host, "launches_per_episode": 3 * T + 2,
           "obs_ring": list(obs.shape)})


if __name__ == "__main__":
    main()
