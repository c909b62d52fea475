#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched dispatch hot path on B200.

One "step" = one full-day rollout (T=148 time slots: Update -> order advance ->
match/reject -> SupplyExpect -> snapshots) of ALL replicas on this GPU, i.e.
R*T env-steps.  Workload = BASELINE.json configs[1] by default:
1024 replicas x 192-grid x 2000 vehicles, synthetic Didi-rate orders
(~200k orders/day/replica, Philox streams generated on device).

  value     env-steps/s with inputs resident in HBM (device-timed, CUDA events): fixed per-replica streams,
            prepared once; observables of the window's last tick only (a hook-free replay)
  value_fresh_streams   every step also GENERATES + PREPARES a fresh stream per replica on the device
            (vds_generate_orders + vds_prepare_orders inside the timed region, no PCIe)
  value_traced          every tick's per-cluster observation (idle before/after match, demand, SupplyExpect)
            is materialised into the observation ring (what a learner reads)
  e2e       the same through the public engine API with HOST inputs: every step copies the vehicle placement
            from pinned host memory and reads the episode returns back
  roofline  dominant kernel: algorithmic bytes (SURVEY 8d formula on measured counters) / measured launch
            time / measured HBM peak
  cpu_baseline  the C oracle (port of the reference loop) on all host cores, bounded sample (rank 0, N=1 only)
  extra_workloads       2-3 timed steps each of BASELINE configs[2], [3], [4] (config3/4/5), N=1 only
  returns_sha256        digest of the all-gathered episode returns of global replicas [0, 1024): equal for
            every N (replica g does not depend on the GPU that hosts it)

--impl reference times the reference's CPU implementation of the path on all host cores: the C port in oracle/
for the line's `value` (same synthetic workload), plus -- when oracle/_ref (the unmodified Python reference +
its data, made by oracle/make_ref.py) travelled with the snapshot -- the Python reference itself on its own
shipped day (BASELINE configs[0]) single-process and all-core, under `reference_python`.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (replicas/GPU, side_m, service_m, neighbor_can_server, vehicles)
    "config2": dict(replicas=1024, side=800, service=800, ncs=False, vehicles=2000,
                    desc="1024 replicas x 192-grid x 2000 vehicles, synthetic Didi-rate orders, depth 0"),
    "config3": dict(replicas=4096, side=800, service=2800, ncs=True, vehicles=5000,
                    desc="4096 replicas x 192-grid x 5000 vehicles, neighbour-search depth 3"),
    "config4": dict(replicas=1024, side=800, service=800, ncs=False, vehicles=2000, policy=0.05,
                    desc="1024 replicas/GPU (8192 on 8 GPUs) x 192-grid x 2000 vehicles, device-resident random-policy "
                         "Dispatch hook every tick, NCCL all-gather of returns"),
    "config5": dict(replicas=2048, side=400, service=400, ncs=False, vehicles=10000,
                    desc="2048 replicas/GPU x 768-grid x 10000 vehicles, depth 0"),
}
SEED = 1234


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def measured_traffic(workload, kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json);
    None if no capture of this workload / kernel has been taken."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
        if t and t["kernel"].split("<")[0] == kernel.split("<")[0]:
            return {"bytes_per_launch": t["bytes_per_launch"], "source": t["capture"]}
    except Exception:
        pass
    return None


# SM clock + throttle reasons DURING the timed region, sampled by a separate PROCESS (NVML every 5 ms with
# wall-clock stamps; the parent keeps the samples that fall inside the timed window).  Round 1 polled from a
# thread of this process: it shared the GIL with the launch loop and cost ~0.25 ms per step at N >= 2.
_SAMPLER = r"""
import json, sys, time
idx, period = int(sys.argv[1]), float(sys.argv[2])
out = {"t": [], "sm": [], "reasons": [], "max": None, "source": "nvml"}
try:
    import pynvml as n
    n.nvmlInit()
    h = n.nvmlDeviceGetHandleByIndex(idx)
    out["max"] = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
    get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
except Exception as e:
    print(json.dumps({"error": repr(e)})); sys.stdout.flush(); sys.exit(0)
print("ready"); sys.stdout.flush()
import select
while True:
    if select.select([sys.stdin], [], [], 0)[0]:
        break
    try:
        out["t"].append(time.time()); out["sm"].append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
        out["reasons"].append(int(get(h)))
    except Exception:
        pass
    time.sleep(period)
print(json.dumps(out)); sys.stdout.flush()
"""


class ClockSampler:
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = index
        if vis:
            try:
                phys = int(vis.split(",")[index])
            except ValueError:
                phys = index
        self.proc = None
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER, str(phys), os.environ.get("BENCH_CLOCK_PERIOD_S", "0.005")],
                                         stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            first = self.proc.stdout.readline().strip()
            if first != "ready":
                self.proc.wait(timeout=2)
                self.proc = None
        except Exception:
            self.proc = None
        self.t0 = self.t1 = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock sampling unavailable"]}
        try:
            self.proc.stdin.write("stop\n"); self.proc.stdin.flush()
            d = json.loads(self.proc.stdout.readline())
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock sampling unavailable"]}
        idx = [i for i, t in enumerate(d["t"]) if self.t0 <= t <= self.t1]
        if not idx and d["t"]:                      # region shorter than one period: nearest sample
            mid = 0.5 * (self.t0 + self.t1)
            idx = [int(np.argmin(np.abs(np.array(d["t"]) - mid)))]
        sm = [d["sm"][i] for i in idx]
        bits = 0
        for i in idx:
            bits |= d["reasons"][i]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": d["max"], "samples": len(sm),
                "reasons": sorted(n for n, b in self.BITS.items() if bits & b), "source": "nvml (separate process)",
                "window_s": self.t1 - self.t0}


_CITY_CACHE = {}


def build_workload(w, replicas, device, first_replica):
    from vehicles_dispatch_simulator_b200.engine import DispatchEngine
    from vehicles_dispatch_simulator_b200.synthetic import DemandTables, synthetic_grid_city
    key = (w["side"], w["service"], w["ncs"])
    if key not in _CITY_CACHE:                      # the 768-grid city takes ~30 s of host NumPy
        city = synthetic_grid_city(side_m=w["side"], service_m=w["service"], neighbor_can_server=w["ncs"])
        _CITY_CACHE[key] = (city, DemandTables(city))
    city, tables = _CITY_CACHE[key]
    eng = DispatchEngine(city, w["vehicles"], replicas=replicas, ticks=tables.ticks, max_orders=tables.max_orders,
                         per_replica_orders=True, max_orders_per_tick=tables.max_orders_per_tick, device=device)
    eng.generate_orders(tables, seed=SEED, first_replica=first_replica)
    loc0 = eng.generate_placement(seed=SEED, first_replica=first_replica)
    return city, tables, eng, loc0


def cpu_port_run(city, eng, loc0, tables, sample_replicas, budget_s, threads, oracles=None):
    """Oracle (C port of the reference loop) on host threads over the first
    `sample_replicas` replicas' streams, repeated until ~budget_s of wall."""
    from concurrent.futures import ThreadPoolExecutor
    from tests.helpers import engine_replica_orders, make_oracle
    S = sample_replicas
    locs = loc0[:S].cpu().numpy().astype(np.int32)
    if oracles is None:
        oracles = []
        for r in range(S):
            minute, pick, drop = engine_replica_orders(eng, r, tables.n_slots)
            oracles.append(make_oracle(city, eng.V, minute, pick, drop))

    def one(r):
        oracles[r].reset(locs[r])
        return oracles[r].run()

    ticks = 0
    t0 = time.perf_counter()
    rounds = 0
    with ThreadPoolExecutor(max_workers=threads) as ex:
        while True:
            ticks += sum(ex.map(one, range(S)))
            rounds += 1
            if time.perf_counter() - t0 >= budget_s:
                break
    wall = time.perf_counter() - t0
    return ticks / wall, wall, rounds, oracles


_JSON_FD = None


def _stdout_only_for_the_json_line():
    """stdout must carry exactly ONE line (the JSON).  Libraries write to fd 1 behind Python's back -- NCCL prints
    "NCCL version ..." there when NCCL_DEBUG asks for it -- so fd 1 is pointed at stderr for the whole run and the
    JSON line goes to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


# ------------------------------------------------------------------ one workload on this rank
class Runner:
    """Episode loop of one workload on this rank.  An episode = reset + rollout (+ fused hook) + stats on the main
    stream -- captured ONCE in a CUDA graph and replayed -- followed by the all-gather of the return record on a SIDE
    stream (double-buffered), so the next episode's kernels are not queued behind the collective."""

    def __init__(self, args, wname, R, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        from vehicles_dispatch_simulator_b200.parallel import ReplicaShard
        self.torch, self.dist = torch, dist
        self.args, self.wname, self.w = args, wname, WORKLOADS[wname]
        self.R, self.rank, self.world = R, rank, world
        self.dev = torch.device("cuda", local_rank)
        self.shard = ReplicaShard(R, rank, world)
        self.city, self.tables, self.eng, self.loc0 = build_workload(self.w, R, local_rank, self.shard.first_replica)
        self.T = self.eng.T
        self.policy = self.w.get("policy")
        self.main = torch.cuda.current_stream(self.dev)
        self.side = torch.cuda.Stream(self.dev, priority=-1)
        self.rec = [torch.zeros((R, 6), dtype=torch.int64, device=self.dev) for _ in range(2)]
        self.ret = [torch.zeros((self.shard.total, 6), dtype=torch.int64, device=self.dev) for _ in range(2)]
        self.ev_rec = [torch.cuda.Event() for _ in range(2)]
        self.ev_done = [torch.cuda.Event() for _ in range(2)]
        self.i = 0
        self.graph = None
        self.kernels_per_episode = None

    # the kernels of one episode, on the current stream
    def run_ticks(self):
        eng, T, policy = self.eng, self.T, self.policy
        if policy and eng.fused and not self.args.policy_per_tick:
            eng.rollout_policy_random(0, T, seed=SEED, first_replica=self.shard.first_replica, prob=policy)
        elif policy:    # hook every tick: one fused tick launch + policy kernel + dispatch primitive per time slot
            for k in range(T):
                eng.tick(k)
                eng.policy_random_dispatch(k, seed=SEED, first_replica=self.shard.first_replica, prob=policy)
        else:
            eng.rollout(0, T)

    def body(self, loc):
        self.eng.reset(loc)
        self.run_ticks()
        self.eng.stats()

    def capture(self):
        """reset + rollout + stats as ONE CUDA graph (stream capture of the C-ABI launches)."""
        torch = self.torch
        l0 = self.eng.launches
        self.body(self.loc0)                         # eager once: counts the launches of one episode
        self.kernels_per_episode = self.eng.launches - l0
        if self.args.no_graph:
            return
        try:
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.body(self.loc0)
            self.graph = g
        except Exception as e:                       # pragma: no cover -- keep the eager launches
            sys.stderr.write(f"bench: CUDA graph capture failed ({e!r}); eager launches\n")
            self.graph = None
            torch.cuda.synchronize(self.dev)

    def episode(self, loc=None):
        """One step.  Returns the (device) all-gathered returns buffer this step will fill."""
        torch = self.torch
        b = self.i & 1
        self.i += 1
        self.main.wait_event(self.ev_done[b])        # rec[b] / ret[b] of step i-2 have been consumed
        if self.graph is not None and loc is None:
            self.graph.replay()
        else:
            self.body(self.loc0 if loc is None else loc)
        self.rec[b].copy_(self.eng._stats_out[:, :6])
        self.ev_rec[b].record(self.main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_rec[b])
            if self.world > 1:
                self.dist.all_gather_into_tensor(self.ret[b], self.rec[b])
            else:
                self.ret[b].copy_(self.rec[b])
            self.ev_done[b].record(self.side)
        return self.ret[b]

    def drain(self):
        self.main.wait_stream(self.side)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, fn, steps, clocks=None):
        """EXACTLY `steps` calls of fn between two events on the main stream (the side stream is joined before the
        closing event), barrier + synchronize on both sides, max over ranks."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if clocks:
            clocks.begin()
        t_host0 = time.perf_counter()
        e0.record(self.main)
        for _ in range(steps):
            fn()
        self.drain()
        e1.record(self.main)
        self.barrier()
        host_ms = 1e3 * (time.perf_counter() - t_host0)
        if clocks:
            clocks.end()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, host_ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item())

    def kernel_times(self, reps):
        """Device time of the dominant kernel(s) per episode, CUDA events on the launch stream."""
        torch, eng, T = self.torch, self.eng, self.T
        w, city, policy, args = self.w, self.city, self.policy, self.args
        if (policy and eng.fused and not args.policy_per_tick) or (not policy and eng.fused):
            evs = []
            for _ in range(reps):
                eng.reset(self.loc0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); self.run_ticks(); b.record()
                evs.append((a, b))
            torch.cuda.synchronize(self.dev)
            name = "rollout+policy" if policy else "rollout"
            kt = {name: float(np.mean([a.elapsed_time(b) for a, b in evs]))}           # ms per launch
            kind = "policy" if policy else ("search" if (w["ncs"] and city.depth_limit > 0) else "local")
            return kt, name, 1, "%s<%d,%s>" % (eng.rollout_kernel_name, eng.rollout_threads, kind)
        if policy:
            names = ("tick", "policy+dispatch")
            fns = lambda k: (lambda: eng.tick(k), lambda: eng.policy_random_dispatch(
                k, seed=SEED, first_replica=self.shard.first_replica, prob=policy))
            dom, kname = "tick", "%s<%d,local> (1-tick windows)" % (eng.rollout_kernel_name, eng.rollout_threads)
        else:
            names = ("update", "match", "supply")
            fns = lambda k: (lambda: eng.update(k), lambda: eng.match(k), lambda: eng.supply_expect(k))
            dom, kname = "match", ("match_nodes_kernel" if eng.search_nodes_active else "match_search_kernel")
        evs = {n: [] for n in names}
        eng.reset(self.loc0)
        for k in range(T):
            for n, fn in zip(names, fns(k)):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                evs[n].append((a, b))
        torch.cuda.synchronize(self.dev)
        kt = {n: float(sum(a.elapsed_time(b) for a, b in evs[n])) for n in names}      # ms per episode
        return kt, dom, T, kname

    def roofline(self, reps):
        eng, R, T, policy = self.eng, self.R, self.T, self.policy
        V, Cn = eng.V, eng.nC
        peak, peak_src = measured_peak_gbs()
        kt, dom, launches_per_episode, kname = self.kernel_times(reps)
        st = eng.stats().cpu().numpy().astype(np.float64)
        O, A, M, P = st[:, 0].sum(), st[:, 7].sum(), st[:, 8].sum(), st[:, 6].sum()
        # algorithmic bytes: B_step = 12V + 12(A+M) + 8O + 16C + P, split per phase
        bytes_k = {"update": 12.0 * V * R * T + 12 * A + 8.0 * Cn * R * T,
                   "match": 12 * M + 8 * O + P + 4.0 * Cn * R * T,
                   "supply": 4.0 * Cn * R * T}
        b_total = sum(bytes_k.values())
        D = st[:, 4].sum()
        if policy:
            bytes_k["policy+dispatch"] = 2.0 * V * R * T + 20.0 * D      # idle flags read + (move record + vehicle record) per move
            b_total = sum(bytes_k.values())
        if dom == "rollout+policy":
            b_dom = b_total                                                # one kernel does the tick AND the hook
        else:
            b_dom = (b_total - bytes_k.get("policy+dispatch", 0.0)) if dom in ("rollout", "tick") else bytes_k["match"]
        ach = b_dom / (kt[dom] * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": kname,
                "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
                "traffic": measured_traffic(self.wname, kname),
                "avg_launch_us": 1e3 * kt[dom] / launches_per_episode,
                "algorithmic_bytes_per_launch": b_dom / launches_per_episode,
                "kernel_ms_per_episode": kt, "dominant_kernel": dom,
                "whole_tick": {"bytes_per_env_step": b_total / (R * T),
                               "achieved": b_total / (sum(kt.values()) * 1e-3) / 1e9,
                               "frac": b_total / (sum(kt.values()) * 1e-3) / 1e9 / peak},
                "per_env_step": {"O_t": O / (R * T), "A_t": A / (R * T), "M_t": M / (R * T), "P_t": P / (R * T)}}

    def close(self):
        self.graph = None
        self.eng.close()


def reference_python(budget_note):
    """The UNMODIFIED Python reference on its own shipped day (BASELINE configs[0]: Kmeans-192 / 2000 vehicles),
    if oracle/_ref travelled: (i) one process, (ii) one process per host core.  SimCity wall time only."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    runner = os.path.join(ROOT, "oracle", "run_ref.py")
    if not (os.path.isdir(os.path.join(ref, "simulator")) and os.path.exists(runner)):
        return {"unavailable": "oracle/_ref not present (python oracle/make_ref.py in the build container)"}
    cores = os.cpu_count() or 1

    def launch(n, tag):
        ps = [subprocess.Popen([sys.executable, runner, "--seed", str(i), "--tag", f"{tag}{i}"], stdout=subprocess.PIPE,
                               stderr=subprocess.DEVNULL, text=True) for i in range(n)]
        outs = []
        for p in ps:
            o, _ = p.communicate(timeout=1500)
            try:
                outs.append(json.loads(o.strip().splitlines()[-1]))
            except Exception:
                outs.append(None)
        return outs

    t0 = time.perf_counter()
    one = launch(1, "s")[0]
    if not one:
        return {"unavailable": "oracle/run_ref.py failed"}
    res = {"config": "Kmeans-192 clusters / 2000 vehicles / shipped 2016-11-01 day (BASELINE configs[0]); SimCity wall only",
           "single_process": {"env_steps_per_s": one["ticks"] / one["simcity_s"], "simcity_s": one["simcity_s"],
                              "setup_s": one["setup_s"], "final": one["final"], "golden_ok": one["golden_ok"]}}
    if os.environ.get("BENCH_REF_ALLCORE", "1") != "0":
        allc = [o for o in launch(cores, "a") if o]
        if allc:
            res["all_cores"] = {"processes": len(allc), "cores": cores,
                                "env_steps_per_s": sum(o["ticks"] for o in allc) / max(o["simcity_s"] for o in allc),
                                "simcity_s_max": max(o["simcity_s"] for o in allc)}
    res["wall_s"] = time.perf_counter() - t0
    res["versions"] = one.get("versions")
    res["note"] = budget_note
    return res


def main():
    _stdout_only_for_the_json_line()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--replicas", type=int, default=None, help="replicas per GPU (default: the workload's)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_workloads (config3/4/5) section")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured episode graph")
    ap.add_argument("--no-ref-python", action="store_true", help="--impl reference: skip the Python reference legs")
    ap.add_argument("--policy-per-tick", action="store_true",
                    help="config4: run the Dispatch hook as separate per-tick launches instead of fused into the rollout kernel")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    w = WORKLOADS[args.workload]
    R = args.replicas or w["replicas"]
    cores = os.cpu_count() or 1

    # ------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        torch.cuda.set_device(0)
        S = min(R, 4 * cores)
        city, tables, eng, loc0 = build_workload(w, S, 0, 0)       # the GPU only GENERATES the input streams here
        vals = []
        oracles = None
        for i in range(args.warmup + args.steps):
            v, wall, rounds, oracles = cpu_port_run(city, eng, loc0, tables, S, 0.0, cores, oracles)
            if i >= args.warmup:
                vals.append((v, wall))
        tot_ticks = sum(v * wl for v, wl in vals)
        tot_wall = sum(wl for _, wl in vals)
        value = tot_ticks / tot_wall
        line = {"impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * tot_wall / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
                "config": {"workload": w["desc"], "replicas_sampled": S, "clusters": city.n_clusters,
                           "vehicles": w["vehicles"], "ticks": tables.ticks,
                           "note": "each step = one full-day episode of the sampled replicas on host threads"},
                "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                 "sample": f"{S} replicas x {tables.ticks} ticks per step, C port of the reference loop, {cores} threads"},
                "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "cluster_ticks_per_sec": value * city.n_clusters}
        if not args.no_ref_python:
            line["reference_python"] = reference_python(
                "the reference is single-threaded Python; it cannot run the 1024-replica synthetic workload in bounded "
                "time, so it is timed on its own shipped day (the value of this line stays the C port on the same "
                "synthetic workload as the GPU arm)")
        emit_json(line)
        return 0

    # ----------------------------------------------------------------- our arm
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    run = Runner(args, args.workload, R, rank, world, local_rank)
    city, tables, eng, loc0, shard, T = run.city, run.tables, run.eng, run.loc0, run.shard, run.T
    policy = run.policy

    # (the clock sampler process is started BEFORE the warm-up: spawning it right before the timed region would leave
    #  the GPU idle for ~0.4 s and the first timed steps would pay for the clock ramp)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    run.capture()
    for _ in range(args.warmup):
        run.episode()
    if world > 1:                      # NCCL sets channels up lazily: settle the collective before anything is timed
        for _ in range(8):
            run.episode()
    run.drain()
    # ---- timed region: EXACTLY K steps, device-timed on the launching stream
    ret_box = []
    elapsed_ms, host_ms = run.timed(lambda: ret_box.append(run.episode()), args.steps, clocks)
    clk = clocks.stop() if clocks else None
    ret = ret_box[-1].clone()
    env_steps = world * R * T * args.steps
    value = env_steps / (elapsed_ms * 1e-3)
    launches = run.kernels_per_episode * args.steps
    digest_n = min(1024, shard.total)
    returns_sha = hashlib.sha256(ret[:digest_n].cpu().numpy().tobytes()).hexdigest()

    # ---- e2e: the call a user of the reference API makes per episode -- Reset() + SimCity() -- with HOST
    #      buffers: the episode's initial vehicle placement (InitVehiclesIntoCluster, simulator.py:249-258)
    #      comes from pinned host memory every step and the episode returns are read back to the host.  The
    #      order streams are bound once, like the reference's CreateAllInstantiate (excluded from its SimCity
    #      timing too).  e2e_stream_upload additionally re-uploads and re-prepares every
    #      replica's order stream every step (PCIe-bound).
    h_loc = torch.empty(loc0.shape, dtype=loc0.dtype, pin_memory=True)
    h_ret = torch.empty((shard.total, 6), dtype=torch.int64, pin_memory=True)
    h_loc.copy_(loc0)
    d_loc = torch.empty_like(loc0)

    def episode_e2e():
        d_loc.copy_(h_loc, non_blocking=True)
        out = run.episode(d_loc)
        run.drain()
        h_ret.copy_(out, non_blocking=True)

    episode_e2e()
    ms, hms = run.timed(episode_e2e, args.steps)
    e2e_value = env_steps / (max(ms, hms) * 1e-3)
    h2d = loc0.numel() * 2
    d2h = h_ret.numel() * 8
    mine = slice(shard.first_replica, shard.first_replica + R)
    assert torch.equal(h_ret[mine], ret[mine].cpu())

    h_pd = torch.empty(eng.order_pd.shape, dtype=eng.order_pd.dtype, pin_memory=True)
    h_toff = torch.empty(eng.tick_off.shape, dtype=eng.tick_off.dtype, pin_memory=True)
    h_pd.copy_(eng.order_pd); h_toff.copy_(eng.tick_off)

    def episode_upload():
        eng.order_pd.copy_(h_pd, non_blocking=True)
        eng.tick_off.copy_(h_toff, non_blocking=True)
        eng._compute_values()
        episode_e2e()

    episode_upload()
    ms, hms = run.timed(episode_upload, max(2, args.steps // 4))
    up_value = world * R * T * max(2, args.steps // 4) / (max(ms, hms) * 1e-3)
    up_h2d = h2d + h_pd.numel() * 4 + h_toff.numel() * 4
    assert torch.equal(h_ret[mine], ret[mine].cpu())

    # ---- the full SURVEY 8d unit of work, device resident
    #  (a) fresh synthetic streams every episode: generator + prepare inside the timed region
    def episode_fresh():
        eng.generate_orders(tables, seed=SEED, first_replica=shard.first_replica, check=False, fused=True)
        run.episode(loc0)

    episode_fresh()
    ms, _ = run.timed(episode_fresh, args.steps)
    fresh_value = env_steps / (ms * 1e-3)
    assert torch.equal(run.ret[(run.i - 1) & 1][mine].cpu(), ret[mine].cpu())     # same seed -> same streams -> same returns
    #  (b) every tick's observation record materialised (the ring covers the episode)
    obs = eng.bind_observations()
    run.episode(loc0)
    ms, _ = run.timed(lambda: run.episode(loc0), args.steps)
    traced_value = env_steps / (ms * 1e-3)
    obs_bytes = obs.numel() * 2
    obs_sum = int(obs[:, :, 3].sum().item())                                      # touch it: the record is really there
    eng.unbind_observations()
    del obs

    roof = run.roofline(max(3, args.steps))

    # ---- CPU baseline + in-bench parity spot check (rank 0, N=1 only)
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not policy:
        S = min(R, 4 * cores)
        v, wall, rounds, oracles = cpu_port_run(city, eng, loc0, tables, S, args.cpu_seconds, cores)
        cpu = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
               "sample": f"first {S} replicas x {T} ticks x {rounds} rounds ({wall:.1f} s), C port of the reference loop, {cores} threads"}
        got = eng.stats().cpu().numpy()
        keep = [0, 1, 2, 3, 4, 5, 7, 8] if (w["ncs"] and city.depth_limit > 0) else list(range(9))   # search path: own lookup count
        parity = all(tuple(got[r][keep]) == tuple(np.asarray(oracles[r].stats())[keep]) for r in range(S)) and \
            all(np.array_equal(eng.order_results(r)[0], oracles[r].order_vehicle()) for r in range(min(S, 4)))
        del oracles

    Cn, V = eng.nC, eng.V
    mean_total, depth = tables.mean_total, city.depth_limit
    stream_gb = (eng.order_pd.numel() * 8) / 1e9
    used_graph = run.graph is not None
    run.close()
    del run, eng

    # ---- the other BASELINE configurations, a few timed steps each (N=1 only; parity for them is in tests/)
    extra = None
    if world == 1 and rank == 0 and not args.no_extra and args.workload == "config2":
        extra = {}
        for name in ("config4", "config3", "config5"):
            try:
                t_build = time.perf_counter()
                ew = WORKLOADS[name]
                rx = Runner(args, name, ew["replicas"], 0, 1, local_rank)
                rx.capture()
                ksteps = 2 if name == "config3" else 3
                rx.episode()
                ms, _ = rx.timed(lambda: rx.episode(), ksteps)
                rf = rx.roofline(2)
                extra[name] = {"workload": ew["desc"], "replicas": ew["replicas"], "steps": ksteps,
                               "value": ew["replicas"] * rx.T * ksteps / (ms * 1e-3), "unit": "env-steps/s",
                               "ms_per_step": ms / ksteps, "kernel": rf["kernel"],
                               "kernel_ms_per_episode": rf["kernel_ms_per_episode"],
                               "roofline_frac": rf["frac"], "roofline_achieved_gbs": rf["achieved"],
                               "whole_tick_frac": rf["whole_tick"]["frac"], "per_env_step": rf["per_env_step"],
                               "returns_sha256": hashlib.sha256(rx.ret[(rx.i - 1) & 1].cpu().numpy().tobytes()).hexdigest(),
                               "build_s": None}
                rx.close()
                del rx
                torch.cuda.empty_cache()
                extra[name]["build_s"] = time.perf_counter() - t_build
            except Exception as e:                   # an extra must never take the headline line down
                extra[name] = {"error": repr(e)}

    if rank == 0:
        line = {"metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32",
                "data": "synthetic",
                "config": {"workload": w["desc"], "replicas_per_gpu": R, "clusters": Cn, "vehicles": V,
                           "ticks": T, "orders_per_day_mean": mean_total, "search_depth": depth,
                           "step": "one full-day rollout of all replicas (R*T env-steps): reset + fused rollout + stats "
                                   "(one CUDA graph) + all-gather of returns on a side stream.  `value` replays FIXED "
                                   "per-replica streams prepared once (order generation and the per-(tick, cluster) "
                                   "grouping are outside the timed region) and materialises per-cluster observables for the "
                                   "last tick only; value_fresh_streams / value_traced time the fuller units of work",
                           "l2": "inputs larger than L2 (order streams + results %.2f GB per step); no explicit flush" % stream_gb,
                           "cuda_graph": used_graph},
                "cluster_ticks_per_sec": value * Cn,
                "host_ms_per_step": host_ms / args.steps,
                "launch_overhead_ms_per_step": elapsed_ms / args.steps - roof["avg_launch_us"] * 1e-3 * (1 if roof["dominant_kernel"].startswith("rollout") else T),
                "returns_sha256": {"replicas": digest_n, "sha256": returns_sha},
                "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "inputs": "per step: vehicle placement u16[R,V] from pinned host memory -> reset -> rollout -> "
                                  "all-gather -> returns int64[R,6] to host; order streams bound once (reference: "
                                  "CreateAllInstantiate, outside SimCity timing)"},
                "e2e_stream_upload": {"value": up_value, "unit": "env-steps/s", "h2d_bytes_per_step": up_h2d,
                                      "d2h_bytes_per_step": d2h,
                                      "inputs": "as e2e plus every replica's order stream + tick offsets re-uploaded and "
                                                "re-prepared every step"},
                "value_fresh_streams": {"value": fresh_value, "unit": "env-steps/s",
                                        "what": "per step: vds_generate_prepared_orders (Philox draw + pricing + per-(tick, cluster) grouping "
                                                "of a fresh stream per replica, one fused kernel) + reset + rollout + stats, all device-resident"},
                "value_traced": {"value": traced_value, "unit": "env-steps/s", "obs_bytes_per_step": obs_bytes,
                                 "what": "per step: reset + rollout writing every tick's [R,4,C] u16 observation record "
                                         "(idle before/after match, demand, SupplyExpect) + stats", "obs_checksum": obs_sum},
                "gpu_launches": launches, "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
                "parity_check_vs_oracle": parity, "extra_workloads": extra}
        emit_json(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
