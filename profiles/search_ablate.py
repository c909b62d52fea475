"""Timing ablations of match_nodes_kernel at one tick (state rebuilt from a snapshot before every launch).
Needs a library built from commit c2 of the round-2 search work (the VDS_SN_DBG knob: bit0 no tie resolution, bit1 no
commit, bit2 no pops, bit3 skeleton only); the shipped kernel has no such parameter.  Results: profiles/r2_search_nodes_summary.md."""
import os, sys
sys.path.insert(0, '/root/repo')
import bench, torch, numpy as np
w = bench.WORKLOADS["config3"]
torch.cuda.set_device(0)
city, tables, eng, loc0 = bench.build_workload(w, 4096, 0, 0)
K = 80
eng.reset(loc0)
for k in range(K):
    eng.update(k); eng.match(k)
eng.update(K)
torch.cuda.synchronize()
snap = {n: eng.tensors[n].clone() for n in ("veh_loc","veh_cluster","veh_arrive","veh_dest","veh_key","idle_live","per_dispatch","stats")}
gc = eng.sn["node_count_exact"].clone()
for dbg in (0, 1, 2, 4, 8, 3, 7, 0):
    for n, t in snap.items(): eng.tensors[n].copy_(t)
    eng.sn["node_count_exact"].copy_(gc)
    os.environ["VDS_SN_DBG"] = str(dbg)
    ts = []
    for rep in range(3):
        for n, t in snap.items(): eng.tensors[n].copy_(t)
        eng.sn["node_count_exact"].copy_(gc)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.match(K); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("dbg", dbg, "match ms", ["%.3f" % t for t in ts])
