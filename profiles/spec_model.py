"""CPU model of the speculative chunk match (match_nodes_kernel, round 2): one config-3 replica is advanced by the C
oracle to a tick, the tick's orders are then matched on the per-node-queue model (a) strictly in order and (b) 32 at a
time with the validity rule of csrc/search_nodes.cuh, and the script reports what the rule costs: rounds per 32-order
chunk, candidate-walk lengths, how often the tie rule fires.  Development aid (test infrastructure: it imports oracle/).
Usage: python profiles/spec_model.py [tick ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle import synth_ref  # noqa: E402
from tests.helpers import make_oracle  # noqa: E402
from vehicles_dispatch_simulator_b200.engine import tick_offsets  # noqa: E402
from vehicles_dispatch_simulator_b200.synthetic import DemandTables, synthetic_grid_city  # noqa: E402

DEAD = 0xFFFFFFFF


def main():
    ticks = [int(x) for x in sys.argv[1:]] or [80]
    K = int(os.environ.get("WALK", 8))
    city = synthetic_grid_city(side_m=800, service_m=2800, neighbor_can_server=True)
    tables = DemandTables(city)
    V = 5000
    minute, pick, drop = synth_ref.replica_orders(tables, 1234, 0)
    loc0 = synth_ref.replica_placement(city.valid_nodes(), V, 1234, 0)
    off, T = tick_offsets(minute, 10)
    o = make_oracle(city, V, minute, pick, drop)
    o.reset(loc0)
    sn = city.search_node_tables()
    rank = sn["node_rank"]; own_l = sn["own_list"]; reg_l = sn["search_list"]
    soff, sidx = city.search_lists()
    n2c = city.node2cluster
    C = city.n_clusters
    rank2c = np.zeros(sn["ranks_padded"], np.int64)
    for n in range(city.n_nodes):
        if n2c[n] >= 0:
            rank2c[rank[n]] = n2c[n]
    print("nodes", city.n_nodes, "own_pitch", own_l.shape[1], "search_pitch", reg_l.shape[1], "orders", len(minute))
    for k in range(T):
        if k not in ticks:
            o.step()
            continue
        o.update()
        loc = o.veh_loc()
        queues = {}
        for c in range(C):
            for pos, v in enumerate(o.idle_list(c)):
                queues.setdefault(int(rank[loc[v]]), []).append((c, pos, int(v)))     # key = (cluster list order)
        live = np.zeros(C, np.int64)
        for q in queues.values():
            live[q[0][0]] += len(q)
        lo, hi = off[k], off[k + 1]
        orders = [(int(pick[i]), int(n2c[pick[i]])) for i in range(lo, hi)]

        def resolve(p, c, queues, live):
            """-> (winner rank or None, walk length in entries, tie-set size, own?)"""
            own = live[c] > 0
            if own:
                lst = own_l[p]
            else:
                if sum(live[s] for s in sidx[soff[c] + 1:soff[c + 1]]) == 0:
                    return None, 0, 0, False
                lst = reg_l[p]
            for i, e in enumerate(lst):
                if e == DEAD:
                    return None, i, 0, own
                if queues.get(int(e & 0xFFFF)):
                    hi16 = e >> 16
                    tied = [int(e & 0xFFFF)]
                    j = i + 1
                    while j < len(lst) and lst[j] != DEAD and (lst[j] >> 16) == hi16:
                        if queues.get(int(lst[j] & 0xFFFF)):
                            tied.append(int(lst[j] & 0xFFFF))
                        j += 1
                    w = min(tied, key=lambda n: queues[n][0][:2])
                    return w, j, len(tied), own
            return None, len(lst), 0, own

        # (a) sequential, with statistics
        import copy
        q1 = copy.deepcopy(queues); l1 = live.copy()
        seq = []; walks = []; ties = []; owns = 0; dead = 0
        for p, c in orders:
            w, walk, nt, own = resolve(p, c, q1, l1)
            if w is None:
                seq.append(-1); dead += walk == 0
                continue
            walks.append(walk); ties.append(nt); owns += own
            cl, _, v = q1[w].pop(0)
            l1[cl] -= 1
            seq.append(v)
        ref = o.match() if False else None
        walks = np.array(walks); ties = np.array(ties)
        print(f"tick {k}: orders {len(orders)} matched {len(walks)} (own-cluster {owns}) nothing-in-reach {dead}")
        print("  walk length (entries incl. tie run) percentiles 50/90/99/max:",
              [int(np.percentile(walks, q)) for q in (50, 90, 99, 100)], " <=%d: %.1f%%" % (K, 100 * (walks <= K).mean()))
        print("  tie-set size: 1: %.1f%%  2-4: %.1f%%  >4: %.1f%%" % (100 * (ties == 1).mean(), 100 * ((ties > 1) & (ties <= 4)).mean(), 100 * (ties > 4).mean()))

        # (b) speculative chunks
        q2 = copy.deepcopy(queues); l2 = live.copy()
        spec = []; rounds_hist = []; coop = 0
        for b in range(0, len(orders), 32):
            chunk = orders[b:b + 32]
            pend = list(range(len(chunk)))
            out = [None] * len(chunk)
            rounds = 0
            while pend:
                rounds += 1
                res = [resolve(chunk[j][0], chunk[j][1], q2, l2) for j in pend]
                # validity in order
                taken = {}
                ncommit = 0
                for idx, j in enumerate(pend):
                    w, walk, nt, own = res[idx]
                    if w is None:
                        ncommit += 1; continue
                    if walk > K and idx > 0:
                        break                                   # long walk: resolved cooperatively at the head of a round
                    kth = taken.get(w, 0)
                    if kth >= len(q2[w]):
                        break                                   # winner node runs empty
                    if kth > 0 and nt > 1:
                        break                                   # tie decided by a popped head
                    if kth > 0 and False:
                        break
                    taken[w] = kth + 1
                    ncommit += 1
                    if walk > K:
                        coop += 1
                for idx in range(ncommit):
                    j = pend[idx]; w = res[idx][0]
                    if w is None:
                        out[j] = -1
                    else:
                        cl, _, v = q2[w].pop(0); l2[cl] -= 1; out[j] = v
                pend = pend[ncommit:]
            rounds_hist.append(rounds)
            spec += out
        assert spec == seq, "speculative schedule diverged from the sequential result"
        rh = np.array(rounds_hist)
        print(f"  speculative: {len(rh)} chunks, rounds per chunk mean {rh.mean():.2f} (p50 {np.percentile(rh, 50):.0f}, p90 {np.percentile(rh, 90):.0f}, max {rh.max()}), "
              f"total rounds {rh.sum()} vs {len(orders) - dead} sequential order steps; cooperative walks {coop}")
        o.match(); o.supply_expect(); o.snapshot_pre_dispatch(); o.end_tick()


if __name__ == "__main__":
    main()
