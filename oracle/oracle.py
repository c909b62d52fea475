"""ctypes wrapper over oracle/libvds_oracle.so (the C restatement).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PARITY_THRESHOLD = 600_000_000_000      # raw int of np.timedelta64(10*MINUTES)


def build(force=False):
    so = os.path.join(_HERE, "libvds_oracle.so")
    src = os.path.join(_HERE, "vds_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", so, src])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        p32 = C.POINTER(C.c_int32)
        L.vdso_create.restype = C.c_void_p
        L.vdso_create.argtypes = [C.c_int] * 7 + [C.c_int64] + [C.c_void_p] * 7
        for name in ("vdso_destroy", "vdso_update", "vdso_match", "vdso_supply_expect",
                     "vdso_snapshot_pre_dispatch", "vdso_end_tick", "vdso_step"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_void_p]
        L.vdso_reset.restype = None
        L.vdso_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.vdso_done.restype = C.c_int
        L.vdso_done.argtypes = [C.c_void_p]
        L.vdso_run.restype = C.c_int
        L.vdso_run.argtypes = [C.c_void_p]
        L.vdso_step_index.restype = C.c_int
        L.vdso_step_index.argtypes = [C.c_void_p]
        L.vdso_dispatch.restype = C.c_int
        L.vdso_dispatch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        for name in ("per_match", "per_dispatch", "later_dispatch", "supply", "n_orders",
                     "order_vehicle", "order_wait", "order_arrive_tick", "veh_loc",
                     "veh_dest", "veh_cluster", "veh_order"):
            f = getattr(L, "vdso_" + name)
            f.restype = p32
            f.argtypes = [C.c_void_p]
        L.vdso_idle_len.restype = C.c_int
        L.vdso_idle_len.argtypes = [C.c_void_p, C.c_int]
        L.vdso_idle_list.restype = C.POINTER(C.c_int)
        L.vdso_idle_list.argtypes = [C.c_void_p, C.c_int]
        L.vdso_arrive_len.restype = C.c_int
        L.vdso_arrive_len.argtypes = [C.c_void_p, C.c_int]
        L.vdso_arrive_list.restype = None
        L.vdso_arrive_list.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.vdso_stats.restype = None
        L.vdso_stats.argtypes = [C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Oracle:
    """One replica of the reference's tick loop, in integers.

    cost_u8[end, start]; node2cluster int32[nodes]; neighbour CSR; orders sorted
    by minute (post-ReadOrder order); period p minutes.
    """

    def __init__(self, cost_u8, node2cluster, nb_off, nb_idx, order_minute, order_pickup,
                 order_delivery, n_vehicles, period=10, depth=0, neighbor_can_server=False,
                 threshold=PARITY_THRESHOLD):
        self.L = lib()
        self.cost = _c(cost_u8, np.uint8)
        self.n2c = _c(node2cluster, np.int32)
        self.nb_off = _c(nb_off, np.int32)
        self.nb_idx = _c(nb_idx, np.int32) if len(nb_idx) else np.zeros(1, np.int32)
        self.om = _c(order_minute, np.int32)
        self.op = _c(order_pickup, np.int32)
        self.od = _c(order_delivery, np.int32)
        self.nC = len(self.nb_off) - 1
        self.V = int(n_vehicles)
        self.N = len(self.om)
        self.p = int(period)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        self.h = self.L.vdso_create(self.cost.shape[0], self.nC, self.V, self.N, self.p, int(depth),
                                    int(bool(neighbor_can_server)), int(threshold),
                                    ptr(self.cost), ptr(self.n2c), ptr(self.nb_off), ptr(self.nb_idx),
                                    ptr(self.om), ptr(self.op), ptr(self.od))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.vdso_destroy(self.h)
            self.h = None

    def reset(self, veh_loc0):
        a = _c(veh_loc0, np.int32)
        assert len(a) == self.V
        self.L.vdso_reset(self.h, a.ctypes.data_as(C.c_void_p))

    def done(self):
        return bool(self.L.vdso_done(self.h))

    def update(self): self.L.vdso_update(self.h)
    def match(self): self.L.vdso_match(self.h)
    def supply_expect(self): self.L.vdso_supply_expect(self.h)
    def snapshot_pre_dispatch(self): self.L.vdso_snapshot_pre_dispatch(self.h)
    def end_tick(self): self.L.vdso_end_tick(self.h)
    def step(self): self.L.vdso_step(self.h)
    def run(self): return self.L.vdso_run(self.h)

    def dispatch(self, veh, node):
        v = _c(veh, np.int32); n = _c(node, np.int32)
        return self.L.vdso_dispatch(self.h, len(v), v.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p))

    @property
    def tick(self): return self.L.vdso_step_index(self.h)

    def _arr(self, name, n):
        p = getattr(self.L, "vdso_" + name)(self.h)
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def per_match(self): return self._arr("per_match", self.nC)
    def per_dispatch(self): return self._arr("per_dispatch", self.nC)
    def later_dispatch(self): return self._arr("later_dispatch", self.nC)
    def supply(self): return self._arr("supply", self.nC)
    def n_orders(self): return self._arr("n_orders", self.nC)
    def order_vehicle(self): return self._arr("order_vehicle", self.N)
    def order_wait(self): return self._arr("order_wait", self.N)
    def order_arrive_tick(self): return self._arr("order_arrive_tick", self.N)
    def veh_loc(self): return self._arr("veh_loc", self.V)
    def veh_dest(self): return self._arr("veh_dest", self.V)
    def veh_cluster(self): return self._arr("veh_cluster", self.V)
    def veh_order(self): return self._arr("veh_order", self.V)

    def idle_list(self, c):
        n = self.L.vdso_idle_len(self.h, c)
        if n == 0:
            return np.zeros(0, np.int32)
        return np.ctypeslib.as_array(self.L.vdso_idle_list(self.h, c), shape=(n,)).astype(np.int32)

    def idle_lists(self):
        return [self.idle_list(c) for c in range(self.nC)]

    def arrive_list(self, c):
        n = self.L.vdso_arrive_len(self.h, c)
        v = np.zeros(max(n, 1), np.int32); m = np.zeros(max(n, 1), np.int64)
        self.L.vdso_arrive_list(self.h, c, v.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p))
        return v[:n], m[:n]

    def stats(self):
        out = np.zeros(9, np.int64)
        self.L.vdso_stats(self.h, out.ctypes.data_as(C.c_void_p))
        return out
