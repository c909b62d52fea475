"""ctypes binding of libvds.so -- the C ABI declared in include/vds.h.

The CUDA library is the product; there is no CPU fallback.  If the shared
object is missing this module raises at import of the engine, loudly.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libvds.so")
SRC_PATH = os.path.join(_HERE, "csrc", "vds.cu")
HDR_PATH = os.path.join(_ROOT, "include", "vds.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

VDS_NUM_STATS = 10
PARITY_THRESHOLD = 600_000_000_000
STAT_NAMES = ("OrderNum", "RejectNum", "TotallyWaitTime", "SumOrderValue", "DispatchNum",
              "TotallyDispatchCost", "Lookups", "Arrivals", "Matches", "Ticks")


class VdsError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("replicas", C.c_int32), ("vehicles", C.c_int32), ("clusters", C.c_int32),
                ("nodes", C.c_int32), ("max_orders", C.c_int32), ("ticks", C.c_int32),
                ("period_min", C.c_int32), ("depth_limit", C.c_int32),
                ("neighbor_can_server", C.c_int32), ("order_replicas", C.c_int32),
                ("max_orders_per_tick", C.c_int32), ("device", C.c_int32),
                ("reject_threshold", C.c_int64)]


class Static(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("cost_u8", "node2cluster", "search_off", "search_idx",
                                          "reach_off", "reach_idx")]


class Orders(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("order_pd", "order_value", "tick_off", "value_total",
                                          "sorted_pd", "sorted_idx", "cluster_off", "tick_value")]


class SearchNodes(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n in ("node_rank", "cluster_base", "own_list", "search_list", "node_count", "run_end",
                                           "node_count_exact", "slot_vehicle", "slot_key", "head_key")]
                + [("ranks_padded", C.c_int32), ("own_pitch", C.c_int32), ("search_pitch", C.c_int32)])


STATE_FIELDS = ("veh_loc", "veh_cluster", "veh_arrive", "veh_dest", "veh_key", "order_res",
                "per_match", "per_dispatch", "idle_live", "supply", "n_orders", "stats",
                "idle_ent", "idle_off", "bucket_off", "bucket_ord", "disp_seq", "trace")


class State(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in STATE_FIELDS]


# every symbol include/vds.h declares (tests check the export list against this)
EXPORTS = ("vds_abi_version", "vds_padded_vehicles", "vds_create", "vds_destroy", "vds_last_error",
           "vds_bind_static", "vds_bind_orders", "vds_bind_state", "vds_compute_order_values",
           "vds_prepare_orders", "vds_reset", "vds_clear_results", "vds_update", "vds_match", "vds_supply_expect", "vds_dispatch", "vds_dispatch_strided", "vds_policy_random", "vds_rollout_policy_random", "vds_rollout", "vds_tick", "vds_rollout_is_fused", "vds_rollout_threads",
           "vds_stats", "vds_sync", "vds_launch_count", "vds_generate_orders", "vds_generate_prepared_orders", "vds_generate_placement",
           "vds_rollout_kernel_name", "vds_padded_nodes", "vds_bind_cluster_nodes", "vds_bind_queues", "vds_bind_observations", "vds_observe", "vds_time_features", "vds_load_orders", "vds_cluster_cost_sums",
           "vds_bind_search_nodes", "vds_search_nodes_active")


def build(force=False, verbose=False):
    """nvcc cross-compiles sm_100a without a GPU; output stays in-tree."""
    deps = [SRC_PATH, HDR_PATH] + [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps)):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + ["-o", LIB_PATH, SRC_PATH]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("VDS_LIB_PATH", LIB_PATH)       # developer knob: A/B-test a kernel variant built elsewhere
    if not os.path.exists(path):
        raise VdsError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                       "The dispatch hot path is CUDA-only; there is no CPU fallback.")
    L = C.CDLL(path)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    sig = {
        "vds_abi_version": (C.c_int, []),
        "vds_padded_vehicles": (C.c_int, [i32]),
        "vds_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
        "vds_destroy": (C.c_int, [vp]),
        "vds_last_error": (C.c_char_p, [vp]),
        "vds_bind_static": (C.c_int, [vp, C.POINTER(Static)]),
        "vds_bind_orders": (C.c_int, [vp, C.POINTER(Orders)]),
        "vds_bind_state": (C.c_int, [vp, C.POINTER(State)]),
        "vds_compute_order_values": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "vds_prepare_orders": (C.c_int, [vp, vp, vp]),
        "vds_reset": (C.c_int, [vp, vp, vp]),
        "vds_clear_results": (C.c_int, [vp, vp]),
        "vds_update": (C.c_int, [vp, i32, vp]),
        "vds_match": (C.c_int, [vp, i32, vp]),
        "vds_supply_expect": (C.c_int, [vp, i32, vp]),
        "vds_dispatch": (C.c_int, [vp, i32, vp, vp, vp, i32, vp]),
        "vds_dispatch_strided": (C.c_int, [vp, i32, vp, vp, vp, i32, vp]),
        "vds_policy_random": (C.c_int, [vp, i32, u64, i64, C.c_uint32, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
        "vds_rollout_policy_random": (C.c_int, [vp, i32, i32, u64, i64, C.c_uint32, vp, vp, vp, vp, vp]),
        "vds_rollout": (C.c_int, [vp, i32, i32, vp]),
        "vds_tick": (C.c_int, [vp, i32, vp]),
        "vds_rollout_is_fused": (C.c_int, [vp]),
        "vds_rollout_threads": (C.c_int, [vp]),
        "vds_rollout_kernel_name": (C.c_char_p, [vp]),
        "vds_padded_nodes": (C.c_int, [i32]),
        "vds_bind_cluster_nodes": (C.c_int, [vp, vp, vp, vp, i32]),
        "vds_bind_queues": (C.c_int, [vp, vp, vp]),
        "vds_stats": (C.c_int, [vp, vp, vp]),
        "vds_sync": (C.c_int, [vp, vp]),
        "vds_launch_count": (i64, [vp]),
        "vds_generate_orders": (C.c_int, [vp, u64, i64, vp, vp, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp]),
        "vds_generate_prepared_orders": (C.c_int, [vp, u64, i64, vp, vp, i32, i32, vp, vp, i32, vp, vp, vp, vp]),
        "vds_generate_placement": (C.c_int, [vp, u64, i64, vp, i32, vp, vp]),
        "vds_bind_observations": (C.c_int, [vp, vp, i32]),
        "vds_observe": (C.c_int, [vp, i32, vp]),
        "vds_time_features": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp, vp]),
        "vds_load_orders": (C.c_int, [vp, vp, vp, vp, i32, i32, vp, vp]),
        "vds_cluster_cost_sums": (C.c_int, [vp, vp, vp, vp, vp]),
        "vds_bind_search_nodes": (C.c_int, [vp, C.POINTER(SearchNodes)]),
        "vds_search_nodes_active": (C.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L
