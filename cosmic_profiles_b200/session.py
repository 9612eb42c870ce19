"""Device sessions shared by the drop-in driver functions."""
import threading

from . import _lib

_lock = threading.Lock()
_contexts = {}

# Public knobs ------------------------------------------------------------------------------------
devices = None          # list of CUDA device ordinals to shard objects over; None = [current default 0]
gpu_centers = True      # centres on the device (cpg_centers); False: the reference's numpy arithmetic on the host
options = {}            # forwarded to cpg_set_option on every context (e.g. {"shells_per_pass": 4})
DEFAULT_OPTIONS = {"shells_per_pass": 0, "large_halo_min": 0, "cluster_size": 0, "radial_order": 1, "shell_groups": 0, "grid_halo_min": 0,
                   "prefix_table": 1, "upload_threads": 8, "direct_results": 1}


def get_context(device=0):
    with _lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = _lib.Context(device)
            _contexts[device] = ctx
        for k, v in dict(DEFAULT_OPTIONS, **options).items():
            ctx.set_option(k, v)
        return ctx


def close_all():
    with _lock:
        for ctx in _contexts.values():
            ctx.close()
        _contexts.clear()
