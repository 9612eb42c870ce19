"""Device sessions shared by the drop-in driver functions."""
import threading

from . import _lib

_lock = threading.Lock()
_groups = {}

# Public knobs ------------------------------------------------------------------------------------
devices = None          # list of CUDA device ordinals to shard objects over; None = [current default 0]
gpu_centers = True      # centres on the device (cpg_centers); False: the reference's numpy arithmetic on the host
options = {}            # forwarded to cpg_set_option on every context (e.g. {"shells_per_pass": 4})
DEFAULT_OPTIONS = {"shells_per_pass": 0, "large_halo_min": 0, "cluster_size": 0, "radial_order": 1, "shell_groups": 0, "grid_halo_min": 0,
                   "prefix_table": 1, "upload_threads": 8, "direct_results": 1, "snapshot_cache": 1, "detect_ranges": 1, "batched_step": 0, "async_step": 2}


allow_duplicate_devices = False   # tests on a single-GPU box: several slots of a cpg_multi handle on one device


def _device_list():
    devs = [int(d) for d in (devices or [0])]
    if len(set(devs)) != len(devs) and not allow_duplicate_devices:
        raise ValueError("session.devices = %r lists a device twice: one context per device" % (devices,))
    return devs


def get_group():
    """The device group of session.devices (created on first use, kept): one Context or a MultiContext that shards
    the objects of a catalogue over the devices.  Callers hold `group.lock` for a whole set-snapshot -> compute ->
    copy-out sequence; the options are applied under that lock (see `locked_group`)."""
    devs = tuple(_device_list())
    with _lock:
        grp = _groups.get(devs)
        if grp is None:
            grp = _lib.DeviceGroup(devs)
            _groups[devs] = grp
        return grp


class locked_group:
    """with locked_group() as grp: ...   -- the session's device group with its lock held and the options applied."""

    def __enter__(self):
        self.grp = get_group()
        self.grp.lock.acquire()
        try:
            for k, v in dict(DEFAULT_OPTIONS, **options).items():
                self.grp.set_option(k, v)
        except Exception:
            self.grp.lock.release()
            raise
        return self.grp

    def __exit__(self, *exc):
        self.grp.lock.release()
        return False


def get_context(device=0):
    """The single-device Context of `device` (a group of one)."""
    with _lock:
        grp = _groups.get((int(device),))
        if grp is None:
            grp = _lib.DeviceGroup((int(device),))
            _groups[(int(device),)] = grp
    with grp.lock:
        for k, v in dict(DEFAULT_OPTIONS, **options).items():
            grp.set_option(k, v)
    return grp.single


def close_all_snapshots():
    """Forget what the devices hold (the next driver call uploads its arrays again); buffers stay allocated."""
    with _lock:
        grps = list(_groups.values())
    for grp in grps:
        with grp.lock:
            grp.set_option("snapshot_cache", 0)     # clears the cache keys
            grp.set_option("snapshot_cache", DEFAULT_OPTIONS["snapshot_cache"])


def close_all():
    with _lock:
        for grp in _groups.values():
            grp.close()
        _groups.clear()
