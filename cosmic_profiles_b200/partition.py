"""Particle-count-balanced partition of objects over GPUs (no collective: objects are independent,
SURVEY.md section 8(e)).  Greedy longest-processing-time assignment on cost = N_h * R."""
import numpy as np


def lpt_partition(cost, n_parts):
    """Return a list of n_parts index arrays (ascending object ids) with balanced total cost."""
    cost = np.asarray(cost, dtype=np.float64)
    order = np.argsort(-cost, kind="stable")
    loads = np.zeros(n_parts)
    owner = np.empty(len(cost), dtype=np.int64)
    for h in order:
        p = int(np.argmin(loads))
        owner[h] = p
        loads[p] += cost[h]
    return [np.flatnonzero(owner == p) for p in range(n_parts)]


def sub_catalogue(idx_cat, obj_size, objs):
    """Index catalogue restricted to `objs` (catalogue order kept inside each object)."""
    off = np.concatenate(([0], np.cumsum(np.asarray(obj_size, dtype=np.int64))))
    if len(objs) == 0:
        return np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32)
    parts = [idx_cat[off[h]:off[h + 1]] for h in objs]
    return np.concatenate(parts).astype(np.int32, copy=False), np.asarray(obj_size)[objs].astype(np.int32)
