"""The per-iteration sampler of the latent optimisation (reference core/data.py:15-73,122-176).

``select_samples`` keeps the reference's signature and consumes numpy's GLOBAL RNG in exactly
the reference's order, so a seeded run draws the same mini-batches.  ``select_sample_indices``
is the same procedure on row indices only (what the device loop needs): identical RNG calls,
identical rows, without copying 300k x 7 values per iteration."""
import numpy as np

from . import errors


def _f32(values):
    """The stored sample sets are float16 (process_sdf2.py:129-133); numpy compares halves in software, 5x slower
    than float32, and the widening is exact -- the > 0 tests pick the same rows."""
    values = np.asarray(values)
    return values.astype(np.float32) if values.dtype == np.float16 else values


def _permutation(n):
    """np.random.permutation(n) on numpy's global legacy generator, same stream and same result, with the shuffle
    done by the library's int64 loop (dj_host_mt19937_permutation) instead of numpy's byte-wise generic one."""
    import ctypes
    from deepjoin_b200 import _lib
    n = int(n)
    if n < 65536:  # below this the state hand-over (get_state / set_state, ~0.5 ms) costs more than it saves
        return np.random.permutation(n)
    state = np.random.get_state()
    if state[0] != "MT19937":
        return np.random.permutation(n)
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos = ctypes.c_int32(int(state[2]))
    out = np.empty(n, dtype=np.int64)
    _lib.check(_lib.lib().dj_host_mt19937_permutation(key.ctypes.data, ctypes.byref(pos), n, out.ctypes.data))
    np.random.set_state((state[0], key, pos.value, state[3], state[4]))
    return out


def _uniform_ratio_rows(n_pts, uniform_ratio, randomize=True):
    """Row ids kept by select_uniform_ratio_samples (core/data.py:122-176), in output order."""
    half = int(n_pts / 2)
    if uniform_ratio != 0.5:
        max_can_select = half
        surface_ends_at = half
        if uniform_ratio > 0.5:
            k = int((max_can_select * (1 - uniform_ratio)) / uniform_ratio)
            sel = _permutation(max_can_select)[:k]  # == np.random.choice(max_can_select, size=k, replace=False)
            rows = np.concatenate([sel, np.arange(surface_ends_at, surface_ends_at + max_can_select)])
        else:
            k = int((max_can_select * uniform_ratio) / (1 - uniform_ratio))
            sel = _permutation(max_can_select)[:k] + surface_ends_at  # == np.random.choice(.., size=k, replace=False)
            rows = np.concatenate([np.arange(max_can_select), sel])
    else:
        rows = np.arange(n_pts)
    if randomize:
        rows = rows[_permutation(rows.shape[0])]
    return rows


def select_sample_indices(values, num_samples=None, uniform_ratio=0.5):
    """Indices such that ``pts[idx], values[idx]`` equals the reference's
    ``select_samples(pts, values, num_samples, uniform_ratio)`` (single value column)."""
    values = _f32(values)
    if values.ndim == 2:
        if values.shape[1] != 1:
            raise NotImplementedError("multi-shape value columns")
        values = values[:, 0]
    rows = _uniform_ratio_rows(values.shape[0], uniform_ratio) if uniform_ratio is not None else np.arange(values.shape[0])
    return rows[_fair_windows(values[rows], num_samples)]


def _fair_windows(d, num_samples):
    """core/data.py:46-73 for one value column: a contiguous window of the positive rows and one of the
    non-positive rows (negatives drawn first), as positions into ``d``."""
    s = d.shape[0] if num_samples is None else num_samples
    pos_inds = np.where(d > 0)[0]
    neg_inds = np.where(d <= 0)[0]
    pos_num, neg_num = (s // 2), (s - (s // 2))
    if len(neg_inds) > neg_num:
        start_ind = np.random.randint(len(neg_inds) - neg_num)
        neg_picked = neg_inds[start_ind: start_ind + neg_num]
    else:
        try:
            neg_picked = np.random.choice(neg_inds, neg_num)
        except ValueError:
            raise errors.NoSamplesError
    if len(pos_inds) > pos_num:
        start_ind = np.random.randint(len(pos_inds) - pos_num)
        pos_picked = pos_inds[start_ind: start_ind + pos_num]
    else:
        try:
            pos_picked = np.random.choice(pos_inds, pos_num)
        except ValueError:
            raise errors.NoSamplesError
    return np.concatenate([pos_picked, neg_picked])


def select_partial_sample_indices(values, mask, num_samples=None, uniform_ratio=0.5):
    """Indices such that ``pts[idx], values[idx]`` equals the reference's
    ``select_partial_samples(pts, values, mask, num_samples, uniform_ratio)`` (core/data.py:76-119):
    ratio selection without shuffling, keep the rows where ``mask`` is set (np.intersect1d leaves them
    in ascending row order), permute, then the fair windows."""
    values = _f32(values)
    if values.ndim == 2:
        if values.shape[1] != 1:
            raise NotImplementedError("multi-shape value columns")
        values = values[:, 0]
    rows = _uniform_ratio_rows(values.shape[0], uniform_ratio, randomize=False)
    keep = np.intersect1d(rows, np.where(np.asarray(mask).reshape(-1))[0])
    keep = keep[_permutation(keep.shape[0])]
    return keep[_fair_windows(values[keep], num_samples)]


def select_partial_samples(pts, values, mask, num_samples=None, uniform_ratio=0.5):
    """core/data.py:76-119 (one value column)."""
    idx = select_partial_sample_indices(values, mask, num_samples, uniform_ratio)
    return pts[idx, :], values[idx, :]


def select_samples(pts, values, num_samples=None, uniform_ratio=0.5):
    """core/data.py:15-73 (one value column)."""
    idx = select_sample_indices(values, num_samples, uniform_ratio)
    return pts[idx, :], values[idx, :]


def clamp_samples(data, clamp_dist, skip_cols=0):
    """core/data.py:207-212."""
    data[:, skip_cols:] = np.clip(data[:, skip_cols:], -clamp_dist, clamp_dist)
    return data
