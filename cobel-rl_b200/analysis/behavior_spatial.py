"""Spatial behaviour analysis (reference: analysis/behavior_spatial.py:9-111).

``get_occupancy_map`` is a histogram over all visited positions; it runs on the device of the
trajectories it is given (the ``[N, T, 2]`` position buffers of a batched run never have to leave the
GPU) and reproduces ``numpy.histogram2d`` bin for bin: bin edges from ``numpy.linspace``, values
binned with ``searchsorted(edges, x, 'right') - 1``, the right-most edge belonging to the last bin,
values outside the range dropped.
"""
import numpy as np
import torch


def get_occupancy_map(trajectories, width, height, bin_size, margins='expand'):
    """analysis/behavior_spatial.py:9-73.  ``trajectories``: a list of ``[T, 2]`` arrays / tensors, or one
    ``[..., 2]`` tensor of positions (e.g. ``[N, T, 2]`` for a batch; rows may be padded with NaN, which falls
    outside every bin).  As in the reference, column 0 is binned along the first (height) axis."""
    assert width > 0 and height > 0, 'Invalid environment dimensions! Dimensions must be positive!'
    assert bin_size > 0 and bin_size <= min(width, height), \
        'Invalid bin size! Bin size must be positive and less than environmental dimensions!'
    assert margins in ['expand', 'include', 'ignore'], \
        "Invalid handling mode for margins! Must be 'expand', 'include' or 'ignore'!"
    bins = np.array([int(height / bin_size), int(width / bin_size)])
    bins += ((np.array([height, width]) - bins * bin_size) > 0.0) * (margins == 'expand')
    if isinstance(trajectories, (list, tuple)):
        parts = [torch.as_tensor(np.asarray(t) if not isinstance(t, torch.Tensor) else t) for t in trajectories]
        for t in parts:
            assert t.dim() == 2 and t.shape[1] == 2, 'Trajectories must be 2-dimensional!'
        pts = torch.cat(parts, dim=0) if parts else torch.zeros((0, 2), dtype=torch.float64)
    else:
        pts = torch.as_tensor(trajectories)
        assert pts.shape[-1] == 2, 'Trajectories must be 2-dimensional!'
        pts = pts.reshape(-1, 2)
    pts = pts.to(torch.float64)
    finite = pts[~torch.isnan(pts).any(dim=1)]
    assert finite.numel() == 0 or float(finite.min()) >= 0.0, 'Invalid coordinates! Coordinates must be non-negative!'
    idx = []
    for d in range(2):
        hi = float(bins[d] * bin_size)
        x = pts[:, d].clamp(min=0.0)
        if margins == 'include':
            x = x.clamp(max=hi)
        edges = torch.as_tensor(np.linspace(0.0, hi, int(bins[d]) + 1), dtype=torch.float64, device=pts.device)
        b = torch.searchsorted(edges, x.contiguous(), right=True) - 1
        b = torch.where(x == edges[-1], torch.full_like(b, int(bins[d]) - 1), b)
        b = torch.where(torch.isnan(pts[:, d]), torch.full_like(b, -1), b)
        idx.append(b)
    ok = (idx[0] >= 0) & (idx[0] < int(bins[0])) & (idx[1] >= 0) & (idx[1] < int(bins[1]))
    flat = (idx[0] * int(bins[1]) + idx[1])[ok]
    occ = torch.bincount(flat, minlength=int(bins[0] * bins[1])).to(torch.float64)
    return occ.reshape(int(bins[0]), int(bins[1]))


def match(sequence, template):
    """analysis/behavior_spatial.py:76-111: for every start position of ``sequence`` the number of elements of
    ``template`` that match when it is laid over the sequence from there on (cut off at the end)."""
    seq, tpl = np.asarray(sequence), np.asarray(template)
    n, m = seq.shape[0], tpl.shape[0]
    out = np.zeros(n, dtype=np.int64)
    for j in range(min(m, n)):
        out[:n - j] += seq[j:] == tpl[j]
    return out
