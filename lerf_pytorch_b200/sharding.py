"""Multi-GPU work partition of the LUT inference path: no collective on the data path.

The reference is single-process (resample/eval_lut_sr.py:489-512 walks the files serially); every output pixel
depends on a bounded input neighbourhood, so the path shards two ways:

* per image  -- images go round-robin to ranks, each rank holds its own LUT copy (cfg-2/3/4);
* row bands  -- rank g computes output rows [oy0, oy1) of ONE frame (cfg-5).  It needs the input rows its taps
  touch (resize_right2d_numpy.py:82-98: ``left``, ``left + 1``) plus 3 rows for the stage-2 reach of modes c/t in all
  rotations plus 3 rows for stage 1 (eval_lut_sr.py:12-18, :541-628) = a 7-row halo; only true image edges clamp.

This module is host logic only (numpy); ``LerfSR(..., rows=(oy0, oy1))`` / ``lerf_sr_fused(oy0, oy1)`` do the work.
``torch.distributed`` is used by callers for barriers and timing reductions, never for pixels.
"""
import numpy as np

from .resize_right2d import sr_axis_tables

STAGE_REACH = 3  # modes c and t read 3 pixels away, in every rotation (eval_lut_sr.py:30-81)


def image_shard(n_images, rank, world):
    """Indices of the images rank ``rank`` of ``world`` processes (round-robin, like a work list)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    return list(range(rank, n_images, world))


def row_bands(out_rows, world, align=1):
    """``world`` contiguous output row bands [(oy0, oy1), ...] covering [0, out_rows); band edges are multiples of
    ``align`` (use the integer scale so a band starts on a cell boundary).  Trailing bands may be empty."""
    if world < 1 or align < 1:
        raise ValueError("bad world/align")
    units = (out_rows + align - 1) // align
    bands = []
    for g in range(world):
        u0 = (units * g) // world
        u1 = (units * (g + 1)) // world
        bands.append((min(u0 * align, out_rows), min(u1 * align, out_rows)))
    return bands


def band_input_rows(in_rows, out_rows, scale, oy0, oy1, support_sz=2):
    """Input rows [r0, r1) that output rows [oy0, oy1) depend on through resampling, stage 2 and stage 1.

    Returns (r0, r1, c0, c1): [c0, c1) are the rows whose hyper codes / features the taps read, [r0, r1) adds the
    stage-2 and stage-1 reach; both are clamped to the image (only true edges clamp)."""
    if not 0 <= oy0 < oy1 <= out_rows:
        raise ValueError("bad band [%d,%d) of %d" % (oy0, oy1, out_rows))
    left, _, _ = sr_axis_tables(in_rows, out_rows, scale, support_sz)
    c0 = int(np.clip(left[oy0], 0, in_rows - 1))
    c1 = int(np.clip(left[oy1 - 1] + support_sz - 1, 0, in_rows - 1)) + 1
    r0 = max(c0 - 2 * STAGE_REACH, 0)
    r1 = min(c1 + 2 * STAGE_REACH, in_rows)
    return r0, r1, c0, c1


def band_halo_rows():
    """Rows of real neighbours a band needs on each side beyond its own taps: 3 (stage 2) + 3 (stage 1); the tap
    footprint itself reaches one more row, hence SURVEY.md's "7-row halo"."""
    return 2 * STAGE_REACH + 1
