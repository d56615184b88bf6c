"""`mvn.utils.multiview` mirror for the one function on the hot path
(lib/mvn/utils/multiview.py:257-269 -> :170-228 -> :72-86)."""
from __future__ import annotations

import torch

from . import ops


def triangulate_batch_of_points_batch_version(proj_matricies_batch, points_batch,
                                              confidences_batch=None, solver='default'):
    """proj_matricies_batch (n,V,3,4), points_batch (n,V,J,2), confidences_batch (n,V,J) or
    None -> (n,J,3) float32.  `solver` ('default' = torch.svd, 'linalg' = torch.linalg.svd in
    the reference) selects between two LAPACK drivers for the same DLT null vector; here both
    map to the fp64 Jacobi kernel `mvg_triangulate` (csrc/offsets_dlt.cu)."""
    if solver not in ('default', 'linalg'):
        raise NotImplementedError(f'Please check solver: {solver}')
    if not points_batch.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    proj = proj_matricies_batch.float().contiguous()
    pts = points_batch.float().contiguous()
    conf = None if confidences_batch is None else confidences_batch.float().contiguous()
    if conf is not None and conf.shape != pts.shape[:3]:
        raise ValueError(f'Please check the size of confidences: {tuple(conf.shape)}')
    return ops.triangulate(proj, pts, conf)
