"""Block-diagonal linear algebra (reference drift/util/blockla.py).

``pinv_dm`` -- the routine ``BeamTransfer.invbeam_m`` uses (beamtransfer.py:344) -- runs
batched on the device (``dsb_pinv_batched``: one-sided Jacobi on the rows of ``[A | I]``);
the block products are plain batched matrix products.
"""

import numpy as np


def pinv_dm(matrix, rcond=None, **kwargs):
    """Pseudo-inverse of every block of ``matrix [nblocks, n, m]`` -> ``[nblocks, m, n]``
    (blockla.py:117-138; ``scipy.linalg.pinv`` semantics: singular values below
    ``rcond * sigma_max`` are dropped, default ``max(n, m) * eps``)."""
    import torch

    from .. import _lib

    if kwargs:
        raise TypeError(f"pinv_dm: unsupported arguments {sorted(kwargs)}")
    matrix = np.asarray(matrix)
    nblocks, n, m = matrix.shape
    if not torch.cuda.is_available():
        raise RuntimeError("driftscan_b200: pinv_dm needs a CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    a = torch.from_numpy(np.ascontiguousarray(matrix, dtype=np.complex128)).to(dev)
    out = torch.empty((nblocks, m, n), dtype=torch.complex128, device=dev)
    _lib.check(_lib.lib.dsb_pinv_batched(a.data_ptr(), nblocks, n, m, -1.0 if rcond is None else float(rcond),
                                         out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    res = out.cpu().numpy()
    return res if np.iscomplexobj(matrix) else res.real.astype(matrix.dtype)


def multiply_dm_v(matrix, vector, conj=False):
    """Block-diagonal matrix times vector (blockla.py:48-82)."""
    nblocks, n, m = matrix.shape
    if conj:
        if vector.shape != (nblocks, n):
            raise Exception("Shapes not compatible.")
        return np.einsum("bnm,bn->bm", matrix.conj(), vector)
    if vector.shape != (nblocks, m):
        raise Exception("Shapes not compatible.")
    return np.einsum("bnm,bm->bn", matrix, vector)


def multiply_dm_dm(matrix1, matrix2):
    """Product of two block-diagonal matrices (blockla.py:85-114)."""
    nblocks, n, m = matrix1.shape
    if matrix2.shape[:2] != (nblocks, m):
        raise Exception("Shapes not compatible.")
    return np.matmul(matrix1, matrix2)
