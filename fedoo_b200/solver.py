"""Device linear solver behind the reference's own extension point.

``Problem.set_solver`` of the reference accepts a callable ``f(A, B, **kargs) -> x`` (fedoo/core/base.py:512-513,
521-537); ``Problem.solve`` hands it the REDUCED system ``MatCB.T @ A @ MatCB`` as a host ``scipy.sparse`` matrix and
the reduced right-hand side (fedoo/core/problem.py:277-298).  ``pcg`` is such a callable: it uploads the system and
runs the Jacobi-preconditioned conjugate gradient of libfdk (``fdk_pcg_jacobi``; what the reference's
``set_solver("cg")`` does with scipy, M = diag(1 / A.diagonal())) on the GPU:

    import fedoo as fd, fedoo_b200
    fedoo_b200.install(fd)
    pb = fd.problem.Linear(assembly)
    pb.set_solver(fedoo_b200.solver.pcg, rtol=1e-10)
    pb.solve()

With ``install`` in place, ``Problem.solve`` short-cuts further when it can (adapter._problem_solve): if the matrix is
one the adapter assembled -- its values are still in HBM -- and the constraints are Dirichlet conditions and / or the
multi-point constraints PeriodicBC generates, the masked / constrained PCG runs on that device matrix directly; the
reduced system is never formed on the host and K is not uploaded
(``info["on_device_matrix"]`` says which way the last solve went).

There is no CPU implementation behind it: without the CUDA extension and a device it raises.
"""

from __future__ import annotations

import numpy as np
import torch

from .core import DeviceCSR, device

info = {"iterations": 0, "relative_residual": 0.0, "on_device_matrix": False}  # of the last call


def pcg(A, B, rtol=1e-8, maxiter=None, check_every=10, **kargs):
    """Solve ``A x = B`` (A symmetric positive definite, host scipy sparse matrix or anything ``scipy.sparse.csr_matrix``
    accepts) on the device; returns a NumPy array.  Unknown keyword arguments of other solvers are ignored."""
    from scipy import sparse

    A = sparse.csr_matrix(A)
    if not A.has_sorted_indices:
        A = A.sorted_indices()
    n = A.shape[0]
    if A.shape[1] != n:
        raise ValueError("square matrix expected")
    b = np.ascontiguousarray(np.asarray(B, dtype=np.float64).reshape(-1))
    if b.size != n:
        raise ValueError("right-hand side does not match the matrix")
    if n == 0:
        return np.zeros(0)
    dev = device()
    idx = torch.int32 if max(A.nnz, n) < 2**31 - 1 else torch.int64
    M = DeviceCSR(
        torch.from_numpy(A.indptr.astype(np.int64)).to(dev).to(idx),
        torch.from_numpy(A.indices.astype(np.int64)).to(dev).to(idx),
        torch.from_numpy(np.ascontiguousarray(A.data, dtype=np.float64)).to(dev),
        (n, n),
    )
    x, it, rel = M.pcg(torch.from_numpy(b).to(dev), rtol=rtol, maxiter=maxiter, check_every=check_every)
    info["iterations"], info["relative_residual"], info["on_device_matrix"] = it, rel, False
    if rel > rtol:
        print(f"Warning: fedoo_b200.solver.pcg: convergence to tolerance not achieved ({rel:.2e} after {it} iterations)")
    return x.cpu().numpy()
