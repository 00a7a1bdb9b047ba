"""TEST INFRASTRUCTURE (oracle): the reference's coefficient-assignment functions restated in NumPy, statement by
statement, from src/dolfinx_external_operator/external_operator.py:286-335.  The reference package itself cannot
be imported here (it needs dolfinx/ufl/basix at import time), and these functions are plain NumPy indexing, so the
restatement IS the reference arithmetic: NumPy's own fancy-assignment rule (last write wins) defines the result.
Only tests/ may import this module."""

from __future__ import annotations

import numpy as np


def assign_non_mixed(x_array: np.ndarray, unrolled_dofmap: np.ndarray, values: np.ndarray) -> None:
    x_array[unrolled_dofmap] = values  # :286-287


def assign_non_mixed_contiguous(x_array: np.ndarray, values: np.ndarray) -> None:
    x_array[:] = values  # :289-290


def assign_mixed_2d(x_array: np.ndarray, info_list, n_points_total: int, values: np.ndarray) -> None:
    if values.ndim == 1:  # :298-300
        n_cells = values.size // n_points_total
        values = values.reshape(n_cells, n_points_total)
    for info in info_list:  # :303-311
        offset, n_pts, flat_dofs = info["offset"], info["n_pts"], info["flat_dofs"]
        block = values[:, offset : offset + n_pts]
        x_array[flat_dofs] = block.reshape(-1)


def assign_mixed_3d(x_array: np.ndarray, info_list, n_points_total: int, comp_size: int, values: np.ndarray) -> None:
    if values.ndim == 1:  # :319-321
        n_cells = values.size // (n_points_total * comp_size)
        values = values.reshape(n_cells, n_points_total, comp_size)
    n_cells = values.shape[0]
    for info in info_list:  # :324-335
        offset, n_pts, flat_dofs = info["offset"], info["n_pts"], info["flat_dofs"]
        dofs_per_cell, val_size = info["dofs_per_cell"], info["val_size"]
        chunk = values[:, offset : offset + n_pts, :]
        block = chunk[:, :, :val_size].reshape(n_cells, dofs_per_cell)
        x_array[flat_dofs] = block.reshape(-1)
