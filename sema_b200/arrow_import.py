"""Importer from the reference's Arrow vector column.

The reference builds ``FixedSizeListArray<Float32, 384>`` (nullable: a failed embedding is a
null row) at ``src/storage/lance_indexer.rs:41-45, 75-76`` and hands it to LanceDB.  This module
takes the same array (pyarrow) and feeds its contiguous values buffer + validity to
``sema_index_append`` without copying the values.
"""
from __future__ import annotations

import numpy as np


def fixed_size_list_to_rows(arr):
    """pyarrow FixedSizeListArray<float32, dim> -> (rows float32[n, dim] view, valid uint8[n] or None)."""
    import pyarrow as pa
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks()
    if not pa.types.is_fixed_size_list(arr.type) or not pa.types.is_float32(arr.type.value_type):
        raise TypeError(f"expected FixedSizeList<float32, dim>, got {arr.type}")
    dim = arr.type.list_size
    n = len(arr)
    values = arr.values  # child array covering [offset*dim, (offset+n)*dim) is what flatten semantics give
    flat = arr.flatten() if arr.null_count == 0 else None
    if flat is not None and len(flat) == n * dim:
        rows = flat.to_numpy(zero_copy_only=True).reshape(n, dim)
    else:
        # null rows still occupy dim slots in the child array; slice it by the parent's offset
        child = values.slice(arr.offset * dim, n * dim)
        rows = np.nan_to_num(child.to_numpy(zero_copy_only=False), nan=0.0).astype(np.float32, copy=False).reshape(n, dim)
    valid = None
    if arr.null_count:
        valid = np.asarray(arr.is_valid().to_numpy(zero_copy_only=False), dtype=np.uint8)
    return rows, valid


def append_arrow(index, arr, normalize: bool = False) -> int:
    """Append a FixedSizeList<float32, dim> vector column to a GpuIndex; returns the first row."""
    rows, valid = fixed_size_list_to_rows(arr)
    if rows.shape[1] != index.dim:
        raise ValueError(f"column has dim {rows.shape[1]}, index has {index.dim}")
    return index.append(rows, valid=valid, normalize=normalize)
