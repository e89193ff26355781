"""Mirror of loader.mojo: WeightLoader hands out tensors sequentially from the flat fp32 file.

Unlike the reference (loader.mojo:21-27 does no bounds check and silently reads past the end of a
short file) `next_tensor` raises when the file is exhausted.
"""
from __future__ import annotations

import numpy as np

from .whisper_tensor import Tensor


class WeightLoader:
    def __init__(self, filename: str | None = None, data: np.ndarray | None = None):
        """WeightLoader(filename) (loader.mojo:10-19); `data` lets tests pass an in-memory image."""
        if data is None:
            try:
                data = np.fromfile(filename, dtype="<f4")
            except OSError as e:  # the reference raises on open failure too (loader.mojo:11)
                raise OSError(f"WeightLoader: cannot read {filename}: {e}") from e
        self.raw_data = np.ascontiguousarray(data, np.float32).reshape(-1)
        self.size = int(self.raw_data.size)
        self.offset = 0
        self.filename = filename

    def next_tensor(self, rows: int, cols: int) -> Tensor:
        count = rows * cols
        if self.offset + count > self.size:
            raise ValueError(f"WeightLoader: file exhausted at offset {self.offset} (+{count} > {self.size})")
        t = Tensor.from_numpy(self.raw_data[self.offset:self.offset + count].reshape(rows, cols))
        self.offset += count
        return t
