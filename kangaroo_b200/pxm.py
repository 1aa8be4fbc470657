"""On-disk outputs of the stereo applications: host-side file writers, numpy only.

* ``SavePXM`` -- include/kangaroo/extra/SavePPM.h:20-39: header ``<type>\\n<w> <h>\\n<num_colors>\\n`` followed by the
  raw rows (``w * sizeof(T)`` bytes each, no pitch padding), for images; ``<w> <h> <d>`` for volumes (:46-58).
* ``SavePDM`` -- applications/stereo/main.cpp:400-410: the ``.pdm`` depth/disparity map,
  ``P7\\n<cols> <rows>\\n4294967295\\n`` + raw float32 data.
``LoadPXM`` reads either back (what a user would diff against).
"""
from __future__ import annotations

import numpy as np


def _host(a) -> np.ndarray:
    return a.numpy() if hasattr(a, "numpy") and not isinstance(a, np.ndarray) else np.asarray(a)


def SavePXM(filename: str, image, ppm_type: str = "P5", num_colors: int = 255) -> None:
    """2-D (h, w[, c]) array / roo.Image -> image file; 3-D volumes use SaveVolumePXM."""
    a = np.ascontiguousarray(_host(image))
    h, w = a.shape[:2]
    with open(filename, "wb") as f:
        f.write(f"{ppm_type}\n{w} {h}\n{num_colors}\n".encode("ascii"))
        f.write(a.tobytes())


def SaveVolumePXM(filename: str, vol, ppm_type: str = "P5", num_colors: int = 255) -> None:
    """(d, h, w) array / roo.Volume -> SavePPM.h:46-58 layout (d outermost, rows contiguous)."""
    a = np.ascontiguousarray(_host(vol))
    d, h, w = a.shape[:3]
    with open(filename, "wb") as f:
        f.write(f"{ppm_type}\n{w} {h} {d}\n{num_colors}\n".encode("ascii"))
        f.write(a.tobytes())


def SavePDM(filename: str, dmap) -> None:
    a = np.ascontiguousarray(_host(dmap), dtype=np.float32)
    rows, cols = a.shape
    with open(filename, "wb") as f:
        f.write(f"P7\n{cols} {rows}\n4294967295\n".encode("ascii"))
        f.write(a.tobytes())


def LoadPXM(filename: str, dtype=None) -> np.ndarray:
    """Reads a file written by SavePXM / SaveVolumePXM / SavePDM.  dtype defaults to uint8 (P5), float32 (P7)."""
    with open(filename, "rb") as f:
        kind = f.readline().strip().decode("ascii")
        dims = [int(t) for t in f.readline().split()]
        f.readline()   # num_colors
        data = f.read()
    dt = np.dtype(dtype if dtype is not None else (np.float32 if kind == "P7" else np.uint8))
    shape = (dims[1], dims[0]) if len(dims) == 2 else (dims[2], dims[1], dims[0])
    n = int(np.prod(shape))
    a = np.frombuffer(data, dtype=dt)
    if a.size != n:   # multi-channel pixels
        return a.reshape(*shape, a.size // n).copy()
    return a.reshape(shape).copy()
