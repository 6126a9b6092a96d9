"""physim's state records as numpy dtypes.

Layout contract: physim-core/src/lib.rs:16-37 (`#[repr(C)] Entity`, 80 B; `Acceleration`, 24 B),
C mirror c_plugin/physim.h:37-54.
"""
import numpy as np

ENTITY = np.dtype(
    [("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("vx", "<f8"), ("vy", "<f8"), ("vz", "<f8"),
     ("radius", "<f8"), ("mass", "<f8"), ("id", "<u8"), ("fixed", "?")],
    align=True,
)
ACCELERATION = np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8")], align=True)

assert ENTITY.itemsize == 80 and ENTITY.fields["id"][1] == 64 and ENTITY.fields["fixed"][1] == 72
assert ACCELERATION.itemsize == 24


def entities(n: int) -> np.ndarray:
    """Zeroed Entity array (Entity::default(): everything 0 / false)."""
    return np.zeros(n, dtype=ENTITY)


def accelerations(n: int) -> np.ndarray:
    return np.zeros(n, dtype=ACCELERATION)
