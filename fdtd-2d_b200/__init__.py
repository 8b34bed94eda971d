"""fdtd2d_b200 -- B200-native 2D FDTD (TM-mode Yee grid) time stepping behind the reference's Python
call surface (skunnavakkam/fdtd-2d, python-src/fdtd.py + main.py).

    from fdtd2d_b200 import grid_init, material_init, update_Hx_Hy, update_Ez, ricker   # drop-in names
    from fdtd2d_b200 import Simulation                                                   # device-resident

All compute runs in libfdtd2d.so (hand-written CUDA for sm_100a, C ABI in include/fdtd2d.h).
There is no CPU fallback.
"""
from . import _lib

DEFAULT_K = 8  # leapfrog steps per HBM round trip when the caller does not choose (fp32; fp64 uses 4)

from ._lib import EXPORTED_SYMBOLS, Fdtd2dError
from .api import (EPSILON0, MU0, capture_snapshot, grid_init, material_init, release_handles, ricker, sinusoidal, update_Ez,
                  update_Hx_Hy)
from .build import LIB_PATH, build
from .distributed import HaloExchange, InProcessSlabs, SlabSimulation, exchange_peer_blobs, slab_rows
from . import dataset, driver, snapshot, structure
from .structure import RegionDrawer
from .dataset import generate_data
from .snapshot import eps_background, make_video_from_frames, seismic_lut
from .simulation import (Simulation, courant_number, ricker_amplitude, sinusoidal_amplitude, source_table)

__all__ = [
    "grid_init", "material_init", "update_Hx_Hy", "update_Ez", "ricker", "sinusoidal", "Simulation",
    "capture_snapshot", "make_video_from_frames", "seismic_lut", "eps_background",
    "courant_number", "ricker_amplitude", "sinusoidal_amplitude", "source_table", "build", "LIB_PATH",
    "Fdtd2dError", "EXPORTED_SYMBOLS", "generate_data", "dataset", "driver", "structure", "RegionDrawer", "SlabSimulation", "InProcessSlabs", "HaloExchange", "slab_rows", "exchange_peer_blobs", "DEFAULT_K", "EPSILON0", "MU0", "release_handles",
]
