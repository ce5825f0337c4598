"""cs420-ray-tracer_b200 -- B200-native (sm_100a) drop-in for the hot path of
shininglegend/cs420-ray-tracer (camera rays -> ray/sphere closest hit -> Phong + shadow rays
-> reflection bounces -> 8-bit RGB).

The product is the C-ABI library ``librt_b200.so`` (``include/rt_b200.h``) and the drop-in
binary ``ray_cuda``; this Python package is only the ctypes mirror of that ABI used by the
tests and by ``bench.py``.  There is no CPU fallback: rendering raises ``RtError`` when the
CUDA library or a B200 is missing.

The directory name contains a hyphen, so import it through ``rtb200.py`` at the repo root
(``import rtb200``) or ``importlib`` (see ``__graft_entry__.py``).
"""
from .api import (  # noqa: F401
    RtError, RtStats, Scene, Renderer, MultiRenderer, load_library, library_path, load_scene, write_ppm,
    band_rows, band_row_list, measure_fp32_peak, host_register, host_unregister, ABI_SYMBOLS,
)
from .ppmtools import read_ppm, compare_rgb, ppm_text  # noqa: F401
from .build import build_all, build_library  # noqa: F401
from .bands import BandGather, PeerFrame  # noqa: F401
