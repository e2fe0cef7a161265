"""pydrodem_b200 — B200-native (sm_100a) DEM hydro-conditioning core behind pydrodem's call boundary.

Hot path: depression fill -> D8 direction/slope -> flat resolution -> D8 accumulation, plus
connected-component labelling, all hand-written CUDA in libhydrocore.so reached through the C ABI of
include/hydrocore.h.  No CPU fallback.
"""
from ._lib import (DIR_NODATA, DIR_NOFLOW, SEED_TAUDEM, SEED_WBT, TABLE_TAUDEM, TABLE_WBT,  # noqa: F401
                   HydrocoreError)
from .core import (accum, breach_depressions, breach_depressions_least_cost, breach_single_cell_pits, d8, empty_raster, fill, fill_d8_accum, fill_eps, fill_single_cell_pits, label,  # noqa: F401
                   resolve_flats, taudem_check, taudem_fix)

__version__ = "0.1.0"
