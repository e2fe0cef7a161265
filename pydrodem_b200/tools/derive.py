"""`tools/derive.py` under its own name, so that `import tools.derive as pdt` becomes
`import pydrodem_b200.tools.derive as pdt`: the stencil functions live in pydrodem_b200.stencils (CUDA kernels behind
the reference's signatures), `DEMdifference` (:22-63) is file plumbing over raster_io."""
import os

from .. import raster_io
from ..stencils import apply_convolve_filter, curvature_D2, slope_D2, slope_D4  # noqa: F401
from . import convert as pyct


def DEMdifference(orig, new, outfile, nodata):
    """new - orig of two GeoTIFFs of the same grid, written next to `new` as `outfile` (tools/derive.py:22-63)"""
    output_path = os.path.join(os.path.dirname(new), outfile)
    orig_array, _ = raster_io.read_raster(orig)
    new_array, _ = raster_io.read_raster(new)
    diff_array = new_array - orig_array
    pyct.npy2tif(diff_array, orig, output_path, nodata)
    return diff_array
