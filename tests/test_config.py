"""The reference's .ini flag system mapped onto the mirrors (pydrodem_b200/config.py).  The text below reproduces the
hot-path sections of config_files/config_local.ini (values as shipped upstream)."""
import numpy as np

from pydrodem_b200 import config as cfg

INI = """
[paths]
base_path = /work/project
WBT_path  = ${paths:base_path}/whitebox-tools

[parameters-noise-removal]
base_DEM        = TDF
filtertype      = DW
vegmodel        = 3Dtable
SWOlevel        = 10
curvebreaks     = [1.0, 2.5]
slopebreaks     = [0, 2, 6, 15, 30]
save_veg_files  = bias

[processes-noise-removal]
veg_remove      = True

[parameters-depression-handling]
buffer           = 200
fill_method      = WL
carve_method     = SC
min_dist         = False
sfill                  = 1.0
flat_increment         = 0.0
maxcost                = 200
radius                 = 50
fill_vol_thresh        = 12000
carve_vol_thresh       = 4000
max_carve_depth        = 100
combined_fill_interval = .2
SWO_cutoff          = 20
initial_burn        = 0.0
SWO_burn            = 1.0
burn_cutoff         = 6.0
gap_dist            = 60

[processes-depression-handling]
verbose_print   = False
del_temp_files  = True
"""


def test_sections_map_onto_the_mirrors(tmp_path):
    p = tmp_path / "config_local.ini"
    p.write_text(INI)
    c = cfg.load(str(p))
    assert c.get("paths", "WBT_path") == "/work/project/whitebox-tools"          # ExtendedInterpolation
    d = cfg.dephandle_options(c)
    assert d == dict(sfill=1.0, fill_method="WL", carvemethod="SC", maxcost=200.0, radius=50, min_dist=False,
                     fill_vol_thresh=12000, carve_vol_thresh=4000, max_carve_depth=100, combined_fill_interval=0.2,
                     flat_increment=0.0, verbose_print=False, del_temp_files=True)
    from pydrodem_b200.models.depression_handling import DepHandleCore
    import inspect
    assert set(d) <= set(inspect.signature(DepHandleCore.__init__).parameters)   # every key is an argument of the core
    n = cfg.noise_removal_options(c)
    assert n["curvebreaks"] == [1.0, 2.5] and n["slopebreaks"] == [0, 2, 6, 15, 30] and n["SWOlevel"] == 10
    assert n["vegmodel"] == "3Dtable" and n["veg_remove"] is True and n["filtertype"] == "DW"
    assert cfg.burn_options(c) == dict(SWO_cutoff=20, initial_burn=0.0, SWO_burn=1.0, burn_cutoff=6.0, gap_dist=60)


def test_filter_weights_build_the_reference_kernels():
    from pydrodem_b200 import stencils
    w = cfg.filter_weights("DW")
    assert [2 * len(x) - 1 for x in w] == [9, 7, 5, 3]
    k5 = stencils.build_kernel(w[2], dtype=np.float32)
    assert k5[2, 2] == 1.0 and k5[1, 1] == np.float32(0.7) and k5[0, 0] == np.float32(0.35)   # the docstring's 5x5 example
    assert np.all(stencils.build_kernel(cfg.filter_weights("mean")[3], dtype=np.float32) == 1.0)


def test_dephandle_core_defaults_are_the_reference_ini():
    """DepHandleCore() without arguments must condition with the reference's own parameters
    (config_files/config_local.ini [parameters-depression-handling]; fixture from tests/golden/make_config_golden.py)."""
    import inspect
    import json
    import os
    from pydrodem_b200.models.depression_handling import DepHandleCore
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "config_local_dephandle.json")) as f:
        ref = json.load(f)
    sig = inspect.signature(DepHandleCore.__init__).parameters
    for key, val in ref.items():
        if key in ("verbose_print", "del_temp_files"):
            continue
        assert key in sig, key
        assert sig[key].default == val, (key, sig[key].default, val)
