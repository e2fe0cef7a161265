"""The reference's flag system for the hot path: the same .ini files (configparser + ExtendedInterpolation, read at
models/depression_handling.py:1428-1431 and models/dem_noise_removal.py:1316-1320) mapped onto the GPU mirrors' arguments.

`dephandle_options(config)` returns the keyword arguments of models.depression_handling.DepHandleCore exactly as the
reference's `__main__` parses them (:1444-1465: same sections, keys and getfloat / getint / getboolean conversions);
`noise_removal_options(config)` the smoothing / vegetation parameters (:1323-1333).  Keys of the reference that steer
GIS stages outside the path (river burning, walls, OSM clipping, paths) are left to the caller's own reading of the
same ConfigParser object."""
import json
from configparser import ConfigParser, ExtendedInterpolation


def load(config_file):
    """ConfigParser(allow_no_value=True, interpolation=ExtendedInterpolation()) on `config_file`, as the reference builds it"""
    config = ConfigParser(allow_no_value=True, interpolation=ExtendedInterpolation())
    if not config.read(config_file):
        raise FileNotFoundError(config_file)
    return config


def dephandle_options(config):
    """[parameters-depression-handling] (+ verbose / del_temp_files of [processes-depression-handling]) ->
    DepHandleCore(**options)"""
    s = "parameters-depression-handling"
    opts = dict(
        sfill=config.getfloat(s, "sfill"),
        fill_method=config.get(s, "fill_method"),
        carvemethod=config.get(s, "carve_method"),
        maxcost=config.getfloat(s, "maxcost"),
        radius=config.getint(s, "radius"),
        min_dist=config.getboolean(s, "min_dist"),
        fill_vol_thresh=config.getint(s, "fill_vol_thresh"),
        carve_vol_thresh=config.getint(s, "carve_vol_thresh"),
        max_carve_depth=config.getint(s, "max_carve_depth"),
        combined_fill_interval=config.getfloat(s, "combined_fill_interval"),
        flat_increment=config.getfloat(s, "flat_increment"),
    )
    p = "processes-depression-handling"
    if config.has_option(p, "verbose_print"):
        opts["verbose_print"] = config.getboolean(p, "verbose_print")
    if config.has_option(p, "del_temp_files"):
        opts["del_temp_files"] = config.getboolean(p, "del_temp_files")
    return opts


def burn_options(config):
    """the stream-burning scalars handed to run_depression_handling per unit (:1489-1495)"""
    s = "parameters-depression-handling"
    return dict(SWO_cutoff=config.getint(s, "SWO_cutoff"), initial_burn=config.getfloat(s, "initial_burn"),
                SWO_burn=config.getfloat(s, "SWO_burn"), burn_cutoff=config.getfloat(s, "burn_cutoff"),
                gap_dist=config.getint(s, "gap_dist"))


def noise_removal_options(config):
    """[parameters-noise-removal] / [processes-noise-removal] -> arguments of stencils.smooth_core and veg
    (models/dem_noise_removal.py:1323-1333)"""
    s = "parameters-noise-removal"
    return dict(
        filtertype=config.get(s, "filtertype"),
        vegmodel=config.get(s, "vegmodel"),
        SWOlevel=config.getint(s, "SWOlevel"),
        curvebreaks=json.loads(config.get(s, "curvebreaks")),
        slopebreaks=json.loads(config.get(s, "slopebreaks")),
        base_DEM=config.get(s, "base_DEM"),
        save_veg_files=config.get(s, "save_veg_files"),
        veg_remove=config.getboolean("processes-noise-removal", "veg_remove"),
    )


def filter_weights(filtertype):
    """CreateDTM.setfweights (models/dem_noise_removal.py:1236-1280): weights of the 9x9, 7x7, 5x5, 3x3 smoothers"""
    if filtertype == "DW":
        return [[1., 0.8, 0.6, 0.4, 0.2], [1., 0.75, 0.5, 0.25], [1., 0.7, 0.35], [1., 0.5]]
    if filtertype == "mean":
        return [[1., 1., 1., 1, 1.], [1., 1., 1., 1.], [1., 1., 1.], [1., 1.]]
    raise ValueError("filtertype must be 'DW' or 'mean'")
