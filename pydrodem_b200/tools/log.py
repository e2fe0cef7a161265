"""Unit discovery and log lines of the reference's `tools/log.py` (find_all_ids :95-106, find_unfinished_ids :66-92,
find_all_basins_IDs :189-195, log_time): which `<unit>_<id>` directories of a work directory still need processing."""
import glob
import os
import time


def _unit_dirs(work_dir, dir_search):
    return sorted(d for d in glob.glob(os.path.join(work_dir, f'*{dir_search}*')) if os.path.isdir(d))


def _unit_id(path, unit, dir_search):
    name = os.path.splitext(os.path.basename(path))[0].replace(dir_search, '')
    return int(name) if unit == 'basin' else name


def find_all_ids(work_dir, unit='tile', dir_search=''):
    """ids of every `<dir_search><id>` directory: ints for basins, strings for tiles"""
    return [_unit_id(d, unit, dir_search) for d in _unit_dirs(work_dir, dir_search)]


def find_unfinished_ids(work_dir, unit='tile', dir_search='', tifname='DTM.tif', id_list=None):
    """ids of the unit directories that do not hold `tifname` yet (the unit's completion marker)"""
    if id_list is None:
        dirs = _unit_dirs(work_dir, dir_search)
    else:
        dirs = [glob.glob(os.path.join(work_dir, f'*{i}*'))[0] for i in id_list]
    return [_unit_id(d, unit, dir_search) for d in dirs if not os.path.exists(os.path.join(d, tifname))]


def find_all_basins_IDs(work_dir, dir_search='basin_*'):
    return [int(os.path.basename(b).replace('basin_', '')) for b in sorted(glob.glob(os.path.join(work_dir, dir_search)))
            if os.path.isdir(b)]


def log_time(text, starttime):
    """'<text> COMPLETE: hh:mm:ss' style line the reference writes into its -l log file"""
    dt = int(time.time() - starttime)
    return "{0} COMPLETE -- elapsed {1:02d}:{2:02d}:{3:02d}\n".format(text, dt // 3600, (dt % 3600) // 60, dt % 60)
