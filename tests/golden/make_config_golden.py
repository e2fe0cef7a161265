"""Run in the build container only (reads /root/reference): the [parameters-depression-handling] values of the
reference's config_files/config_local.ini, as the reference's own __main__ converts them
(models/depression_handling.py:1444-1465), for the CPU test that pins DepHandleCore's defaults."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pydrodem_b200 import config

from configparser import ConfigParser, ExtendedInterpolation
# (the shipped ini repeats `output_dtm_dir` in [paths]; a strict parser -- the reference's own -- rejects the file, so
#  it is read non-strictly here: the depression-handling section is unaffected)
cfg = ConfigParser(allow_no_value=True, interpolation=ExtendedInterpolation(), strict=False)
cfg.read("/root/reference/config_files/config_local.ini")
opts = config.dephandle_options(cfg)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config_local_dephandle.json")
with open(out, "w") as f:
    json.dump(opts, f, indent=1, sort_keys=True)
print(opts)
