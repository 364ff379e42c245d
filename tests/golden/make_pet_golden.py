"""Generates tests/golden/ref_pet.json from the UNMODIFIED reference's Thornthwaite helpers
(``/root/reference/Current_Release/v0-2A/simplyP/inputs.py:315-508``), imported through oracle/reference_live.py.
The reference's ``daily_PET`` wrapper itself (``inputs.py:232-312``) does not run under pandas 3 (``resample('M')``,
``pd.DatetimeIndex(freq=...)``), which is recorded in the fixture; its arithmetic is all in the helpers pinned here.
Run in the build container only:  python tests/golden/make_pet_golden.py
"""
import json
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import reference_live as rl  # noqa: E402

sp = rl.load()
ri = sp.inputs
z = np.load(os.path.join(HERE, "..", "..", "simplyp_b200", "data", "tarland", "tarland_met.npz"), allow_pickle=True)
idx = pd.date_range("1981-01-01", periods=len(z["T_air"]), freq="D")
t_air = pd.Series(z["T_air"].astype(float), index=idx)
out = {"latitude_deg": 57.1, "years": {}}
lat = ri.deg2rad(57.1)
out["dlh_normal"] = [float(v) for v in ri.monthly_mean_daylight_hours(lat, year=1983)]
out["dlh_leap"] = [float(v) for v in ri.monthly_mean_daylight_hours(lat, year=1984)]
import calendar
for year in (1981, 1984, 2003, 2004, 2010):
    tm = t_air[str(year)].groupby(t_air[str(year)].index.month).mean().to_numpy()
    dlh = out["dlh_leap"] if calendar.isleap(year) else out["dlh_normal"]
    out["years"][str(year)] = {"monthly_t": [float(v) for v in tm],
                               "pet_mm_month": [float(v) for v in ri.annual_thornthwaite(tm, dlh, year=year)]}
try:
    met = pd.DataFrame({"T_air": t_air["2003":"2004"]})
    ref = ri.daily_PET(57.1, met)
    out["daily_PET_2003_2004"] = [float(v) for v in ref["PET"].to_numpy()]
    out["daily_PET_status"] = "ran"
except Exception as e:  # pandas-3 incompatibilities of the wrapper
    out["daily_PET_status"] = "reference wrapper does not run here: %s: %s" % (type(e).__name__, str(e)[:120])
with open(os.path.join(HERE, "ref_pet.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out["daily_PET_status"])
