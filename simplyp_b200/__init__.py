"""simplyp_b200 — B200-native SimplyP daily mass-balance integration.

Flat re-export of the same public names as the reference package
(reference ``simplyP/__init__.py:1-31``), minus the matplotlib plotting
functions which are out of scope.
"""
from .helper_functions import UC_Q, UC_Qinv, UC_C, UC_Cinv, UC_V, lin_interp
from .inputs import read_input_data, snow_hydrol_inputs, daily_PET
from .model import run_simply_p, derived_P_species, sum_to_waterbody
from .stats import goodness_of_fit_stats

__version__ = "0.1.0"
