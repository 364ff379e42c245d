"""Unit conversions and small helpers, same names and meaning as the reference's
``simplyP/helper_functions.py`` (UC_Q :6, UC_Qinv :19, UC_C :32, UC_Cinv :46,
UC_V :59, lin_interp :77).  All of them broadcast over numpy arrays / pandas
objects exactly as the reference's do because they are plain arithmetic.
"""

_SECONDS_PER_DAY = 86400


def UC_Q(Q_mmd, A_catch):
    """Discharge mm/day -> m3/day for a catchment of ``A_catch`` km2."""
    return Q_mmd * 1000 * A_catch


def UC_Qinv(Q_m3s, A_catch):
    """Discharge m3/s -> mm/day for a catchment of ``A_catch`` km2."""
    return Q_m3s * _SECONDS_PER_DAY / (1000 * A_catch)


def UC_C(C_kgmm, A_catch):
    """Concentration kg/mm -> mg/l (1 mm over 1 km2 is 1e6 l; 1 kg is 1e6 mg)."""
    return C_kgmm / A_catch


def UC_Cinv(C_mgl, A_catch):
    """Concentration mg/l -> kg/mm."""
    return C_mgl * A_catch


def UC_V(V_mm, A_catch, outUnits):
    """Depth in mm over the catchment -> volume in 'm3' or 'l'."""
    factor = {"m3": 10 ** 3, "l": 10 ** 6}[outUnits]
    return V_mm * factor * A_catch


def lin_interp(x, x0, x1, y0, y1):
    """Value at ``x`` on the straight line through (x0, y0) and (x1, y1)."""
    return y0 + (y1 - y0) * (x - x0) / (x1 - x0)
