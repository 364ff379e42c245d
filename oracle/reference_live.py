"""TEST INFRASTRUCTURE — live import of the UNMODIFIED reference (this container only).

Imports ``/root/reference/Current_Release/v0-2A/simplyP`` as it is, under three
process-local compatibility shims, so that the reference's own ``run_simply_p``
(``model.py:193-827``), ``ode_f`` (``:58-187``), ``f_x`` (``:23-37``),
``discretized_soilP`` (``:39-56``) and ``goodness_of_fit_stats``
(``visualise_results.py:387-474``) run against SciPy's real LSODA.  It is used

* by ``tests/golden/make_golden.py`` to generate the committed fixtures, and
* by the ``not gpu`` tests (when ``/root/reference`` exists) to pin
  ``oracle/simplyp_oracle.py`` against the reference itself.

Nothing here is shipped or measured as product; ``/root/reference`` does not
exist on the GPU box, so nothing under ``-m gpu``, ``smoke()`` or ``bench.py``
may import this module.

Shims (none edits a reference file):
1. stub ``matplotlib``/``seaborn`` modules (pulled in by ``simplyP/__init__.py:25``),
2. ``np.NaN`` (removed in numpy 2; used at ``model.py:311-313,549``),
3. an ``.ix`` indexer on DataFrame/Series (removed in pandas 1.0; used at
   ``model.py:497-498,524-528`` and ``visualise_results.py:445,447``):
   an integer key on a non-integer axis is positional, anything else is by label.

``set_tolerances(rtol, atol)`` rebinds the module-level name ``simplyP.model.odeint``
(what ``run_simply_p`` resolves at ``model.py:640``) to a wrapper that overrides the
hard-coded ``rtol=0.01`` — that is how the tight-tolerance oracle runs are made.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np
import pandas as pd

REFERENCE_ROOT = "/root/reference"
REFERENCE_PKG_DIR = os.path.join(REFERENCE_ROOT, "Current_Release", "v0-2A")

_state = {"module": None, "nfe": 0, "ncalls": 0}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_PKG_DIR, "simplyP"))


# --------------------------------------------------------------------------- shims
class _Anything(types.ModuleType):
    """Module stub whose every attribute is a no-op callable/stub."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        stub = _Anything(self.__name__ + "." + name)
        setattr(self, name, stub)
        return stub

    def __call__(self, *a, **k):
        return None


def _install_plot_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].ticker = sys.modules["matplotlib.ticker"]


class _Ix:
    def __init__(self, obj):
        self._obj = obj

    def _split(self, key):
        obj = self._obj
        if isinstance(obj, pd.Series):
            return (key,)
        if isinstance(key, tuple):
            return key
        return (key, slice(None))

    @staticmethod
    def _positional(axis_index, k):
        return isinstance(k, (int, np.integer)) and not pd.api.types.is_integer_dtype(axis_index)

    def __getitem__(self, key):
        obj = self._obj
        if isinstance(obj, pd.Series):
            if self._positional(obj.index, key):
                return obj.iloc[key]
            return obj.loc[key]
        r, c = self._split(key)
        rpos = self._positional(obj.index, r)
        cpos = self._positional(obj.columns, c)
        if rpos and cpos:
            return obj.iloc[r, c]
        if rpos:
            c_i = obj.columns.get_loc(c) if not isinstance(c, slice) else c
            return obj.iloc[r, c_i]
        if cpos:
            return obj.loc[r].iloc[c]
        return obj.loc[r, c]

    def __setitem__(self, key, value):
        obj = self._obj
        if isinstance(obj, pd.Series):
            if self._positional(obj.index, key):
                obj.iloc[key] = value
            else:
                obj.loc[key] = value
            return
        r, c = self._split(key)
        if self._positional(obj.index, r):
            obj.iloc[r, obj.columns.get_loc(c)] = value
        else:
            obj.loc[r, c] = value


def _install_numpy_pandas_shims():
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    if not hasattr(pd.DataFrame, "ix"):
        pd.DataFrame.ix = property(lambda self: _Ix(self))
    if not hasattr(pd.Series, "ix"):
        pd.Series.ix = property(lambda self: _Ix(self))


# --------------------------------------------------------------------------- import
def load():
    """Import and return the unmodified reference package ``simplyP``."""
    if _state["module"] is not None:
        return _state["module"]
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_PKG_DIR)
    _install_plot_stubs()
    _install_numpy_pandas_shims()
    if REFERENCE_PKG_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_PKG_DIR)
    with contextlib.redirect_stdout(io.StringIO()):
        import simplyP  # noqa: F401  (the reference)
    _state["module"] = simplyP
    _state["orig_odeint"] = simplyP.model.odeint
    return simplyP


def set_tolerances(rtol=None, atol=None, mxstep=50000):
    """Force (rtol, atol) on the reference's odeint call; ``rtol=None`` restores the reference's own."""
    sp = load()
    import scipy.integrate

    if rtol is None:
        def _counting(f, y0, t, args=(), full_output=0, rtol=None, mxstep=5000):
            y, info = scipy.integrate.odeint(f, y0, t, args=args, full_output=1, rtol=rtol, mxstep=mxstep)
            _state["nfe"] += int(info["nfe"][-1])
            _state["ncalls"] += 1
            return y, info
        sp.model.odeint = _counting
        return

    def _forced(f, y0, t, args=(), full_output=0, rtol_=rtol, atol_=atol, **_ignored):
        y, info = scipy.integrate.odeint(f, y0, t, args=args, full_output=1,
                                         rtol=rtol_, atol=atol_, mxstep=mxstep)
        _state["nfe"] += int(info["nfe"][-1])
        _state["ncalls"] += 1
        return y, info

    sp.model.odeint = _forced


def solver_counters(reset=False):
    out = {"nfe": _state["nfe"], "ncalls": _state["ncalls"]}
    if reset:
        _state["nfe"] = 0
        _state["ncalls"] = 0
    return out


def run_simply_p(met_df, p_struc, p_SU, p_LU, p_SC, p, dynamic_options, step_len=1.0,
                 rtol=None, atol=None, quiet=True):
    """Run the reference's own ``run_simply_p`` on deep copies of the inputs.

    ``rtol=None`` keeps the reference's hard-coded tolerance (rtol=0.01, default atol).
    """
    sp = load()
    set_tolerances(rtol, atol)
    args = (met_df.copy(deep=True), p_struc.copy(deep=True), p_SU.copy(deep=True), p_LU.copy(deep=True),
            p_SC.copy(deep=True).astype(object), p.copy(deep=True), dynamic_options.copy(deep=True))
    try:
        if quiet:
            with contextlib.redirect_stdout(io.StringIO()):
                res = sp.run_simply_p(*args, step_len=step_len)
        else:
            res = sp.run_simply_p(*args, step_len=step_len)
    finally:
        set_tolerances(None)
    return res


def goodness_of_fit_stats(p_SU, df_R_dict, obs_dict):
    sp = load()
    with contextlib.redirect_stdout(io.StringIO()):
        return sp.goodness_of_fit_stats(p_SU, df_R_dict, obs_dict)


# --------------------------------------------------------------------------- daily_PET wrapper (inputs.py:232-312)
class _PandasProxy(types.ModuleType):
    """Stands in for the name ``pd`` inside the reference's ``inputs`` module: everything is pandas, except that
    ``DatetimeIndex(freq=, start=, end=)`` — a constructor form pandas dropped in 1.0 — is answered by
    ``pd.date_range`` (what that form always did)."""

    def __getattr__(self, name):
        return getattr(pd, name)

    @staticmethod
    def DatetimeIndex(*args, **kw):
        if "start" in kw or "end" in kw:
            kw.pop("dayfirst", None)
            return pd.date_range(start=kw.pop("start", None), end=kw.pop("end", None), freq=kw.pop("freq", None), **kw)
        return pd.DatetimeIndex(*args, **kw)


@contextlib.contextmanager
def _month_end_alias():
    """``Series.resample('M')`` (month end; renamed 'ME' in pandas 2.2 and refused in 3) for the duration of a call."""
    orig = pd.Series.resample

    def resample(self, rule, *a, **k):
        return orig(self, "ME" if rule == "M" else rule, *a, **k)

    pd.Series.resample = resample
    try:
        yield
    finally:
        pd.Series.resample = orig


def daily_PET(latitude, met_df):
    """The reference's own ``daily_PET`` wrapper (``inputs.py:232-312``), unmodified, on a deep copy of ``met_df``.
    Two more process-local shims let it run under pandas 3: the month-end alias and the ``DatetimeIndex`` form."""
    sp = load()
    mod = sp.inputs
    saved = mod.pd
    mod.pd = _PandasProxy("pandas_proxy")
    try:
        with _month_end_alias(), contextlib.redirect_stdout(io.StringIO()):
            return mod.daily_PET(latitude, met_df.copy(deep=True))
    finally:
        mod.pd = saved
