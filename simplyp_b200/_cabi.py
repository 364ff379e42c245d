"""ctypes binding of ``libsimplyp_b200.so`` (``include/simplyp_b200.h``).

This is the stub a maintainer of the reference would add to call the CUDA path from Python (see
``INTEGRATION.md``).  There is no CPU fallback: if the shared library is missing, or no CUDA device
is visible, the compute entry points raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import packing as pk

_HERE = os.path.dirname(os.path.abspath(__file__))
# SIMPLYP_B200_LIB lets a developer A/B-test another build of the same library (still CUDA-only).
LIB_PATH = os.environ.get("SIMPLYP_B200_LIB") or os.path.join(_HERE, "lib", "libsimplyp_b200.so")

EXPORTS = [
    "simplyp_abi_version", "simplyp_version", "simplyp_last_error", "simplyp_device_count",
    "simplyp_default_options", "simplyp_topology_levels", "simplyp_workspace_bytes",
    "simplyp_run_device", "simplyp_calibrate_device", "simplyp_run_host", "simplyp_calibrate_host",
    "simplyp_release_cache", "simplyp_launch_count", "simplyp_measure_fp64_peak",
    "simplyp_measure_fp64_latency", "simplyp_sum_to_waterbody_device", "simplyp_thornthwaite_pet_device",
    "simplyp_calibrate_gather_device", "simplyp_peer_alloc", "simplyp_peer_free", "simplyp_ipc_export",
    "simplyp_ipc_import", "simplyp_ipc_close",
]
MAX_RANKS = 8


class SimplypDims(C.Structure):
    _fields_ = [("n_members", C.c_int32), ("n_sc", C.c_int32), ("n_days", C.c_int32),
                ("n_sc_param_sets", C.c_int32), ("n_obs_series", C.c_int32), ("reserved", C.c_int32 * 3)]


class SimplypOptions(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("step_len", C.c_double),
                ("max_steps_per_day", C.c_int32), ("dynamic_epc0", C.c_int32),
                ("dynamic_erodibility", C.c_int32), ("run_mode_cal", C.c_int32), ("sc_qr0", C.c_int32),
                ("strict_quirks", C.c_int32), ("threads_per_block", C.c_int32), ("lanes_per_item", C.c_int32),
                ("pilot_days", C.c_int32), ("rank_stats", C.c_int32), ("snow_on_device", C.c_int32), ("reserved", C.c_int32 * 1)]


class SimplypPeerGather(C.Structure):
    _fields_ = [("n_ranks", C.c_int32), ("rank", C.c_int32), ("member_offset", C.c_int64),
                ("n_members_total", C.c_int64), ("step", C.c_int64), ("stats_bufs", C.c_void_p * MAX_RANKS),
                ("flag_bufs", C.c_void_p * MAX_RANKS)]


class SimplypError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (built by ``__graft_entry__.build()``); raise if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SimplypError("%s not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    dp, ip, lp, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_void_p
    lib.simplyp_abi_version.restype = C.c_int
    lib.simplyp_version.restype = C.c_char_p
    lib.simplyp_last_error.restype = C.c_char_p
    lib.simplyp_device_count.restype = C.c_int
    lib.simplyp_default_options.argtypes = [C.POINTER(SimplypOptions)]
    lib.simplyp_default_options.restype = None
    lib.simplyp_topology_levels.argtypes = [C.c_int32, ip, ip, ip]
    lib.simplyp_topology_levels.restype = C.c_int
    lib.simplyp_workspace_bytes.argtypes = [C.POINTER(SimplypDims), C.c_int]
    lib.simplyp_workspace_bytes.restype = C.c_int64
    lib.simplyp_run_device.argtypes = [C.POINTER(SimplypDims), C.POINTER(SimplypOptions), vp, vp, vp, ip, ip,
                                       vp, vp, vp, vp]
    lib.simplyp_run_device.restype = C.c_int
    lib.simplyp_calibrate_device.argtypes = [C.POINTER(SimplypDims), C.POINTER(SimplypOptions), vp, vp, vp, ip, ip,
                                             vp, vp, vp, vp, vp, vp]
    lib.simplyp_calibrate_device.restype = C.c_int
    lib.simplyp_run_host.argtypes = [C.c_int, C.POINTER(SimplypDims), C.POINTER(SimplypOptions), dp, dp, dp, ip, ip,
                                     dp, lp]
    lib.simplyp_run_host.restype = C.c_int
    lib.simplyp_calibrate_host.argtypes = [C.c_int, C.POINTER(SimplypDims), C.POINTER(SimplypOptions), dp, dp, dp,
                                           ip, ip, dp, ip, dp, lp]
    lib.simplyp_calibrate_host.restype = C.c_int
    lib.simplyp_release_cache.restype = None
    lib.simplyp_launch_count.restype = C.c_int64
    lib.simplyp_measure_fp64_peak.argtypes = [C.c_int, C.c_int]
    lib.simplyp_measure_fp64_peak.restype = C.c_double
    lib.simplyp_measure_fp64_latency.argtypes = [C.c_int]
    lib.simplyp_measure_fp64_latency.restype = C.c_double
    lib.simplyp_sum_to_waterbody_device.argtypes = [C.POINTER(SimplypDims), vp, vp, vp, vp, C.c_int32, vp, vp]
    lib.simplyp_sum_to_waterbody_device.restype = C.c_int
    lib.simplyp_thornthwaite_pet_device.argtypes = [C.c_int32, C.c_int32, vp, C.c_int32, vp, vp, C.c_double, vp,
                                                    C.c_int32, vp]
    lib.simplyp_thornthwaite_pet_device.restype = C.c_int
    lib.simplyp_calibrate_gather_device.argtypes = [C.POINTER(SimplypDims), C.POINTER(SimplypOptions), vp, vp, vp, ip,
                                                    ip, vp, vp, C.POINTER(SimplypPeerGather), vp, vp, vp]
    lib.simplyp_calibrate_gather_device.restype = C.c_int
    lib.simplyp_peer_alloc.argtypes = [C.c_int64, C.POINTER(C.c_void_p)]
    lib.simplyp_peer_alloc.restype = C.c_int
    lib.simplyp_peer_free.argtypes = [vp]
    lib.simplyp_peer_free.restype = C.c_int
    lib.simplyp_ipc_export.argtypes = [vp, C.c_char_p]
    lib.simplyp_ipc_export.restype = C.c_int
    lib.simplyp_ipc_import.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.simplyp_ipc_import.restype = C.c_int
    lib.simplyp_ipc_close.argtypes = [vp]
    lib.simplyp_ipc_close.restype = C.c_int
    if lib.simplyp_abi_version() != 4:
        raise SimplypError("ABI version mismatch")
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise SimplypError("simplyp_b200 error %d: %s" % (rc, load().simplyp_last_error().decode()))


def require_device():
    lib = load()
    if lib.simplyp_device_count() <= 0:
        raise SimplypError("no CUDA device visible: simplyp_b200 has no CPU fallback")
    return lib


def default_options(**overrides):
    opt = SimplypOptions()
    load().simplyp_default_options(C.byref(opt))
    for k, v in overrides.items():
        if not hasattr(opt, k):
            raise AttributeError(k)
        setattr(opt, k, v)
    return opt


def make_dims(n_members, n_sc, n_days, n_sc_param_sets=1, n_obs_series=0, n_edges=0):
    d = SimplypDims()
    d.n_members, d.n_sc, d.n_days = int(n_members), int(n_sc), int(n_days)
    d.n_sc_param_sets, d.n_obs_series = int(n_sc_param_sets), int(n_obs_series)
    d.reserved[0] = int(n_edges)
    return d


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _c64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (shape, a.shape))
    return a


def _topology_arrays(parent_offsets, parent_ids):
    po = np.ascontiguousarray(parent_offsets, dtype=np.int32)
    pid = np.ascontiguousarray(parent_ids, dtype=np.int32)
    if pid.size == 0:
        pid = np.zeros(1, dtype=np.int32)
    return po, pid


def topology_levels(parent_offsets, parent_ids):
    po, pid = _topology_arrays(parent_offsets, parent_ids)
    S = len(po) - 1
    lv = np.zeros(max(S, 1), dtype=np.int32)
    n = load().simplyp_topology_levels(S, _iptr(po), _iptr(pid), _iptr(lv))
    if n < 0:
        _check(n)
    return n, lv[:S]


# ------------------------------------------------------------------------------------------ host-buffer calls
def _result_buffer(given, shape, dtype):
    """A caller-supplied result buffer (e.g. pinned host memory) must be C-contiguous with the expected shape/dtype."""
    if given is None:
        return np.zeros(shape, dtype=dtype) if dtype == np.int64 else np.empty(shape, dtype=dtype)
    if tuple(given.shape) != tuple(shape) or given.dtype != dtype or not given.flags["C_CONTIGUOUS"]:
        raise ValueError("result buffer must be a C-contiguous %s array of shape %s" % (np.dtype(dtype).name, tuple(shape)))
    return given


def run_host(forcing, member_params, sc_params, parent_offsets, parent_ids, opt, device=0, want_diag=True, out=None,
             diag=None):
    """numpy in / numpy out through ``simplyp_run_host``.  Returns (out [M][S][D][25], diag [M][S][4]); ``out`` /
    ``diag`` may be given (pinned host memory makes the copies asynchronous DMA)."""
    lib = require_device()
    forcing = _c64(forcing)
    member_params = _c64(member_params)
    sc_params = _c64(sc_params)
    if sc_params.ndim == 2:
        sc_params = sc_params[None]
    M, D = member_params.shape[0], forcing.shape[0]
    Msc, S = sc_params.shape[0], sc_params.shape[1]
    po, pid = _topology_arrays(parent_offsets, parent_ids)
    dims = make_dims(M, S, D, Msc, 0, int(po[-1]))
    out = _result_buffer(out, (M, S, D, pk.NOUT), np.float64)
    diag = _result_buffer(diag, (M, S, pk.NDIAG), np.int64)
    rc = lib.simplyp_run_host(device, C.byref(dims), C.byref(opt), _dptr(forcing), _dptr(member_params),
                              _dptr(sc_params), _iptr(po), _iptr(pid), _dptr(out),
                              diag.ctypes.data_as(C.POINTER(C.c_int64)) if want_diag else None)
    _check(rc)
    return out, diag


def calibrate_host(forcing, member_params, sc_params, parent_offsets, parent_ids, obs, obs_desc, opt,
                   device=0, want_diag=True, stats=None, diag=None):
    """numpy in / numpy out through ``simplyp_calibrate_host``.  Returns (stats [M][V][10], diag [M][S][4])."""
    lib = require_device()
    forcing = _c64(forcing)
    member_params = _c64(member_params)
    sc_params = _c64(sc_params)
    if sc_params.ndim == 2:
        sc_params = sc_params[None]
    obs = _c64(obs)
    obs_desc = np.ascontiguousarray(obs_desc, dtype=np.int32)
    M, D = member_params.shape[0], forcing.shape[0]
    Msc, S = sc_params.shape[0], sc_params.shape[1]
    V = obs.shape[0]
    po, pid = _topology_arrays(parent_offsets, parent_ids)
    dims = make_dims(M, S, D, Msc, V, int(po[-1]))
    stats = _result_buffer(stats, (M, V, pk.NSTAT), np.float64)
    diag = _result_buffer(diag, (M, S, pk.NDIAG), np.int64)
    rc = lib.simplyp_calibrate_host(device, C.byref(dims), C.byref(opt), _dptr(forcing), _dptr(member_params),
                                    _dptr(sc_params), _iptr(po), _iptr(pid), _dptr(obs), _iptr(obs_desc),
                                    _dptr(stats), diag.ctypes.data_as(C.POINTER(C.c_int64)) if want_diag else None)
    _check(rc)
    return stats, diag


# ------------------------------------------------------------------------------------------ device-pointer calls
def run_device(dims, opt, forcing_ptr, member_ptr, sc_ptr, parent_offsets, parent_ids, out_ptr, diag_ptr,
               ws_ptr, stream_ptr):
    """Raw device-pointer call (torch ``.data_ptr()`` integers); enqueues on ``stream_ptr``."""
    lib = require_device()
    po, pid = _topology_arrays(parent_offsets, parent_ids)
    _check(lib.simplyp_run_device(C.byref(dims), C.byref(opt), forcing_ptr, member_ptr, sc_ptr, _iptr(po), _iptr(pid),
                                  out_ptr, diag_ptr, ws_ptr, stream_ptr))


def calibrate_device(dims, opt, forcing_ptr, member_ptr, sc_ptr, parent_offsets, parent_ids, obs_ptr, desc_ptr,
                     stats_ptr, diag_ptr, ws_ptr, stream_ptr):
    lib = require_device()
    po, pid = _topology_arrays(parent_offsets, parent_ids)
    _check(lib.simplyp_calibrate_device(C.byref(dims), C.byref(opt), forcing_ptr, member_ptr, sc_ptr, _iptr(po),
                                        _iptr(pid), obs_ptr, desc_ptr, stats_ptr, diag_ptr, ws_ptr, stream_ptr))


def calibrate_gather_device(dims, opt, forcing_ptr, member_ptr, sc_ptr, parent_offsets, parent_ids, obs_ptr, desc_ptr,
                            gather, diag_ptr, ws_ptr, stream_ptr):
    """Calibration fused with the all-gather of the statistics over peer memory (``gather``: SimplypPeerGather)."""
    lib = require_device()
    po, pid = _topology_arrays(parent_offsets, parent_ids)
    _check(lib.simplyp_calibrate_gather_device(C.byref(dims), C.byref(opt), forcing_ptr, member_ptr, sc_ptr, _iptr(po),
                                               _iptr(pid), obs_ptr, desc_ptr, C.byref(gather), diag_ptr, ws_ptr,
                                               stream_ptr))


def peer_alloc(n_bytes):
    p = C.c_void_p()
    _check(require_device().simplyp_peer_alloc(int(n_bytes), C.byref(p)))
    return int(p.value)


def peer_free(ptr):
    _check(load().simplyp_peer_free(C.c_void_p(ptr)))


def ipc_export(ptr):
    buf = C.create_string_buffer(64)
    _check(load().simplyp_ipc_export(C.c_void_p(ptr), buf))
    return bytes(buf.raw)


def ipc_import(handle):
    p = C.c_void_p()
    _check(load().simplyp_ipc_import(C.create_string_buffer(bytes(handle), 64), C.byref(p)))
    return int(p.value)


def ipc_close(ptr):
    _check(load().simplyp_ipc_close(C.c_void_p(ptr)))


def sum_to_waterbody_device(dims, out_ptr, sc_ptr, member_ptr, reaches_ptr, n_reaches, wb_ptr, stream_ptr):
    lib = require_device()
    _check(lib.simplyp_sum_to_waterbody_device(C.byref(dims), out_ptr, sc_ptr, member_ptr, reaches_ptr,
                                               int(n_reaches), wb_ptr, stream_ptr))


def thornthwaite_pet_device(n_days, n_months, t_air_ptr, t_stride, month_start_ptr, year_is_leap_ptr, latitude_deg,
                            pet_ptr, pet_stride, stream_ptr):
    """Reference ``daily_PET`` (``inputs.py:232-312``) on device pointers; asynchronous on ``stream_ptr``."""
    _check(load().simplyp_thornthwaite_pet_device(int(n_days), int(n_months), C.c_void_p(t_air_ptr), int(t_stride),
                                                  C.c_void_p(month_start_ptr), C.c_void_p(year_is_leap_ptr),
                                                  float(latitude_deg), C.c_void_p(pet_ptr), int(pet_stride),
                                                  C.c_void_p(stream_ptr)))


def workspace_bytes(dims, calibrate, rank_stats=False):
    n = load().simplyp_workspace_bytes(C.byref(dims), (1 if calibrate else 0) | (2 if (calibrate and rank_stats) else 0))
    if n < 0:
        _check(int(n))
    return int(n)


def launch_count():
    return int(load().simplyp_launch_count())


def measure_fp64_latency(device=0):
    require_device()
    return float(load().simplyp_measure_fp64_latency(device))


def measure_fp64_peak(device=0, repeats=3):
    require_device()
    return float(load().simplyp_measure_fp64_peak(device, repeats))
