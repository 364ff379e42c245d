"""Device-resident execution of the integrator: torch owns memory and streams, the C-ABI does the work.

``Engine`` keeps forcing / parameters / observations resident in HBM as torch tensors and calls the
``*_device`` entry points of ``libsimplyp_b200.so`` with raw pointers on torch's current stream.
PyTorch is plumbing only (allocation, streams, ``torch.distributed``); every arithmetic step of the
path runs in the hand-written kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from . import packing as pk


def _torch():
    import torch
    return torch


class Engine:
    """One engine per process / GPU."""

    def __init__(self, device=None):
        torch = _torch()
        _cabi.require_device()
        if not torch.cuda.is_available():
            raise _cabi.SimplypError("torch sees no CUDA device: simplyp_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self._ws = None

    # ------------------------------------------------------------------ helpers
    def to_device(self, a, dtype=None):
        torch = _torch()
        if isinstance(a, torch.Tensor):
            return a.to(self.device).contiguous()
        a = np.ascontiguousarray(a, dtype=dtype)
        return torch.from_numpy(a).to(self.device)

    def _workspace(self, nbytes):
        torch = _torch()
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
        return self._ws

    def free_bytes(self):
        torch = _torch()
        free, _total = torch.cuda.mem_get_info(self.device)
        return int(free)

    @staticmethod
    def _shapes(forcing, member_params, sc_params):
        D = forcing.shape[0]
        M = member_params.shape[0]
        if sc_params.dim() == 2:
            sc_params = sc_params.unsqueeze(0)
        Msc, S = sc_params.shape[0], sc_params.shape[1]
        assert forcing.shape[1] == pk.NF and member_params.shape[1] == pk.NP_MEMBER and sc_params.shape[2] == pk.NP_SC
        return D, M, Msc, S, sc_params

    # ------------------------------------------------------------------ full output
    def run(self, forcing, member_params, sc_params, parent_offsets, parent_ids, opt, out=None, diag=None):
        """All tensors on this device, fp64.  Returns (out [M][S][D][25], diag [M][S][4] int64); asynchronous."""
        torch = _torch()
        D, M, Msc, S, sc_params = self._shapes(forcing, member_params, sc_params)
        n_edges = int(parent_offsets[-1])
        dims = _cabi.make_dims(M, S, D, Msc, 0, n_edges)
        if out is None:
            out = torch.empty((M, S, D, pk.NOUT), dtype=torch.float64, device=self.device)
        if diag is None:
            diag = torch.zeros((M, S, pk.NDIAG), dtype=torch.int64, device=self.device)
        ws = self._workspace(_cabi.workspace_bytes(dims, False))
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            _cabi.run_device(dims, opt, forcing.data_ptr(), member_params.data_ptr(), sc_params.data_ptr(),
                             parent_offsets, parent_ids, out.data_ptr(), diag.data_ptr(), ws.data_ptr(), stream)
        return out, diag

    # ------------------------------------------------------------------ fused statistics
    def calibrate(self, forcing, member_params, sc_params, parent_offsets, parent_ids, obs, obs_desc, opt,
                  stats=None, diag=None, max_workspace_bytes=None, peer_gather=None):
        """Returns (stats [M][V][10], diag [M][S][4]); asynchronous on the current stream.

        With ``peer_gather`` (an :class:`ensemble.PeerGather`; the members given here are this rank's shard) the
        statistics of ALL ranks come back, [M_total][V][10]: the kernel stores them into every rank's buffer and a flag
        exchange ends the call on the stream — no collective follows.

        The workspace grows with the members: M*S*D*32 bytes of flux exchange for networks (S > 1) and M*V*D*8
        bytes of simulated values with ``opt.rank_stats``; members are processed in chunks that keep it under
        ``max_workspace_bytes`` (default: half of the free HBM).
        """
        torch = _torch()
        D, M, Msc, S, sc_params = self._shapes(forcing, member_params, sc_params)
        V = obs.shape[0]
        n_edges = int(parent_offsets[-1])
        if stats is None and peer_gather is None:
            stats = torch.empty((M, V, pk.NSTAT), dtype=torch.float64, device=self.device)
        if diag is None:
            diag = torch.zeros((M, S, pk.NDIAG), dtype=torch.int64, device=self.device)
        chunk = M
        per_member = (32 * S * D if S > 1 else 0) + (8 * V * D if opt.rank_stats else 0)
        if per_member > 0:
            budget = max_workspace_bytes if max_workspace_bytes is not None else self.free_bytes() // 2
            chunk = max(1, min(M, int(budget // per_member)))
        if peer_gather is not None:
            if chunk < M or opt.rank_stats or M != peer_gather.hi - peer_gather.lo:
                raise _cabi.SimplypError("peer_gather: one launch of this rank's whole shard, without rank statistics")
            with torch.cuda.device(self.device):
                stream = torch.cuda.current_stream().cuda_stream
                dims = _cabi.make_dims(M, S, D, 1 if Msc == 1 else M, V, n_edges)
                ws = self._workspace(_cabi.workspace_bytes(dims, True, False))
                desc, gathered = peer_gather.descriptor()
                _cabi.calibrate_gather_device(dims, opt, forcing.data_ptr(), member_params.data_ptr(),
                                              sc_params.data_ptr(), parent_offsets, parent_ids, obs.data_ptr(),
                                              obs_desc.data_ptr(), desc, diag.data_ptr(), ws.data_ptr(), stream)
            return gathered, diag
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            for m0 in range(0, M, chunk):
                m1 = min(M, m0 + chunk)
                dims = _cabi.make_dims(m1 - m0, S, D, 1 if Msc == 1 else (m1 - m0), V, n_edges)
                ws = self._workspace(_cabi.workspace_bytes(dims, True, bool(opt.rank_stats)))
                scp = sc_params if Msc == 1 else sc_params[m0:m1]
                _cabi.calibrate_device(dims, opt, forcing.data_ptr(), member_params[m0:m1].data_ptr(),
                                       scp.data_ptr(), parent_offsets, parent_ids, obs.data_ptr(),
                                       obs_desc.data_ptr(), stats[m0:m1].data_ptr(), diag[m0:m1].data_ptr(),
                                       ws.data_ptr(), stream)
        return stats, diag

    # ------------------------------------------------------------------ receiving water body
    WATERBODY_COLUMNS = ["Q_cumecs", "Msus_kg/day", "TDP_kg/day", "PP_kg/day", "SS_mgl", "TDP_mgl", "PP_mgl",
                         "TP_mgl", "TP_kg/day", "SRP_mgl", "SRP_kg/day"]

    def sum_to_waterbody(self, out, member_params, sc_params, reaches):
        """Reference ``sum_to_waterbody`` (``model.py:851-900``) for every member of a full-output run, on the
        device: ``out`` [M][S][D][25] from :meth:`run`, ``reaches`` = run-order indices of the reaches flagged
        ``In_final_flux? == 1``.  Returns [M][D][11] (columns ``WATERBODY_COLUMNS``); asynchronous."""
        torch = _torch()
        M, S, D = out.shape[0], out.shape[1], out.shape[2]
        if sc_params.dim() == 2:
            sc_params = sc_params.unsqueeze(0)
        dims = _cabi.make_dims(M, S, D, sc_params.shape[0], 0, 0)
        r = torch.as_tensor(np.asarray(reaches, dtype=np.int32), device=self.device)
        wb = torch.empty((M, D, len(self.WATERBODY_COLUMNS)), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            _cabi.sum_to_waterbody_device(dims, out.data_ptr(), sc_params.data_ptr(), member_params.data_ptr(),
                                          r.data_ptr(), int(r.numel()), wb.data_ptr(), stream)
        return wb

    # ------------------------------------------------------------------ Thornthwaite PET
    def thornthwaite_pet(self, t_air, month_start, year_is_leap, latitude_deg, out=None, out_stride=1):
        """Reference ``daily_PET`` (``inputs.py:232-312``) on the device: ``t_air`` [D] (device or host),
        ``month_start`` [n_months + 1] day index at which each calendar month of the record starts,
        ``year_is_leap`` [n_months / 12]; returns PET [D] in mm/day (or writes ``out`` with ``out_stride``,
        e.g. column 1 of a forcing matrix: ``out=forcing[:, 1], out_stride=4``); asynchronous."""
        torch = _torch()
        t = self.to_device(t_air, np.float64)
        ms = torch.as_tensor(np.asarray(month_start, dtype=np.int32), device=self.device)
        lp = torch.as_tensor(np.asarray(year_is_leap, dtype=np.int32), device=self.device)
        D, NM = int(t.numel()), int(ms.numel()) - 1
        if out is None:
            out = torch.empty(D, dtype=torch.float64, device=self.device)
            out_stride = 1
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            _cabi.thornthwaite_pet_device(D, NM, t.data_ptr(), 1, ms.data_ptr(), lp.data_ptr(), latitude_deg,
                                          out.data_ptr(), out_stride, stream)
        return out
