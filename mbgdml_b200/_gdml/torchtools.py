"""``GDMLTorchPredict`` on the device kernels (reference mbgdml/_gdml/torchtools.py:401-1063).

The reference class is a ``torch.nn.Module`` that evaluates an sGDML model with einsum-based PyTorch code; here the
same interface -- ``forward(Rs, return_E=True)`` with ``Rs`` a (M, N, 3) tensor, returning ``(Es, Fs)`` tensors on
the device of ``Rs`` -- runs through the Matern contraction kernels of the prediction path.  CUDA tensors are
consumed and produced in place (no host round trip); the call is ordered on torch's current stream.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from .. import _engine, _native
from ..models import gdmlModel
from ..utils import gen_combs


class GDMLTorchPredict(nn.Module):
    """PyTorch front-end of ``GDMLPredict``; holds no trainable parameters (torchtools.py:401-405)."""

    def __init__(self, model, lat_and_inv=None, batch_size=None, n_perm_batches=1, max_memory=None,
                 max_processes=None, log_level=None):  # pylint: disable=unused-argument
        super().__init__()
        model = dict(model)
        if lat_and_inv is not None:  # the kernels clamp pair differences to this lattice (_gdml/desc.py:42-76)
            lat = lat_and_inv[0]
            model["lattice"] = lat.detach().cpu().numpy() if hasattr(lat, "detach") else np.asarray(lat)
        model.setdefault("r_unit", np.array("Angstrom"))
        self.dim_d, self.n_train = model["R_desc"].shape[:2]
        self.n_perms, self.n_atoms = model["perms"].shape
        self.dim_i = 3 * self.n_atoms
        self._model_dict = model
        self.R_d_desc = None  # training mode: Jacobians of the training geometries (torchtools.py:702-732)
        self._model = gdmlModel(model, comp_ids=np.array(["mol"]))
        self._std = float(self._model.stdev)
        self._c = float(self._model.integ_c)
        self._Z = np.asarray(self._model.Z)
        self._eids = np.zeros(self.n_atoms, dtype=np.int64)

    def get_n_perm_batches(self):  # memory knobs of the reference: nothing to tune here (torchtools.py:584-605)
        return 1

    def set_n_perm_batches(self, n_perm_batches):  # pylint: disable=unused-argument
        return None

    @staticmethod
    def _host(x):
        return np.asarray(x.detach().cpu().numpy() if hasattr(x, "detach") else x, dtype=np.float64)

    def set_R_d_desc(self, R_d_desc):  # pylint: disable=invalid-name
        """Descriptor Jacobians (n_train, D, 3) of the training geometries (torchtools.py:702-732): with them the
        training points can be evaluated by index and ``set_alphas`` can reconfigure the model."""
        self.R_d_desc = self._host(R_d_desc)

    def set_alphas(self, alphas, alphas_E=None):  # pylint: disable=invalid-name
        """New regression parameters (torchtools.py:734-849): ``R_d_desc_alpha = d_desc_dot_vec(R_d_desc, alphas)``."""
        assert self.R_d_desc is not None
        a = self._host(alphas).reshape(-1, self.n_atoms, 3)
        i, j = np.tril_indices(self.n_atoms, k=-1)
        model = dict(self._model_dict)
        model["R_d_desc_alpha"] = np.einsum("tkc,tkc->tk", self.R_d_desc, a[:, j, :] - a[:, i, :])
        if alphas_E is not None:
            model["alphas_E"] = self._host(alphas_E)
        self._model_dict = model
        self._model = gdmlModel(model, comp_ids=np.array(["mol"]))

    def forward(self, Rs_or_train_idxs, return_E=True):
        """(M, N, 3) coordinates -> ``(Es (M,), Fs (M, N, 3))``, or ``(Fs,)`` when ``return_E`` is False
        (torchtools.py:999-1063).  A 1-D tensor selects training points by index (torchtools.py:861-877; needs
        ``set_R_d_desc``): their geometries are rebuilt from the model's descriptors and the Jacobians."""
        Rs = Rs_or_train_idxs
        if Rs.dim() == 1:
            if self.R_d_desc is None:
                raise ValueError("set_R_d_desc() must be called before training points can be evaluated by index")
            from ..gdml_predict import geometries_from_desc

            idx = Rs.detach().cpu().numpy().astype(np.int64)
            R_desc = np.asarray(self._model_dict["R_desc"], dtype=np.float64).T[idx]
            Rs = torch.from_numpy(geometries_from_desc(R_desc, self.R_d_desc[idx])).to(Rs.device)
        if Rs.dim() != 3 or Rs.shape[1] != self.n_atoms or Rs.shape[2] != 3:
            raise ValueError(f"Rs must be (M, {self.n_atoms}, 3)")
        M = int(Rs.shape[0])
        src = [gen_combs([np.array([0])])]
        if Rs.is_cuda:
            dev = Rs.device
            R = Rs.detach().to(torch.float64).contiguous()
            ctx = _native.get_context(dev.index if dev.index is not None else torch.cuda.current_device())
            Es = torch.empty(M, dtype=torch.float64, device=dev)
            Fs = torch.empty((M, self.n_atoms, 3), dtype=torch.float64, device=dev)
            if M:
                ts = torch.cuda.current_stream(dev)
                if ts.cuda_stream == 0:
                    # legacy default stream: the library's own (non-blocking) stream is not ordered against it
                    ts.synchronize()
                    ctx.set_stream(0)
                else:
                    ctx.set_stream(ts.cuda_stream)  # stream-ordered with the producer of Rs and the consumer of (Es, Fs)
                try:
                    _engine.predict_total(ctx, self._Z, R, self._eids, [self._model], src, [[]], None,
                                          device_out=(Es, Fs, None))
                    if ts.cuda_stream == 0:
                        ctx.synchronize()
                finally:
                    ctx.set_stream(0)
        else:
            ctx = _native.get_context()
            E, F, _, _ = _engine.predict_total(ctx, self._Z, Rs.detach().to(torch.float64).numpy(), self._eids,
                                               [self._model], src, [[]], None)
            Es, Fs = torch.from_numpy(E), torch.from_numpy(F)
        return (Es, Fs) if return_E else (Fs,)


class GDMLTorchAssemble(nn.Module):
    """Kernel-matrix assembly with the reference's torch front-end (torchtools.py:72-398): ``J`` lists the column
    blocks of the final matrix -- training-point indices ``j`` (``j >= n_train``: the energy-constraint column
    ``j % n_train``), or tuples ``(first column in out, j, partials kept)`` for "fancy" column subsets, as
    ``GDMLTrain._assemble_kernel_mat`` builds them (train.py:1190-1245) -- and ``forward(J_indx)`` fills the columns
    of ``out`` (NumPy array or tensor, rows x columns of the final matrix) that the selected entries own.  The blocks
    come from the device kernels behind ``assemble_kernel_mat`` (csrc/kernel_mat.cuh): one launch per ``forward``
    instead of one einsum chain per column block."""

    def __init__(self, J, tril_perms_lin, sig, use_E_cstr, R_desc_torch, R_d_desc_torch, out, callback=None):
        super().__init__()
        self.callback = callback
        self.n_train, self.dim_d = R_d_desc_torch.shape[:2]
        self.n_atoms = int((1 + np.sqrt(8 * self.dim_d + 1)) / 2)
        self.dim_i = 3 * self.n_atoms
        self.sig = float(sig)
        self.tril_perms_lin = np.asarray(tril_perms_lin)
        self.n_perms = len(self.tril_perms_lin) // self.dim_d
        self.use_E_cstr = bool(use_E_cstr)
        self._R_desc = self._host(R_desc_torch)
        self._R_d_desc = self._host(R_d_desc_torch)
        self.J = list(J)
        self.out = out

    @staticmethod
    def _host(x):
        return np.ascontiguousarray(x.detach().cpu().numpy() if hasattr(x, "detach") else x, dtype=np.float64)

    def _columns(self, entry):
        """(columns of ``out``, columns of the full matrix) of one entry of ``J`` (torchtools.py:119-153)."""
        n_ff = self.n_train * self.dim_i
        if isinstance(entry, tuple):
            K_j, j, keep = entry
            j = int(j)
            if j < self.n_train:
                keep = np.arange(self.dim_i)[keep] if isinstance(keep, slice) else np.asarray(keep, dtype=np.int64)
                return int(K_j) + np.arange(len(keep)), j * self.dim_i + keep
            return np.array([int(K_j)]), np.array([n_ff + j % self.n_train])
        j = int(entry)
        if j < self.n_train:
            cols = j * self.dim_i + np.arange(self.dim_i)
            return cols, cols
        col = np.array([n_ff + j % self.n_train])
        return col, col

    def forward(self, J_indx):
        from .train import assemble_kernel_mat

        picked = [self._columns(self.J[int(i)]) for i in J_indx]
        if not picked:
            return
        dst = np.concatenate([p[0] for p in picked])
        src = np.concatenate([p[1] for p in picked])
        order = np.argsort(src, kind="stable")
        K = assemble_kernel_mat(self._R_desc, self._R_d_desc, self.tril_perms_lin, self.sig,
                                use_E_cstr=self.use_E_cstr, col_idxs=src[order])
        if hasattr(self.out, "detach"):
            idx = torch.from_numpy(dst[order]).to(self.out.device)
            self.out[:, idx] = torch.from_numpy(K).to(device=self.out.device, dtype=self.out.dtype)
        else:
            self.out[:, dst[order]] = K
        if self.callback is not None:
            for d, _ in picked:
                self.callback(len(d))
