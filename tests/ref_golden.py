"""Loader of tests/golden/ref_backend.npz: inputs + outputs of the reference's own compiled code (make_ref_golden.py)."""
import ctypes as C
import os

import numpy as np

from conftest import GOLDEN
from lfvio import backend as B
from lfvio import prestep as PS


def load():
    return np.load(os.path.join(GOLDEN, "ref_backend.npz"))


def window(g, opts, prefix="w_") -> B.WindowProblem:
    """Rebuild the stored WindowProblem (state, factor records, prior) around the given solver options."""
    n_proj = len(g[prefix + "proj"]) // C.sizeof(B.ProjFactor)
    n_imu = len(g[prefix + "imu"]) // C.sizeof(B.ImuFactor)
    proj = list((B.ProjFactor * n_proj).from_buffer_copy(g[prefix + "proj"].tobytes()))
    imu = list((B.ImuFactor * n_imu).from_buffer_copy(g[prefix + "imu"].tobytes()))
    pr = B.PriorData()
    ids, idx = g[prefix + "prior_ids"], g[prefix + "prior_idx"]
    n = len(g[prefix + "prior_r0"])
    pr.block_id[: len(ids)], pr.block_idx[: len(ids)] = ids, idx
    pr.x0[: len(g[prefix + "prior_x0"])] = g[prefix + "prior_x0"]
    pr.J0[: n * n] = g[prefix + "prior_J0"].ravel()
    pr.r0[:n] = g[prefix + "prior_r0"]
    pr.c.n, pr.c.n_blocks = n, len(ids)
    opts.row, opts.tr = int(g[prefix + "row_tr"][0]), float(g[prefix + "row_tr"][1])
    return B.WindowProblem(len(g[prefix + "pose"]), g[prefix + "pose"], g[prefix + "sb"], g[prefix + "ex"], g[prefix + "feat"],
                           float(g[prefix + "td"][0]), proj, imu, pr, opts)


def preintegration_inputs(g):
    n = len(g["pre_intervals"]) // C.sizeof(PS.ImuInterval)
    ivs = list((PS.ImuInterval * n).from_buffer_copy(g["pre_intervals"].tobytes()))
    want = list((B.ImuFactor * n).from_buffer_copy(g["pre_out"].tobytes()))
    return ivs, g["pre_dt"], g["pre_acc"], g["pre_gyr"], tuple(g["pre_noise_g"][:4]), g["pre_noise_g"][4:], want
