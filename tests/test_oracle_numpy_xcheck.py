"""CPU: the C++ oracle's smoisturev / trimb against an INDEPENDENT NumPy restatement of the same Fortran
(tests/np_restatement.py) on states taken from a running simulation, incl. frozen and saturated
layers.  Both are fp64 with the same libm pow, so they must agree to rounding."""
import ctypes as C

import numpy as np

from cable_b200 import lib
from oracle.pyoracle import Oracle
from np_restatement import smoisturev as smoisturev_np, trimb as trimb_np
from util import DELS, make_case


def test_trimb_numpy_vs_oracle():
    from oracle import pyoracle
    L = pyoracle.load(True)
    rng = np.random.default_rng(3)
    for kmax in (6, 9):
        n = 40
        a = -rng.uniform(0.0, 1.0, (kmax, n)); c = -rng.uniform(0.0, 1.0, (kmax, n)); b = 1.0 - a - c
        rhs = rng.normal(280.0, 10.0, (kmax, n))
        want = trimb_np(a, b, c, rhs)
        got = np.ascontiguousarray(rhs.copy())
        L.oracle_trimb(n, np.ascontiguousarray(a).ctypes.data, np.ascontiguousarray(b).ctypes.data,
                       np.ascontiguousarray(c).ctypes.data, got.ctypes.data, kmax)
        np.testing.assert_allclose(got, want, rtol=1e-14, atol=0)


def test_smoisturev_numpy_vs_oracle():
    for new_speed in (0, 1):
        cfg = lib.default_cfg(); cfg.l_new_runoff_speed = new_speed
        cfg, grid, T, F = make_case(500, cfg=cfg, start_doy=30)
        o = Oracle(T, cfg, cr_math=True)
        o._lib.oracle_run_smoisturev.argtypes = [C.c_void_p, C.c_float]
        o._lib.oracle_run_smoisturev.restype = None
        for k in range(10):
            F.fill(T, k)
            o.cbm(k + 1, DELS)
        assert (T["ssnow_wbice"] > 0.05).any(), "case must contain frozen layers"
        # a state smoisturev could be handed: after cbm, with infiltration fluxes of this step
        args = dict(dels=DELS, wb=T["ssnow_wb"].copy(), wbice=T["ssnow_wbice"].copy(), tgg=T["ssnow_tgg"].copy(),
                    gammzz=T["ssnow_gammzz"].copy(), fwtop=[T["ssnow_fwtop1"][0].copy(), T["ssnow_fwtop2"][0].copy(), T["ssnow_fwtop3"][0].copy()],
                    ssat=T["soil_ssat"][0], sfc=T["soil_sfc"][0], hyds=T["soil_hyds"][0], hsbh=T["soil_hsbh"][0], ibp2=T["soil_ibp2"][0],
                    i2bp3=T["soil_i2bp3"][0], pwb_min=T["soil_pwb_min"][0], zse=np.array(list(cfg.zse), np.float32),
                    zshh=np.array(list(cfg.zshh), np.float32), frozen_limit=cfg.frozen_limit, l_new_runoff_speed=bool(new_speed))
        want = smoisturev_np(**args)
        o._lib.oracle_run_smoisturev(o._h, DELS)
        for name, key in (("ssnow_wb", "wb"), ("ssnow_wbice", "wbice"), ("ssnow_wblf", "wblf")):
            np.testing.assert_allclose(T[name], want[key], rtol=2e-13, atol=1e-300, err_msg=name)
        np.testing.assert_allclose(T["ssnow_tgg"], want["tgg"], rtol=2e-7, err_msg="tgg")
        np.testing.assert_allclose(T["ssnow_rnof2"][0], want["rnof2"], rtol=2e-7, atol=1e-12, err_msg="rnof2")
        assert (want["rnof2"] > 0).any() and np.abs(want["wb"] - args["wb"]).max() > 1e-6     # the routine did something
