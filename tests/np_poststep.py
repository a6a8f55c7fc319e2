"""Independent NumPy restatement of the per-tile statements the offline driver runs after CALL cbm:
the dels scaling of smelt / runoff (src/offline/cable_serial.F90:602-605), sumcflux for icycle == 0
(src/science/casa-cnp/casa_sumcflux.F90:76-102), mass_balance and energy_balance (src/offline/cable_checks.F90:472-618,
soil_struc = 'default', qrecharge = 0 offline).  Written from the Fortran alone: default REAL = float32, REAL(r_2) =
float64, Fortran operation order and promotion, x**4 by repeated multiplication.  State arrays are the bound (k, mp)
fields; the driver-side arrays live in the dict D under the names of oracle.pyoracle.DRIVER_ARRAYS.
TEST INFRASTRUCTURE ONLY."""
import numpy as np

F, D64 = np.float32, np.float64
SBOLTZ, EMLEAF, EMSOIL = F(5.67e-8), F(1.0), F(1.0)                    # cable_phys_constants_mod.F90:25-27


def scale_by_dels(T, dels):
    for n in ("ssnow_smelt", "ssnow_rnof1", "ssnow_rnof2", "ssnow_runoff"):
        T[n][0][:] = T[n][0] * F(dels)


def sumcflux(T, Dr, ktau, kstart, dels):
    dels = F(dels)
    pairs = (("sumpn", "fpn"), ("sumrd", "frday"), ("dsumpn", "fpn"), ("dsumrd", "frday"), ("sumrpw", "frpw"), ("sumrpr", "frpr"),
             ("sumrp", "frp"), ("dsumrp", "frp"), ("sumrs", "frs"))
    for acc, flux in pairs:
        x = T["canopy_" + flux][0] * dels
        Dr[acc][:] = x if ktau == kstart else Dr[acc] + x
    T["canopy_fnee"][0][:] = T["canopy_fpn"][0] + T["canopy_frs"][0] + T["canopy_frp"][0]


def mass_balance(T, Dr, ktau, dels):
    s = lambda n: T[n][0]
    dels = F(dels)
    if ktau == 1:
        Dr["owb"][:] = s("ssnow_wbtot")
    delwb = s("ssnow_wbtot") - Dr["owb"]
    Dr["owb"][:] = s("ssnow_wbtot")
    surface = s("met_precip") - s("canopy_delwc") - s("ssnow_snowd") + s("ssnow_osnowd") - s("ssnow_runoff")
    evap = (s("canopy_fevw").astype(D64) + s("canopy_fevc") + s("canopy_fes") / s("ssnow_cls").astype(D64)) * D64(dels) \
        / s("air_rlam").astype(D64)
    Dr["wbal"][:] = (surface.astype(D64) - evap - delwb - 0.0).astype(F)
    if ktau == 1:
        for n in ("wbal_tot", "precip_tot", "rnoff_tot", "evap_tot"):
            Dr[n][:] = F(0.)
    if ktau > 10:
        Dr["wbal_tot"][:] = Dr["wbal_tot"] + Dr["wbal"]
        Dr["precip_tot"][:] = Dr["precip_tot"] + s("met_precip")
        Dr["rnoff_tot"][:] = Dr["rnoff_tot"] + s("ssnow_rnof1") + s("ssnow_rnof2")
        Dr["evap_tot"][:] = (Dr["evap_tot"].astype(D64)
                             + (s("canopy_fev").astype(D64) + s("canopy_fes") / s("ssnow_cls").astype(D64)) * D64(dels)
                             / s("air_rlam").astype(D64)).astype(F)


def energy_balance(T, Dr):
    s = lambda n: T[n][0]
    p4 = lambda x: (x * x) * (x * x)
    fsd, alb, transd, tv = T["met_fsd"], T["rad_albedo"], s("rad_transd"), s("canopy_tv")
    Dr["radbal"][:] = (fsd[0] + fsd[1] + s("met_fld") - alb[0] * fsd[0] - alb[1] * fsd[1]
                       - (EMSOIL * SBOLTZ * transd * p4(s("ssnow_otss"))) - (EMLEAF * SBOLTZ * (F(1) - transd) * p4(tv))
                       - s("canopy_fnv") - s("canopy_fns"))
    Dr["ebalsoil"][:] = ((s("canopy_fns").astype(D64) - s("canopy_fes")) - s("canopy_fhs").astype(D64)
                         - s("canopy_ga").astype(D64)).astype(F)
    Dr["ebalveg"][:] = s("canopy_fnv") - s("canopy_fev") - s("canopy_fhv")
    q = T["rad_qcan"]
    head = ((q[0] + q[1]) + (q[2] + q[3]) + s("rad_qssabs") + s("met_fld") - SBOLTZ * EMLEAF * p4(tv) * (F(1) - transd)
            - s("rad_flws") * transd - s("canopy_fev"))
    Dr["ebal"][:] = (head.astype(D64) - s("canopy_fes") - s("canopy_fh").astype(D64) - s("canopy_ga").astype(D64)).astype(F)
    Dr["ebal_tot"][:] = Dr["ebal_tot"] + Dr["ebal"]
    Dr["radbalsum"][:] = Dr["radbalsum"] + Dr["radbal"]
