"""Independent NumPy restatement of the ORCHESTRATION of cbm (src/offline/cbl_model_driver_offline.F90:108-213) and
soil_snow (src/science/soilsnow/cbl_soilsnow_main.F90:28-207): the statements between the calls, written from the Fortran
alone with the declarations' kinds (default REAL = float32, REAL(r_2) = float64).  The routines they call are passed in as
callables (the test hands in the oracle's single-routine hooks, each of which has its own NumPy cross-check), so what this
file pins is call order, the infiltration / puddle arithmetic and the bookkeeping either side.  Fields are the bound
(k, mp) arrays, updated in place.  TEST INFRASTRUCTURE ONLY."""
import numpy as np

F, D = np.float32, np.float64
DENSITY_LIQ, DENSITY_ICE, HL = F(1000.0), F(921.0), F(2.5014e6)          # cable_phys_constants_mod.F90:29,43,44
CGSNOW, CSICE, CSWAT = F(2090.0), F(2.100e3), F(4.218e3)                 # :38,41,42
LAKES = 16                                                               # cable_surface_types.F90:31


def cbm_head(T, zse):
    """cbl_model_driver_offline.F90:116-122 (lake refill) and :183-186 (albedo_T is checked by the caller after Albedo)."""
    wb1, sfc = T["ssnow_wb"][0], T["soil_sfc"][0]
    lake = (T["veg_iveg"][0] == LAKES) & (wb1 < sfc)
    zse1 = F(zse[0])
    T["ssnow_wbtot1"][0][lake] = (wb1.astype(F) * DENSITY_LIQ * zse1)[lake]
    wb1[lake] = sfc[lake]
    T["ssnow_wbtot2"][0][lake] = (wb1.astype(F) * DENSITY_LIQ * zse1)[lake]
    T["ssnow_wb_lake"][0][:] = T["ssnow_wb_lake"][0] + np.maximum(T["ssnow_wbtot2"][0] - T["ssnow_wbtot1"][0], F(0.))
    return int(lake.sum())


def cbm_tail(T):
    """cbl_model_driver_offline.F90:200-213 -> deltss, fev, fe, rnet, trad (EXP/** evaluated in float64, rounded once)."""
    tss, tv, transd = T["ssnow_tss"][0], T["canopy_tv"][0], T["rad_transd"][0]
    deltss = tss - T["ssnow_otss"][0]
    fev = (T["canopy_fevc"][0] + T["canopy_fevw"][0].astype(D)).astype(F)
    fe = (fev.astype(D) + T["canopy_fes"][0]).astype(F)
    rnet = T["canopy_fns"][0] + T["canopy_fnv"][0]
    p4 = lambda x: (x * x) * (x * x)
    arg = (F(1.) - transd) * p4(tv) + transd * p4(tss)
    trad = np.power(arg.astype(D), D(F(0.25))).astype(F)
    return dict(ssnow_deltss=deltss, canopy_fev=fev, canopy_fe=fe, canopy_rnet=rnet, rad_trad=trad)


def soil_snow(dels, T, zse, R, first_call=False):
    """cbl_soilsnow_main.F90:62-203 for cable_runtime%offline, redistrb = .FALSE.; R = the called routines, each working in
    place on T (first_call = the process's first call ever): snowcheck(), snowdensity(dels), snow_accum(dels), snow_melting(dels) -> snowmlt, snowl_adjust(),
    stempv(dels), remove_trans(), soilfreeze(), surfbv(dels)."""
    dels = F(dels)
    zse = np.asarray(zse, F)
    ms = zse.shape[0]
    s = lambda n: T[n][0]
    zsetot = F(0.)
    for k in range(ms):
        zsetot = zsetot + zse[k]
    tggav = np.zeros_like(s("ssnow_tggav"))
    for k in range(ms):
        tggav = tggav + ((zse[k] / zsetot) * T["ssnow_tgg"][k])
        T["soil_heat_cap_lower_limit"][k][:] = np.maximum(F(0.01), s("soil_css") * s("soil_rhosoil"))
    s("ssnow_tggav")[:] = tggav
    s("ssnow_t_snwlr")[:] = F(0.05)
    for n in ("fwtop1", "fwtop2", "fwtop3", "runoff", "rnof1", "rnof2", "smelt"):
        s("ssnow_" + n)[:] = F(0.0)
    T["ssnow_dtmlt"][...] = F(0.0)
    s("ssnow_osnowd")[:] = s("ssnow_snowd")
    T["ssnow_wbliq"][...] = T["ssnow_wb"] - T["ssnow_wbice"]
    ssat = s("soil_ssat")
    if first_call:                                                                   # IF (ktau <= 1), SAVE ktau (:60-62, :92-96)
        css, rhosoil, wb1, wbice1 = s("soil_css"), s("soil_rhosoil"), T["ssnow_wb"][0], T["ssnow_wbice"][0]
        xx = css * rhosoil
        heat = (((F(1.0) - ssat) * css * rhosoil).astype(D) + (wb1 - wbice1) * D(CSWAT) * D(DENSITY_LIQ)
                + wbice1 * D(CSICE) * D(DENSITY_ICE))
        T["ssnow_gammzz"][0][:] = (np.maximum(heat, xx.astype(D)) * D(zse[0])
                                   + ((F(1.) - s("ssnow_isflag").astype(F)) * CGSNOW * s("ssnow_snowd")).astype(D))
    for k in range(ms):
        T["ssnow_wblf"][k][:] = np.maximum(D(0.01), T["ssnow_wb"][k] - T["ssnow_wbice"][k]) / ssat.astype(D)
        T["ssnow_wbfice"][k][:] = T["ssnow_wbice"][k].astype(F) / ssat
    R["snowcheck"]()
    R["snowdensity"](dels)
    R["snow_accum"](dels)
    s("ssnow_smelt")[:] = R["snow_melting"](dels)
    R["snowl_adjust"]()
    R["stempv"](dels)
    isflag = s("ssnow_isflag")
    surface_temp = lambda: ((1 - isflag).astype(F) * T["ssnow_tgg"][0] + isflag.astype(F) * T["ssnow_tggsn"][0])
    s("ssnow_tss")[:] = surface_temp()
    s("ssnow_smelt")[:] = s("ssnow_smelt") + R["snow_melting"](dels)
    R["remove_trans"]()
    R["soilfreeze"]()
    totwet = s("canopy_precis") + s("ssnow_smelt")
    weting = (totwet.astype(D) + np.maximum(D(0.), s("ssnow_pudsto").astype(D) - s("canopy_fesp") / D(HL) * D(dels))).astype(F)
    cap = lambda k: D(F(0.95)) * (ssat.astype(D) - T["ssnow_wb"][k]) * D(zse[k]) * D(DENSITY_LIQ)
    sinfil1 = np.minimum(cap(0), weting.astype(D))
    sinfil2 = np.minimum(cap(1), (weting - sinfil1.astype(F)).astype(D))
    sinfil3 = np.minimum(cap(2), (weting - sinfil1.astype(F) - sinfil2.astype(F)).astype(D))
    s("ssnow_fwtop1")[:] = (sinfil1 / D(dels) - s("canopy_segg").astype(D)).astype(F)
    s("ssnow_fwtop2")[:] = (sinfil2 / D(dels)).astype(F)
    s("ssnow_fwtop3")[:] = (sinfil3 / D(dels)).astype(F)
    pudsto = np.maximum(D(0.), weting.astype(D) - sinfil1 - sinfil2 - sinfil3).astype(F)
    rnof1 = np.maximum(F(0.), pudsto - s("ssnow_pudsmx"))
    s("ssnow_rnof1")[:] = rnof1
    s("ssnow_pudsto")[:] = pudsto - rnof1
    R["surfbv"](dels)
    s("ssnow_smelt")[:] = s("ssnow_smelt") / dels
    s("ssnow_tss")[:] = surface_temp()
    T["ssnow_wbliq"][...] = T["ssnow_wb"] - T["ssnow_wbice"]
    sd = T["ssnow_sdepth"]
    s("ssnow_totsdepth")[:] = (sd[0] + sd[1]) + sd[2]
    wbtot = np.zeros(ssat.shape[0], D)
    for k in range(ms):
        wbtot = wbtot + (T["ssnow_wbliq"][k] * D(DENSITY_LIQ) + T["ssnow_wbice"][k] * D(DENSITY_ICE)) * D(zse[k])
    s("ssnow_wbtot")[:] = wbtot
