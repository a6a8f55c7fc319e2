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
    stempv(dels), remove_trans(), soilfreeze(), surfbv(dels) and, with redistrb, hydraulic_redistribution(dels)."""
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
    if "hydraulic_redistribution" in R:                                              # IF (redistrb) (:186-187)
        R["hydraulic_redistribution"](dels)
    s("ssnow_smelt")[:] = s("ssnow_smelt") / dels
    s("ssnow_tss")[:] = surface_temp()
    T["ssnow_wbliq"][...] = T["ssnow_wb"] - T["ssnow_wbice"]
    sd = T["ssnow_sdepth"]
    s("ssnow_totsdepth")[:] = (sd[0] + sd[1]) + sd[2]
    wbtot = np.zeros(ssat.shape[0], D)
    for k in range(ms):
        wbtot = wbtot + (T["ssnow_wbliq"][k] * D(DENSITY_LIQ) + T["ssnow_wbice"][k] * D(DENSITY_ICE)) * D(zse[k])
    s("ssnow_wbtot")[:] = wbtot


def hydraulic_redistribution(dels, T, zse, wilt_param, satu_param):
    """cbl_hyd_redistrib.F90:13-221 (redistrb = .TRUE.): every working variable default REAL, ssnow%wb REAL(r_2); EXP / **
    with a real exponent evaluated in float64 and rounded once.  Updates T['ssnow_wb'] in place."""
    dels = F(dels)
    zse = np.asarray(zse, F)
    ms = zse.shape[0]
    s = lambda n: T[n][0]
    wb, wbice, froot = T["ssnow_wb"], T["ssnow_wbice"], T["veg_froot"]
    swilt, sfc, ssat = s("soil_swilt"), s("soil_sfc"), s("soil_ssat")
    n_hr, wpsy50, n_vg, alpha_vg, crt = F(3.22), F(-1.0), F(2.06), F(0.00423), F(125.0)
    m_vg = F(1.0) - F(1.0) / n_vg
    wilt_param, satu_param = F(wilt_param), F(satu_param)
    pw = lambda x, y: np.power(np.asarray(x, D), D(y)).astype(F)
    zsetot = F(0.)
    for k in range(ms):
        zsetot = zsetot + zse[k]
    totalice = np.zeros(swilt.shape[0], F)
    for k in range(ms):
        totalice = (totalice.astype(D) + wbice[k] * D(zse[k]) / D(zsetot)).astype(F)
    dtran = np.where((s("canopy_fevc") < 10.0) & (totalice < F(1.e-2)), F(1.0), F(0.0)).astype(F)
    hr_pft = (s("veg_iveg") == 2) | (s("veg_iveg") == 7)               # evergreen_broadleaf, c4_grassland

    def potentials():
        wpsy, c_hr = [], []
        for k in range(ms):
            s_vg = np.minimum(F(1.0), np.maximum(F(1.0E-4), wb[k].astype(F) - swilt) / (ssat - swilt))
            with np.errstate(all="ignore"):
                w = -(F(1.0) / alpha_vg * pw(pw(s_vg, -(F(1.0) / m_vg)) - F(1.0), F(1) / n_vg) * F(100) * F(1.0E-6))
                c = F(1.) / (F(1) + pw(w / wpsy50, n_hr))
            wpsy.append(w); c_hr.append(c)
        return wpsy, c_hr

    def exchange(k, j, hr_term, avail_of):
        hkj = hr_term * F(1.0E-2) / F(3600.0) * dels
        hjk = F(-1.0) * hkj
        hkj = hkj / zse[k]
        hjk = hjk / zse[j]
        hkj = np.where(hr_pft, hkj, F(0.0)).astype(F)
        hjk = np.where(hr_pft, hjk, F(0.0)).astype(F)
        down, up = hkj < F(0.0), ~(hkj < F(0.0)) & (hjk < F(0.0))
        # WHERE (hr_perTime(:,k,j) < 0): layer k gives to layer j
        t1 = np.maximum(np.maximum(hkj, F(-1.0) * wilt_param * avail_of(k)),
                        F(-1.0) * satu_param * np.maximum(D(0.0), ssat.astype(D) - wb[j]).astype(F) * zse[j] / zse[k])
        # ELSEWHERE (hr_perTime(:,j,k) < 0): layer j gives to layer k
        t2 = np.maximum(np.maximum(hjk, F(-1.0) * wilt_param * avail_of(j)),
                        F(-1.0) * satu_param * np.maximum(D(0.0), ssat.astype(D) - wb[k]).astype(F) * zse[k] / zse[j])
        new_kj = np.where(down, t1, np.where(up, F(-1.0) * t2 * zse[j] / zse[k], hkj)).astype(F)
        new_jk = np.where(down, F(-1.0) * t1 * zse[k] / zse[j], np.where(up, t2, hjk)).astype(F)
        wb[k][:] = wb[k] + new_kj.astype(D)
        wb[j][:] = wb[j] + new_jk.astype(D)

    # deep -> shallow pairs, limited by a third of the way from wilting point to field capacity (:98-143)
    wpsy, c_hr = potentials()
    avail1 = lambda l: np.maximum(D(0.0), wb[l] - (swilt + (sfc - swilt) / F(3.)).astype(D)).astype(F)
    for k in range(ms - 1, 1, -1):
        for j in range(k - 1, 0, -1):
            froot_x = np.maximum(F(0.01), np.maximum(froot[k], froot[j]))
            hr_term = crt * (wpsy[j] - wpsy[k]) * np.maximum(c_hr[k], c_hr[j]) * (froot[k] * froot[j]) / (F(1) - froot_x) * dtran
            exchange(k, j, hr_term, avail1)
    # shallow -> deep pairs, only the water above field capacity moves (:146-219)
    dtran = np.where(s("met_tk") < F(273.16) + F(5.), F(0.0), dtran).astype(F)
    wpsy, c_hr = potentials()
    avail2 = lambda l: np.maximum(D(0.0), wb[l] - sfc.astype(D)).astype(F)
    for k in range(0, ms - 2):
        for j in range(k + 1, ms - 1):
            froot_x = np.maximum(F(0.01), np.maximum(froot[k], froot[j]))
            hr_term = (crt * (wpsy[j] - wpsy[k]) * np.maximum(c_hr[k], c_hr[j])
                       * (np.maximum(F(0.01), froot[k]) * np.maximum(F(0.01), froot[j])) / (F(1) - froot_x) * dtran)
            exchange(k, j, hr_term, avail2)
