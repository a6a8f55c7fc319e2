"""Independent NumPy / scalar restatements of the snow-pack routines at the head of soil_snow, written from the Fortran
alone as cross-checks of the C++ oracle:
  snowdensity   src/science/soilsnow/cbl_snowDensity.F90:9-102
  snow_accum    src/science/soilsnow/cbl_snowAccum.F90:10-188
  snow_melting  src/science/soilsnow/cbl_snowMelt.F90:9-120
Default REAL = np.float32, REAL(r_2) = np.float64, Fortran operation order; EXP evaluated in float64 and rounded once (the
correctly rounded oracle build's convention).  State is a dict of registry-layout arrays ((k, mp), layer index first),
modified in place.  TEST INFRASTRUCTURE ONLY."""
import numpy as np

F, D = np.float32, np.float64
TFRZ, CHLF, CHL, CSWAT, CGSNOW, DENSITY_LIQ = F(273.16), F(0.334e6), F(2.5014e6), F(4.218e3), F(2090.0), F(1000.0)


def _exp(x):
    with np.errstate(all="ignore"):
        return np.exp(np.asarray(x, D)).astype(F)


def snowdensity(S, dels, max_ssdn, max_sconds):
    dels, max_ssdn, max_sconds = F(dels), F(max_ssdn), F(max_sconds)
    ssdn, tgg1, tggsn, snowd, isflag = S["ssnow_ssdn"], S["ssnow_tgg"][0], S["ssnow_tggsn"], S["ssnow_snowd"][0], S["ssnow_isflag"][0]
    smass, sdepth, sconds, t_snwlr = S["ssnow_smass"], S["ssnow_sdepth"], S["ssnow_sconds"], S["ssnow_t_snwlr"][0]
    sc = lambda d: np.maximum(F(0.2), np.minimum(F(2.876e-6) * (d * d) + F(0.074), max_sconds))
    merge = lambda d: np.where(d >= F(150.0), F(0.046), F(0.0)).astype(F)
    with np.errstate(all="ignore"):
        m0 = (snowd > F(0.1)) & (isflag == 0)
        tmin1 = np.minimum(TFRZ, tgg1)
        d = ssdn[0]
        d1 = np.minimum(max_ssdn, np.maximum(F(120.0), d + dels * d * F(3.1e-6) * _exp(F(-0.03) * (F(273.15) - tmin1) - merge(d) * (d - F(150.0)))))
        d1 = np.minimum(max_ssdn, d1 + dels * F(9.806) * d1 * F(0.75) * snowd
                        / (F(3.0e7) * _exp(F(0.021) * d1 + F(0.081) * (F(273.15) - np.minimum(TFRZ, tgg1)))))
        d1 = np.where(S["soil_isoilm"][0] != 9, np.minimum(F(450.0), d1), d1).astype(F)
        m1 = isflag == 1
        e = [None] * 3
        for l in range(3):
            dl = ssdn[l]
            e[l] = dl + dels * dl * F(3.1e-6) * _exp(F(-0.03) * (F(273.15) - np.minimum(TFRZ, tggsn[l])) - merge(dl) * (dl - F(150.0)))
        den = lambda l: F(3.0e7) * _exp(F(.021) * e[l] + F(0.081) * (F(273.15) - np.minimum(TFRZ, tggsn[l])))
        e[0] = e[0] + dels * F(9.806) * e[0] * t_snwlr * e[0] / den(0)
        e[1] = e[1] + dels * F(9.806) * e[1] * (t_snwlr * e[0] + F(0.5) * smass[1]) / den(1)
        e[2] = e[2] + dels * F(9.806) * e[2] * (t_snwlr * e[0] + smass[1] + F(0.5) * smass[2]) / den(2)
        ssdnn1 = (e[0] * smass[0] + e[1] * smass[1] + e[2] * smass[2]) / snowd
        for l in range(3):
            sdepth[l] = np.where(m1, smass[l] / e[l], sdepth[l])
            sconds[l] = np.where(m1, sc(e[l]), np.where(m0, sc(d1), sconds[l]))
            ssdn[l] = np.where(m1, e[l], np.where(m0, d1, ssdn[l]))
        S["ssnow_ssdnn"][0] = np.where(m1, ssdnn1, np.where(m0, d1, S["ssnow_ssdnn"][0]))


def snow_accum(S, dels, max_ssdn):
    dels, max_ssdn = F(dels), F(max_ssdn)
    precis, precip_sn = S["canopy_precis"][0], S["met_precip_sn"][0]
    snowd, osnowd, isflag, isoilm = S["ssnow_snowd"][0], S["ssnow_osnowd"][0], S["ssnow_isflag"][0], S["soil_isoilm"][0]
    ssdn, ssdnn, tgg, tggsn, dtmlt = S["ssnow_ssdn"], S["ssnow_ssdnn"][0], S["ssnow_tgg"], S["ssnow_tggsn"], S["ssnow_dtmlt"]
    smass, sdepth, gammzz = S["ssnow_smass"], S["ssnow_sdepth"], S["ssnow_gammzz"]
    fess, fes_cor, cls = S["canopy_fess"][0], S["canopy_fes_cor"][0], S["ssnow_cls"][0]
    segg, evapsn = S["canopy_segg"][0], S["ssnow_evapsn"][0]
    mx, mn = max, min
    for i in range(snowd.shape[0]):
        if precis[i] > 0 and isflag[i] == 0:
            snowd[i] = mx(snowd[i] + precip_sn[i], F(0.0))
            precis[i] = precis[i] - precip_sn[i]
            ssdn[0, i] = mx(F(120.0), ssdn[0, i] * osnowd[i] / mx(F(0.01), snowd[i]) + F(120.0) * precip_sn[i] / mx(F(0.01), snowd[i]))
            ssdnn[i] = ssdn[0, i]
            if precis[i] > 0 and tgg[0, i] < TFRZ:
                snowd[i] = mx(snowd[i] + precis[i], F(0.0))
                dt = precis[i] * CHLF / (F(gammzz[0, i]) + CSWAT * precis[i])
                tgg[0, i] = tgg[0, i] + dt
                dtmlt[0, i] = dtmlt[0, i] + dt
                ssdn[0, i] = mn(max_ssdn, mx(F(120.0), ssdn[0, i] * osnowd[i] / mx(F(0.01), snowd[i])
                                             + DENSITY_LIQ * precis[i] / mx(F(0.01), snowd[i])))
                if isoilm[i] != 9:
                    ssdn[0, i] = mn(F(450.0), ssdn[0, i])
                precis[i] = F(0.0)
                ssdnn[i] = ssdn[0, i]
        if precis[i] > 0 and isflag[i] > 0:
            snowd[i] = mx(snowd[i] + precip_sn[i], F(0.0))
            precis[i] = precis[i] - precip_sn[i]
            osm = smass[0, i]
            smass[0, i] = smass[0, i] + precip_sn[i]
            ssdn[0, i] = mx(F(120.0), ssdn[0, i] * osm / smass[0, i] + F(120.0) * precip_sn[i] / smass[0, i])
            sdepth[0, i] = mx(F(0.02), smass[0, i] / ssdn[0, i])
            if precis[i] > 0:
                snowd[i] = mx(snowd[i] + precis[i], F(0.0))
                for l in range(3):
                    sgamm = ssdn[l, i] * CGSNOW * sdepth[l, i]
                    osm = smass[l, i]
                    dt = precis[i] * CHLF * osm / (sgamm * osnowd[i])
                    tggsn[l, i] = tggsn[l, i] + dt
                    if l == 0:
                        dtmlt[0, i] = dtmlt[0, i] + dt
                    smass[l, i] = smass[l, i] + precis[i] * osm / osnowd[i]
                    ssdn[l, i] = mx(F(120.0), mn(ssdn[l, i] * osm / smass[l, i] + DENSITY_LIQ * (F(1.0) - osm / smass[l, i]), max_ssdn))
                    if isoilm[i] != 9:
                        ssdn[l, i] = mn(F(450.0), ssdn[l, i])
                    sdepth[l, i] = smass[l, i] / ssdn[l, i]
                precis[i] = F(0.0)
    segg[:] = ((fess + fes_cor) / D(CHL)).astype(F)
    evapsn[:] = F(0.0)
    for i in np.flatnonzero(cls == F(1.1335)):
        ftot = fess[i] + fes_cor[i]                                     # r_2
        evapsn[i] = F(D(dels) * ftot / D(CHL + CHLF))
        xxx = evapsn[i]
        if isflag[i] == 0 and ftot > 0.0:
            evapsn[i] = mn(snowd[i], xxx)
        if isflag[i] > 0 and ftot > 0.0:
            evapsn[i] = mn(F(0.9) * smass[0, i], xxx)
        snowd[i] = snowd[i] - evapsn[i]
        if isflag[i] > 0:
            smass[0, i] = smass[0, i] - evapsn[i]
            sdepth[0, i] = mx(F(0.02), smass[0, i] / ssdn[0, i])
        segg[i] = (CHL + CHLF) * (xxx - evapsn[i]) / CHL / dels


def snow_melting(S, dels, max_ssdn):
    """-> snowmlt (mp) f32.  Per tile, in the Fortran's statement order (the WHERE masks are evaluated where they stand)."""
    dels, max_ssdn = F(dels), F(max_ssdn)
    snowd, isflag, isoilm = S["ssnow_snowd"][0], S["ssnow_isflag"][0], S["soil_isoilm"][0]
    tgg, tggsn, gammzz, dtmlt = S["ssnow_tgg"], S["ssnow_tggsn"], S["ssnow_gammzz"], S["ssnow_dtmlt"]
    ssdn, sdepth, smass = S["ssnow_ssdn"], S["ssnow_sdepth"], S["ssnow_smass"]
    mp = snowd.shape[0]
    snowmlt = np.zeros(mp, F)
    for j in range(mp):
        if snowd[j] > 0 and isflag[j] == 0 and tgg[0, j] >= TFRZ:
            snowflx = F(D(tgg[0, j] - TFRZ) * gammzz[0, j])
            snowmlt[j] = min(snowflx / CHLF, snowd[j])
            dtmlt[0, j] = F(D(dtmlt[0, j]) + D(snowmlt[j] * CHLF) / gammzz[0, j])
            snowd[j] = snowd[j] - snowmlt[j]
            tgg[0, j] = F(D(tgg[0, j]) - D(snowmlt[j] * CHLF) / gammzz[0, j])
        if snowd[j] > 0 and isflag[j] > 0:
            sm = [F(0.0)] * 4                                            # smelt1(j, 0:3)
            for k in (1, 2, 3):
                l = k - 1
                sgamm = ssdn[l, j] * CGSNOW * sdepth[l, j]
                snowflx = sm[k - 1] * CHLF / dels
                tggsn[l, j] = tggsn[l, j] + (snowflx * dels + sm[k - 1] * CSWAT * (TFRZ - tggsn[l, j])) / (sgamm + CSWAT * sm[k - 1])
                osm = smass[l, j]
                smass[l, j] = smass[l, j] + sm[k - 1]
                ssdn[l, j] = max(F(120.0), min(ssdn[l, j] * osm / smass[l, j] + DENSITY_LIQ * (F(1.0) - osm / smass[l, j]), max_ssdn))
                if isoilm[j] != 9:
                    ssdn[l, j] = min(F(450.0), ssdn[l, j])
                sdepth[l, j] = smass[l, j] / ssdn[l, j]
                sgamm = smass[l, j] * CGSNOW
                sm[k - 1] = F(0.0); sm[k] = F(0.0)
                if tggsn[l, j] > TFRZ:
                    snowflx = (tggsn[l, j] - TFRZ) * sgamm
                    sm[k] = min(snowflx / CHLF, F(0.6) * smass[l, j])
                    dtmlt[l, j] = dtmlt[l, j] + sm[k] * CHLF / sgamm
                    smass[l, j] = smass[l, j] - sm[k]
                    tggsn[l, j] = tggsn[l, j] - sm[k] * CHLF / sgamm
                    sdepth[l, j] = smass[l, j] / ssdn[l, j]
            snowmlt[j] = sm[1] + sm[2] + sm[3]
            snowd[j] = snowd[j] - snowmlt[j]
    return snowmlt


def snowcheck(S, snmin):
    """cbl_snowCheck.F90:9-100: the one-layer <-> three-layer regime of the pack."""
    snmin = F(snmin)
    snowd, isflag, ssdnn, t_snwlr = S["ssnow_snowd"][0], S["ssnow_isflag"][0], S["ssnow_ssdnn"][0], S["ssnow_t_snwlr"][0]
    ssdn, tggsn, tgg, sdepth, smass = S["ssnow_ssdn"], S["ssnow_tggsn"], S["ssnow_tgg"], S["ssnow_sdepth"], S["ssnow_smass"]
    for j in range(snowd.shape[0]):
        if snowd[j] <= 0:
            isflag[j] = 0
            ssdn[:, j] = F(120.0); ssdnn[j] = F(120.0); tggsn[:, j] = TFRZ
            sdepth[0, j] = snowd[j] / ssdn[0, j]; sdepth[1, j] = F(0.0); sdepth[2, j] = F(0.0)
            smass[0, j] = snowd[j]; smass[1, j] = F(0.0); smass[2, j] = F(0.0)
        elif snowd[j] < snmin * ssdnn[j]:
            if isflag[j] == 1:
                ssdn[0, j] = ssdnn[j]
                tgg[0, j] = tggsn[0, j]
            isflag[j] = 0
            ssdnn[j] = min(F(400.0), max(F(120.0), ssdn[0, j]))
            tggsn[:, j] = min(TFRZ, tgg[0, j])
            sdepth[0, j] = snowd[j] / ssdn[0, j]; sdepth[1, j] = F(0.0); sdepth[2, j] = F(0.0)
            smass[0, j] = snowd[j]; smass[1, j] = F(0.0); smass[2, j] = F(0.0)
            ssdn[:, j] = ssdnn[j]
        else:
            if isflag[j] == 0:
                tggsn[:, j] = min(TFRZ, tgg[0, j])
                ssdn[1, j] = ssdn[0, j]; ssdn[2, j] = ssdn[0, j]
                sdepth[0, j] = t_snwlr[j]
                smass[0, j] = t_snwlr[j] * ssdn[0, j]
                smass[1, j] = (snowd[j] - smass[0, j]) * F(0.4)
                smass[2, j] = (snowd[j] - smass[0, j]) * F(0.6)
                sdepth[1, j] = smass[1, j] / ssdn[1, j]
                sdepth[2, j] = smass[2, j] / ssdn[2, j]
                ssdnn[j] = (ssdn[0, j] * smass[0, j] + ssdn[1, j] * smass[1, j] + ssdn[2, j] * smass[2, j]) / snowd[j]
            isflag[j] = 1


def snowl_adjust(S, max_ssdn):
    """cbl_snowl_adjust.F90:9-156: re-partition of mass between the three layers (excd, excm, frac, xfrac are r_2)."""
    max_ssdn = F(max_ssdn)
    snowd, isflag, ssdnn, t_snwlr = S["ssnow_snowd"][0], S["ssnow_isflag"][0], S["ssnow_ssdnn"][0], S["ssnow_t_snwlr"][0]
    ssdn, tggsn, sdepth, smass = S["ssnow_ssdn"], S["ssnow_tggsn"], S["ssnow_sdepth"], S["ssnow_smass"]
    for j in np.flatnonzero(isflag > 0):
        if sdepth[0, j] > t_snwlr[j]:
            excd = D(sdepth[0, j] - t_snwlr[j])
            excm = excd * D(ssdn[0, j])
            sdepth[0, j] = sdepth[0, j] - F(excd)
            smass[0, j] = smass[0, j] - F(excm)
            osm = smass[1, j]
            smass[1, j] = max(F(0.01), smass[1, j] + F(excm))
            ssdn[1, j] = F(max(120.0, min(D(max_ssdn), D(ssdn[1, j] * osm / smass[1, j]) + D(ssdn[0, j]) * excm / D(smass[1, j]))))
            sdepth[1, j] = smass[1, j] / ssdn[1, j]
            tggsn[1, j] = F(D(tggsn[1, j] * osm / smass[1, j]) + D(tggsn[0, j]) * excm / D(smass[1, j]))
            smass[2, j] = max(F(0.01), snowd[j] - smass[0, j] - smass[1, j])
        else:
            excd = D(t_snwlr[j] - sdepth[0, j])
            excm = excd * D(ssdn[1, j])
            osm = smass[0, j]
            smass[0, j] = smass[0, j] + F(excm)
            sdepth[0, j] = t_snwlr[j]
            ssdn[0, j] = F(max(120.0, min(D(max_ssdn), D(ssdn[0, j] * osm / smass[0, j]) + D(ssdn[1, j]) * excm / D(smass[0, j]))))
            tggsn[0, j] = F(D(tggsn[0, j] * osm / smass[0, j]) + D(tggsn[1, j]) * excm / D(smass[0, j]))
            smass[1, j] = max(F(0.01), smass[1, j] - F(excm))
            sdepth[1, j] = smass[1, j] / ssdn[1, j]
            smass[2, j] = max(F(0.01), snowd[j] - smass[0, j] - smass[1, j])
    for j in np.flatnonzero(isflag > 0):
        frac = D(smass[1, j] / max(F(0.02), smass[2, j]))
        xfrac = D(F(2.0) / F(3.0)) / frac
        if xfrac > 1.0:
            excm = (xfrac - 1.0) * D(smass[1, j])
            osm = smass[1, j]
            smass[1, j] = max(F(0.01), smass[1, j] + F(excm))
            tggsn[1, j] = tggsn[1, j] * osm / smass[1, j] + tggsn[2, j] * F(excm) / smass[1, j]
            ssdn[1, j] = max(F(120.0), min(max_ssdn, ssdn[1, j] * osm / smass[1, j] + ssdn[2, j] * F(excm) / smass[1, j]))
            smass[2, j] = max(F(0.01), snowd[j] - smass[0, j] - smass[1, j])
            sdepth[2, j] = max(F(0.02), smass[2, j] / ssdn[2, j])
        else:
            excm = (1.0 - xfrac) * D(smass[1, j])
            smass[1, j] = max(F(0.01), smass[1, j] - F(excm))
            sdepth[1, j] = max(F(0.02), smass[1, j] / ssdn[1, j])
            osm = smass[2, j]
            smass[2, j] = max(F(0.01), snowd[j] - smass[0, j] - smass[1, j])
            tggsn[2, j] = tggsn[2, j] * osm / smass[2, j] + tggsn[1, j] * F(excm) / smass[2, j]
            ssdn[2, j] = max(F(120.0), min(max_ssdn, ssdn[2, j] * osm / smass[2, j] + ssdn[1, j] * F(excm) / smass[2, j]))
            sdepth[2, j] = smass[2, j] / ssdn[2, j]
        isflag[j] = 1
        ssdnn[j] = (ssdn[0, j] * sdepth[0, j] + ssdn[1, j] * sdepth[1, j] + ssdn[2, j] * sdepth[2, j]) \
            / (sdepth[0, j] + sdepth[1, j] + sdepth[2, j])
