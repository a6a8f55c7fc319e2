"""GPU: the CUDA path, called through the C ABI, against the oracle on identical seeded inputs.

Tolerances are BASELINE.json's north star: 1e-4 relative for fp32 fields, 1e-6 for fp64 fields, evaluated per
element against max(|a|,|b|, 1e-3 * field scale) (util.field_errors).

Two oracle builds are used on purpose:
  * the correctly rounded build (every fp32 intrinsic = fp64 evaluation rounded once) -- the device evaluates its
    fp32 intrinsics the same way, so this comparison isolates LOGIC: it must hold on every element;
  * the host-libm build (what a gfortran build of the reference links) -- intrinsics differ by <= 1 ulp and the
    reference's own thresholds (|dT_leaf| > 0.1 K, rtsoil limiter, ...) can flip for a handful of tiles, so the
    criterion is the fraction of elements inside the tolerance.
"""
import ctypes as C

import numpy as np
import pytest

from cable_b200 import lib, synth
from cable_b200.cbm import CableB200, cbm as cbm_dropin, derived_types, release
from cable_b200.registry import FIELDS, ROLE, FLAG
from oracle.pyoracle import Oracle
from util import DELS, RTOL_F32, RTOL_F64, compare_tiles, make_case, output_fields, water_balance, energy_balances

pytestmark = pytest.mark.gpu


def run_pair(cfg, grid, T_ref, forcing, nsteps, cr_math=True, dels=DELS, device_cfg=None):
    """Step oracle (in T_ref) and device (returns T_gpu) side by side through the drop-in call."""
    cfg.output_level = 2
    T_gpu = {k: v.copy() for k, v in T_ref.items()}
    o = Oracle(T_ref, cfg, cr_math=cr_math)
    with CableB200(grid.mp, device_cfg or cfg) as h:
        h.bind(T_gpu); h.upload_params(); h.upload_state()
        for k in range(nsteps):
            forcing.fill(T_ref, k)
            for n in synth.FORCING_FIELDS:
                T_gpu[n][...] = T_ref[n]
            o.cbm(k + 1, dels)
            h.cbm(k + 1, dels)
        ctr = h.counters()
    return T_gpu, o, ctr


def assert_logic_parity(T_ref, T_gpu):
    res = compare_tiles(T_ref, T_gpu)
    bad = {n: r for n, r in res.items() if r[0] > r[1]}
    assert not bad, f"fields outside tolerance vs correctly rounded oracle: {bad}"
    return max(r[0] for r in res.values())


@pytest.mark.parametrize("gs", [lib.C.c_int(0).value, 1])
def test_every_field_matches_cr_oracle(gs):
    """All 176 state/diagnostic fields, 16 steps, 5 000 tiles incl. lakes, glaciers, snow, frozen soil."""
    cfg = lib.default_cfg(); cfg.gs_switch = gs
    cfg, grid, T, F = make_case(1000, cfg=cfg)
    T_gpu, o, ctr = run_pair(cfg, grid, T, F, 16)
    worst = assert_logic_parity(T, T_gpu)
    assert worst < 1e-5
    # per step: kernel A's CBL_FASTDIV build, the ordinary kernel A over the blocks it handed back (normally none), kernel B
    assert ctr.kernel_launches == 48 and ctr.n_dryleaf_warn == o.warnings()
    # the synthetic case must actually exercise the branches we claim to cover
    assert (T["ssnow_isflag"] == 1).any() and (T["ssnow_snowd"] > 0).any() and (T["ssnow_wbice"] > 0).any()
    assert (T["veg_iveg"] == 16).any() and (T["soil_isoilm"] == 9).any() and (T["canopy_vlaiw"] > 0.001).any()


def test_libm_oracle_fraction_within_tolerance():
    """Against the host-libm oracle (closest to a gfortran build of the reference): >= 99 % of the elements of every
    fp32 field within 1e-4, fp64 fields (all derived from fp32 chains) >= 90 % within 1e-6 and >= 99 % within 1e-4."""
    cfg, grid, T, F = make_case(2000)
    T_gpu, o, _ = run_pair(cfg, grid, T, F, 12, cr_math=False)
    for f in output_fields():
        if f.name in ("bal_drybal", "bal_wetbal"):
            continue
        a, b = T[f.name].astype(np.float64), T_gpu[f.name].astype(np.float64)
        floor = 1e-3 * max(np.abs(a).max(), 1e-30)
        rel = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
        assert np.mean(rel <= 1e-4) >= 0.99, (f.name, float(np.mean(rel <= 1e-4)), float(rel.max()))
        if f.dtype == np.float64:
            assert np.mean(rel <= RTOL_F64) >= 0.90, (f.name, float(np.mean(rel <= RTOL_F64)))


@pytest.mark.parametrize("switches", [dict(fwsoil_switch=1), dict(fwsoil_switch=2), dict(ssnow_potev=1),
                                      dict(diag_soil_resp_on=0), dict(l_new_runoff_speed=1, l_new_reduce_soilevp=1),
                                      dict(gs_switch=1, fwsoil_switch=1),
                                      # the XSW instantiation of the kernels (cbm_kernel.cuh)
                                      dict(litter=1), dict(l_rev_corr=1), dict(litter=1, l_rev_corr=1, ssnow_potev=1),
                                      dict(soil_thermal_fix=1), dict(l_new_roughness_soil=1), dict(redistrb=1),
                                      dict(call_climate=1), dict(call_climate=1, gs_switch=1),
                                      dict(gs_switch=1, litter=1, l_rev_corr=1, soil_thermal_fix=1, l_new_roughness_soil=1,
                                           redistrb=1, call_climate=1)])
def test_switch_matrix(switches):
    """cable_user switches on the supported path."""
    cfg = lib.default_cfg()
    for k, v in switches.items():
        setattr(cfg, k, v)
    cfg, grid, T, F = make_case(300, cfg=cfg, start_doy=200)
    T_gpu, o, _ = run_pair(cfg, grid, T, F, 6)
    assert_logic_parity(T, T_gpu)


def test_new_roughness_soil_carries_us_at_output_level_1():
    """l_new_roughness_soil makes canopy%us an input of the next step (cable_roughness.F90:197): the device must keep it
    at every output level, and the initial value must come from the host array (0.1, cable_parameters.F90:1213)."""
    cfg = lib.default_cfg(); cfg.l_new_roughness_soil = 1
    cfg, grid, T, F = make_case(300, cfg=cfg, start_doy=200)
    dev_cfg = lib.default_cfg(); dev_cfg.l_new_roughness_soil = 1; dev_cfg.output_level = 1
    T_gpu = {k: v.copy() for k, v in T.items()}
    o = Oracle(T, cfg, cr_math=True)
    with CableB200(grid.mp, dev_cfg) as h:
        h.bind(T_gpu); h.upload_params(); h.upload_state()
        for k in range(6):
            F.fill(T, k)
            for n in synth.FORCING_FIELDS:
                T_gpu[n][...] = T[n]
            o.cbm(k + 1, DELS)
            h.cbm(k + 1, DELS)
    fields = [f for f in output_fields() if f.role == ROLE["STATE"] or f.star()]
    res = compare_tiles(T, T_gpu, fields)
    bad = {n: r for n, r in res.items() if r[0] > r[1]}
    assert not bad, bad


def test_fastdiv_build_is_bit_identical_and_falls_back(monkeypatch):
    """Kernel A's CBL_FASTDIV build (IEEE divisions / square roots without the slow-path scaffolding, cable_fast.cu) must
    give the same bits as the ordinary build; and a tile whose operands leave the fast path's window (here: a
    subnormal-scale CO2 concentration, so csx and the quantities derived from it are ~1e-38) must come out identical
    too, through the flag -> recompute path."""
    def run(fast, tiny_ca):
        monkeypatch.setenv("CABLE_B200_FASTDIV", str(fast))
        cfg, grid, T, F = make_case(300)
        cfg.output_level = 2
        with CableB200(grid.mp, cfg) as h:
            h.bind(T); h.upload_params(); h.upload_state()
            for k in range(4):
                F.fill(T, k)
                if tiny_ca:
                    T["met_ca"][0][::7] = np.float32(3e-38)
                h.cbm(k + 1, DELS)
            redo = h.counters().n_fastdiv_redo_blocks
        return T, redo
    for tiny in (False, True):
        (a, ra), (b, rb) = run(1, tiny), run(0, tiny)
        assert rb == 0 and (ra > 0) == tiny, (tiny, ra, rb)       # the fallback runs exactly when operands leave the window
        for f in output_fields():
            assert np.array_equal(a[f.name], b[f.name], equal_nan=True), (tiny, f.name)


def test_fastdiv_decaying_numerators_stay_on_the_fast_path(monkeypatch):
    """Quantities that decay through the subnormal range over a long run (canopy storage, wet fraction, the halving
    soil wetness factor, soil ice) reach divisions of kernel A as numerators below the fast path's 2^-100 window.  Those
    sites divide through the fp64 chain (dvw) or the operator (dvx), cbm_consts.cuh: same bits as the ordinary build and
    no block handed back."""
    def run(fast):
        monkeypatch.setenv("CABLE_B200_FASTDIV", str(fast))
        cfg, grid, T, F = make_case(300)
        cfg.output_level = 2
        T["canopy_cansto"][0][::5] = np.float32(1e-42); T["canopy_cansto"][0][1::5] = np.float32(3e-33)
        T["canopy_oldcansto"][...] = T["canopy_cansto"]                     # D10: cbm restores cansto from oldcansto
        T["ssnow_owetfac"][0][::4] = np.float32(2e-36); T["ssnow_wetfac"][0][::4] = np.float32(2e-36)
        T["ssnow_wbice"][0][::3] = 1e-310; T["ssnow_wbice"][0][1::3] = 4e-300
        with CableB200(grid.mp, cfg) as h:
            h.bind(T); h.upload_params(); h.upload_state()
            for k in range(6):
                F.fill(T, k)
                h.cbm(k + 1, DELS)
            redo = h.counters().n_fastdiv_redo_blocks
        return T, redo
    (a, ra), (b, rb) = run(1), run(0)
    assert ra == 0 and rb == 0, (ra, rb)
    for f in output_fields():
        assert np.array_equal(a[f.name], b[f.name], equal_nan=True), f.name


@pytest.mark.parametrize("mp_case", [(1, 1), (1, 3), (7, 5), (26, 5)])      # 1, 3, 35, 130 tiles: ragged vs the 128-thread block
def test_ragged_and_tiny_sizes(mp_case):
    nland, nap = mp_case
    cfg, grid, T, F = make_case(nland, nap=nap)
    T_gpu, _, _ = run_pair(cfg, grid, T, F, 4)
    assert_logic_parity(T, T_gpu)


def test_single_site_half_hourly():
    """BASELINE config 1 shape: one tile, dels = 1800 s, two days."""
    cfg = lib.default_cfg()
    grid = synth.make_grid(1, 1, site_lat=-35.66)
    T = synth.make_tiles(grid, cfg, single_pft=2)
    F = synth.Forcing(grid, T, 1800.0, start_doy=15)
    T_gpu, _, _ = run_pair(cfg, grid, T, F, 96, dels=1800.0)
    assert_logic_parity(T, T_gpu)


def test_fused_and_split_kernels_agree(monkeypatch):
    cfg, grid, T, F = make_case(400)
    Tb = {k: v.copy() for k, v in T.items()}
    monkeypatch.setenv("CABLE_B200_SPLIT", "0")
    Ta_gpu, _, ca = run_pair(cfg, grid, T, F, 5)
    monkeypatch.setenv("CABLE_B200_SPLIT", "1")
    Tb_gpu, _, cb = run_pair(cfg, grid, Tb, F, 5)
    assert ca.kernel_launches == 5 and cb.kernel_launches == 15
    for f in output_fields():
        np.testing.assert_array_equal(Ta_gpu[f.name], Tb_gpu[f.name], err_msg=f.name)


def test_resident_stepping_equals_dropin_and_is_deterministic():
    """step() from a device-resident forcing ring (the benchmark's inner loop) == cable_b200_cbm() per step;
    and two identical runs are bitwise identical."""
    cfg, grid, T0, F = make_case(500)
    cfg.n_forcing_slots = 4; cfg.output_level = 1
    nsteps = 8
    fs = []
    for k in range(nsteps):
        F.fill(T0, k); fs.append({n: T0[n].copy() for n in synth.FORCING_FIELDS})
    results = []
    for mode in ("dropin", "ring", "ring"):
        T = {k: v.copy() for k, v in T0.items()}
        with CableB200(grid.mp, cfg) as h:
            h.bind(T); h.upload_params(); h.upload_state()
            for k in range(nsteps):
                for n, a in fs[k].items():
                    T[n][...] = a
                if mode == "dropin":
                    h.cbm(k + 1, DELS)
                else:
                    h.set_forcing_async(k % 4); h.step(k + 1, DELS, k % 4); h.sync()   # host arrays reused: sync before refill
            h.download_state(); h.download_diag(star_only=True)
        results.append(T)
    for f in FIELDS:
        if f.role == ROLE["STATE"] or (f.role == ROLE["DIAG"] and f.flags & FLAG["STAR"]):
            np.testing.assert_array_equal(results[0][f.name], results[1][f.name], err_msg=f.name)
            np.testing.assert_array_equal(results[1][f.name], results[2][f.name], err_msg=f.name)


def test_tile_independence_under_permutation():
    """No tile-to-tile dependence in cbm: shuffling the tiles permutes the results and nothing else."""
    cfg, grid, T, F = make_case(300)
    rng = np.random.default_rng(5)
    perm = rng.permutation(grid.mp)
    Tp = {k: np.ascontiguousarray(v[:, perm]) for k, v in T.items()}
    outs = []
    for tiles, p in ((T, None), (Tp, perm)):
        c = lib.default_cfg(); c.output_level = 1
        with CableB200(grid.mp, c) as h:
            h.bind(tiles); h.upload_params(); h.upload_state()
            for k in range(4):
                F.fill(T, k)
                for n in synth.FORCING_FIELDS:
                    tiles[n][...] = T[n] if p is None else T[n][:, p]
                h.cbm(k + 1, DELS)
        outs.append(tiles)
    for n in ("canopy_fe", "canopy_fh", "ssnow_tgg", "ssnow_wb", "canopy_fpn", "ssnow_snowd"):
        np.testing.assert_array_equal(outs[0][n][:, perm], outs[1][n], err_msg=n)


def test_dropin_signature_mirror():
    """cable_b200.cbm.cbm(...) takes the reference argument list (cbl_model_driver_offline.F90:38-40)."""
    cfg, grid, T, F = make_case(120)
    Tg = {k: v.copy() for k, v in T.items()}
    cfg.output_level = 1
    o = Oracle(T, cfg, cr_math=True)
    ty = derived_types(Tg)
    for k in range(3):
        F.fill(T, k)
        for n in synth.FORCING_FIELDS:
            Tg[n][...] = T[n]
        o.cbm(k + 1, DELS)
        cbm_dropin(k + 1, DELS, ty.air, ty.bgc, ty.canopy, ty.met, ty.bal, ty.rad, ty.rough, ty.soil, ty.ssnow, None, ty.veg,
                   None, ty.scr.xk, ty.scr.c1, ty.scr.rhoch, cfg=cfg)
    release(ty.ssnow)
    star = [f for f in output_fields() if f.role == ROLE["STATE"] or f.flags & FLAG["STAR"]]
    res = compare_tiles(T, Tg, star)
    assert all(r[0] <= r[1] for r in res.values()), {n: r for n, r in res.items() if r[0] > r[1]}


def test_error_behaviour_on_device():
    cfg, grid, T, F = make_case(10)
    with CableB200(grid.mp, cfg) as h:
        with pytest.raises(lib.CableError) as e:           # nothing bound yet
            h.upload_params()
        assert e.value.code == -4
        h.bind(T)
        h.upload_params(); h.upload_state()
        with pytest.raises(lib.CableError):                 # slot out of range
            h.step(1, DELS, 99)
        with pytest.raises(lib.CableError):                 # dels must be positive
            h.step(1, 0.0, 0)
        bad = T["soil_swilt_vec"].copy(); bad[2, 0] += 0.01     # per-layer soil parameters are not a spread
        T2 = dict(T); T2["soil_swilt_vec"] = bad
        h.bind({"soil_swilt_vec": bad})
        with pytest.raises(lib.CableError) as e:
            h.upload_params()
        assert e.value.code == -5


def test_full_size_properties():
    """BASELINE config 3 size (62 000 land points x 5 = 310 000 tiles): the reference's closure invariants hold on the
    device output, everything stays finite, and a strided sample of tiles matches the oracle."""
    cfg = lib.default_cfg(); cfg.output_level = 2
    cfg, grid, T, F = make_case(62000, cfg=cfg, start_doy=172)
    sample = np.arange(0, grid.mp, 97)                       # ~3 200 tiles spread over the whole grid
    Ts = {k: np.ascontiguousarray(v[:, sample]) for k, v in T.items()}
    o = Oracle(Ts, cfg, cr_math=True)
    with CableB200(grid.mp, cfg) as h:
        h.bind(T); h.upload_params(); h.upload_state()
        for k in range(6):
            F.fill(T, k)
            for n in synth.FORCING_FIELDS:
                Ts[n][...] = T[n][:, sample]
            wb_prev = T["ssnow_wbtot"][0].copy()
            h.cbm(k + 1, DELS)
            o.cbm(k + 1, DELS)
            for f in output_fields():
                assert np.all(np.isfinite(T[f.name])), f.name
            if k >= 1:
                radbal, ebalsoil, ebalveg, ebal = energy_balances(T)
                assert np.abs(radbal).max() < 5e-3 and np.abs(ebalsoil).max() < 1e-3
                assert np.abs(ebalveg).max() < 5e-3 and np.abs(ebal).max() < 5e-3
                wbal = water_balance(T, DELS, wb_prev)
                normal = T["veg_iveg"][0] < 16
                assert np.abs(wbal[normal]).max() < 5e-2 and abs(wbal[normal].mean()) < 2e-3
    res = compare_tiles(Ts, {k: np.ascontiguousarray(v[:, sample]) for k, v in T.items()})
    bad = {n: r for n, r in res.items() if r[0] > r[1]}
    assert not bad, bad


def test_device_grid_reduction_matches_reference_rule():
    """patch -> grid-cell area-weighted mean (cable_grid_reductions.F90:66-73) on the device."""
    import torch
    from cable_b200.sharding import grid_cell_average
    cfg, grid, T, F = make_case(300)
    cfg.output_level = 1
    with CableB200(grid.mp, cfg, device=0) as h:
        h.bind(T); h.upload_params(); h.upload_state()
        F.fill(T, 0); h.cbm(1, DELS)
        d_pf = torch.from_numpy(grid.patchfrac).cuda(); d_cs = torch.from_numpy(grid.cstart).cuda()
        d_ce = torch.from_numpy(grid.cend).cuda(); out = torch.zeros(grid.nland, device="cuda")
        h.grid_reduce("canopy_fe", 0, d_pf.data_ptr(), d_cs.data_ptr(), d_ce.data_ptr(), grid.nland, out.data_ptr())
        h.sync()
        got = out.cpu().numpy()
    want = grid_cell_average(T["canopy_fe"][0], grid.patchfrac, grid.cstart, grid.cend)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-6)     # fmad=false on device; same order of accumulation


def test_pipelined_resident_step_is_bit_identical(monkeypatch):
    """The resident step as chunk chains on several streams with no join between steps (cable_b200_step, pipe_chunk) against
    the unpipelined launch on the same inputs: identical bits in every state / driver-visible field, in the driver's
    accumulators and in every step's grid-cell output block.  The offline-driver loop rides the pipeline (post-step statements
    and the output reduction run per chunk on the chain streams; a land point that straddles two chunks couples neighbours);
    some steps go without post_step / output so that both the joined and the unjoined orders are exercised, and a ragged grid
    (1..5 patches per land point) puts land points across the chunk edges."""
    from util import ragged_case
    rows = [("canopy_fe", 0, "mean"), ("ssnow_tgg", 2, "mean"), ("ssnow_wb", 0, "mean"), ("ssnow_runoff", 0, "mean"),
            ("bal_wbal", 0, "mean"), ("canopy_fnee", 0, "mean")]

    def run(streams):
        monkeypatch.setenv("CABLE_B200_PIPE_STREAMS", str(streams))
        monkeypatch.setenv("CABLE_B200_PIPE_CHUNK", "20480")          # ~5 chunks (small-range kernels), edges inside land points
        cfg = lib.default_cfg(); cfg.n_forcing_slots = 4; cfg.output_level = 1
        cfg, grid, T, F, idx = ragged_case(32000, cfg=cfg, start_doy=100)
        assert grid.mp > 4 * 20480 and np.any((grid.cstart < 20480) & (grid.cend >= 20480))
        conv = lib.MetConvert(tair_offset=0.0, psurf_scale=0.01, rainf_scale=DELS, co2_scale=1.0e-6, snowf_from_tair=1)
        slices = [np.ascontiguousarray(F.land_slice(k), np.float32) for k in range(10)]
        outs = []
        with CableB200(grid.mp, cfg) as h:
            h.bind(T); h.upload_params(); h.upload_state()
            h.driver_init(grid.cstart, grid.cend, grid.patchfrac, grid.lat[grid.tile2land])
            h.output_plan(rows)
            T["veg_vlai"][0] = F.lai(0)[idx]; h.upload_lai()
            for k in range(10):
                h.set_met_async(k % 4, slices[k], conv)      # the ring: a slot is rewritten while earlier steps are still in flight
                h.step(k + 1, DELS, k % 4)
                if k not in (0, 5):
                    h.post_step(k + 1, 1, DELS)
                if k not in (1, 5, 6):
                    out = np.zeros((len(rows), grid.nland), np.float32)
                    h.output_fetch_async(out); outs.append(out)
                    if k % 2:
                        h.output_wait()
            h.output_wait()
            h.download_state(); h.download_diag()
            acc = {n: h.driver_download(n) for n in ("sum_flux_sumpn", "bal_wbal", "bal_ebal")}
            launches = h.counters().kernel_launches
        return T, acc, launches, outs
    Ta, acca, la, oa = run(0)
    Tb, accb, lb, ob = run(4)
    assert lb > la                                            # chunks x (A fast, A, B) per step against 2 chains
    for f in output_fields():
        assert np.array_equal(Ta[f.name], Tb[f.name], equal_nan=True), f.name
    for n in acca:
        assert np.array_equal(acca[n], accb[n], equal_nan=True), n
    assert len(oa) == len(ob) == 7
    for k, (x, y) in enumerate(zip(oa, ob)):
        assert np.array_equal(x, y, equal_nan=True) and np.isfinite(x).all() and np.abs(x).max() > 0, k
