#!/usr/bin/env python
"""bench.py -- tile-timesteps/sec of cbm() on B200, with roofline, end-to-end and CPU-baseline figures.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2]): global 0.5 degree GSWP3-shape synthetic forcing, ONE grid of 62 000 land
points x 5 tiles = 310 000 tiles, 3-hourly (dels = 10 800 s), sharded over the N GPUs as contiguous land-point
blocks (rule of master_decomp, src/offline/cable_mpimaster.F90:1428-1463) -- strong scaling: 310 k / 155 k / 78 k /
39 k tiles per GPU at N = 1 / 2 / 4 / 8.  `--nland 250000` is configs[3] (1.25 M tiles, 156 k per GPU at N = 8).
`--scaling weak` gives every rank its own --nland-point grid instead.  A "step" is one cbm() pass over every tile
of the grid.  There is no data-path collective: NCCL only gathers grid-cell-reduced diagnostics to rank 0 once
per output interval (= once per timed region).

value : device-resident rate -- a year-shaped forcing ring already in HBM, one fused kernel launch per step.
e2e   : the reference-facing call cable_b200_cbm() with HOST buffers -- per step the forcing goes H2D from
        pinned host arrays and prognostic state + driver-visible diagnostics come back D2H.
roofline : algorithmic bytes (648 B / tile-step) x tiles per launch / mean kernel time
        (CUDA events on the launching stream, recorded inside the library), against the measured HBM peak.
cpu_baseline : the C++ oracle (restatement of the reference, glibc libm) on all host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

DELS = 10800.0
ALGO_BYTES_PER_TILE_STEP = 648          # forcing 68 + state 252 r + 252 w + per-tile params 76
NLAND = 62000                           # configs[2]: the 0.5 degree global grid
NAP = 5
RING = 8                                # forcing ring = one model day of 3-hourly steps
# rows the output module writes by default (src/offline/cable_diagnostics.F90 registrations): (field, component, method)
OUTPUT_ROWS = (
    [("canopy_fe", 0, "mean"), ("canopy_fh", 0, "mean"), ("canopy_ga", 0, "mean"), ("rad_rnet", 0, "mean"),
     ("rad_swnet", 0, "mean"), ("rad_lwnet", 0, "mean"), ("canopy_fev", 0, "mean"), ("canopy_fevc", 0, "mean"),
     ("canopy_fevw", 0, "mean"), ("canopy_fes", 0, "mean"), ("canopy_fhv", 0, "mean"), ("canopy_fhs", 0, "mean"),
     ("canopy_fnee", 0, "mean"), ("canopy_fpn", 0, "mean"), ("canopy_frday", 0, "mean"), ("canopy_frp", 0, "mean"),
     ("canopy_frs", 0, "mean"), ("canopy_fgpp", 0, "mean"), ("canopy_fnpp", 0, "mean"), ("ssnow_runoff", 0, "mean"),
     ("ssnow_rnof1", 0, "mean"), ("ssnow_rnof2", 0, "mean"), ("ssnow_smelt", 0, "mean"), ("ssnow_snowd", 0, "mean"),
     ("ssnow_totsdepth", 0, "mean"), ("rad_albedo", 0, "mean"), ("rad_albedo", 1, "mean"), ("rad_trad", 0, "mean"),
     ("canopy_tv", 0, "mean"), ("canopy_tscrn", 0, "mean"), ("canopy_qscrn", 0, "mean"), ("canopy_cansto", 0, "mean"),
     ("canopy_through", 0, "mean"), ("canopy_epot", 0, "mean"), ("ssnow_tss", 0, "mean"), ("canopy_fwsoil", 0, "mean"),
     ("bal_wbal", 0, "mean"), ("bal_ebal", 0, "mean")]
    + [("ssnow_tgg", k, "mean") for k in range(6)] + [("ssnow_wb", k, "mean") for k in range(6)])


# what an UNCHANGED serialdrv reads back after CALL cbm (cable_b200_set_output_mask): the fields its output module registers
# (OUTPUT_ROWS) + the inputs of mass_balance / energy_balance (cable_checks.F90:472-618) + runoff scaling (cable_serial.F90:602-605)
DRIVER_READS = sorted({r[0] for r in OUTPUT_ROWS if not r[0].startswith("bal_")} | {
    "ssnow_smelt", "ssnow_rnof1", "ssnow_rnof2", "ssnow_runoff", "canopy_tscrn", "canopy_fpn", "canopy_frday", "canopy_frp",
    "canopy_frpw", "canopy_frpr", "canopy_frs", "canopy_fnee", "canopy_delwc", "ssnow_snowd", "ssnow_osnowd", "canopy_fevw",
    "canopy_fev", "ssnow_cls", "air_rlam", "rad_albedo", "rad_transd", "ssnow_otss", "canopy_tv", "canopy_fnv", "canopy_fns",
    "canopy_fhs", "canopy_ga", "canopy_fhv", "canopy_fh", "rad_qcan", "rad_qssabs", "rad_flws", "ssnow_wbtot", "canopy_fevc",
    "canopy_fes", "canopy_cansto"})


def measured_peak_hbm() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def profiled_metrics() -> dict | None:
    """Per-launch DRAM traffic and pipe utilisation of the two step kernels from the committed `ncu --set full`
    capture (profiles/, written by tools/ncu_metrics_json.py) -- attached to the roofline object, never measured here."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_metrics.json")))
    if not files:
        return None
    try:
        with open(files[-1]) as fh:
            m = json.load(fh)
        ks = m["kernels"]
        return {"file": os.path.relpath(files[-1], ROOT), "tiles": m.get("tiles"),
                "flops_per_tile_step": m.get("flops_per_tile_step"),
                "traffic_bytes_per_step": sum(k["dram_bytes"] for k in ks),
                "warp_inst_per_step": sum(k.get("warp_inst_executed") or 0.0 for k in ks),
                "kernels": [{k2: k[k2] for k2 in ("kernel", "duration_ms", "dram_bytes", "issue_active_pct", "pipe_fp64_pct",
                                                    "pipe_fma_fp32_pct", "pipe_alu_pct", "pipe_xu_pct", "dram_pct_of_peak",
                                                    "active_threads_per_warp_inst", "icache_hit_pct")} for k in ks]}
    except Exception:
        return None


def pipe_roof(prof: dict | None, tile_steps_per_s_per_gpu: float, clocks: dict | None) -> dict | None:
    """Second candidate bound: executed fp64 / fp32 flops per tile-step, COUNTED by ncu in the committed
    capture, x this run's measured throughput, against the non-tensor pipe peaks of one B200 (148 SMs x 64 DFMA or
    128 FFMA lanes per cycle -- ncu's sm__sass_thread_inst_executed_op_{dfma,ffma}_pred_on.peak_sustained -- x 2 flops
    x the SM clock sampled during the timed region)."""
    f = (prof or {}).get("flops_per_tile_step")
    if not f:
        return None
    mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
    peak64, peak32 = 148 * 64 * 2 * mhz * 1e6 / 1e12, 148 * 128 * 2 * mhz * 1e6 / 1e12
    a64, a32 = f["fp64"] * tile_steps_per_s_per_gpu / 1e12, f["fp32"] * tile_steps_per_s_per_gpu / 1e12
    return {"flops_per_tile_step_fp64": f["fp64"], "flops_per_tile_step_fp32": f["fp32"], "counted": f.get("how"),
            "fp64": {"achieved": a64, "peak": peak64, "unit": "TFLOP/s", "frac": a64 / peak64},
            "fp32": {"achieved": a32, "peak": peak32, "unit": "TFLOP/s", "frac": a32 / peak32}}


def issue_roof(prof: dict | None, mp: int, kern_ms: float, clocks: dict | None) -> dict | None:
    """The bound that actually bites: warp-instruction issue slots.  Executed warp instructions per
    tile-step COUNTED by ncu in the committed capture (smsp__inst_executed.sum over the step's launches / tiles of that
    capture) x this run's tiles per step / this run's event-timed step, against 148 SMs x 4 schedulers x 1 warp
    instruction per cycle at the SM clock sampled during the timed region."""
    if not prof or not prof.get("warp_inst_per_step") or not prof.get("tiles") or kern_ms <= 0:
        return None
    mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
    per_tile = prof["warp_inst_per_step"] / prof["tiles"]
    peak = 148 * 4 * mhz * 1e6
    achieved = per_tile * mp / (kern_ms * 1e-3)
    return {"warp_inst_per_tile_step": per_tile, "warp_inst_per_step": per_tile * mp, "achieved": achieved, "peak": peak,
            "unit": "warp-inst/s", "frac": achieved / peak, "sm_mhz": mhz,
            "floor_ms_at_this_instruction_count": per_tile * mp / peak * 1e3,
            "counted": f"ncu smsp__inst_executed.sum, {prof.get('file')}"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self._halt.is_set():
                    break
        except Exception:
            pass

    def stop(self) -> dict:
        self._halt.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_text(nland_total: int, tiles_total: int) -> str:
    res = "0.5deg" if nland_total <= 100000 else "0.25deg"
    return (f"global {res} GSWP3-shape synthetic forcing: ONE grid of {nland_total} land points x {NAP} tiles = {tiles_total} "
            f"tiles, dels={int(DELS)}s, leuning/standard/HDM/icycle=0 (cable.nml)")


# ------------------------------------------------------------------------------------------------------------------
def cpu_oracle_rate(nland: int, nsteps: int, nproc: int, warm: int = 1) -> tuple[float, float, int]:
    """Oracle throughput on `nproc` host processes stepping ONE `nland`-point grid (the b200 arm's grid: same seed,
    same forcing), split into contiguous land-point blocks exactly like the reference MPI driver does it (master_decomp,
    cable_mpimaster.F90:1428-1445), one worker per core.  -> (tile-steps/s, seconds, tiles per step)."""
    import multiprocessing as mp_
    ctx = mp_.get_context("fork")
    with ctx.Pool(nproc) as pool:
        res = pool.starmap(_cpu_worker, [(r, nproc, nland, nsteps, warm) for r in range(nproc)])
    tmax = max(t for t, _ in res)
    tiles = sum(n for _, n in res)
    return tiles * nsteps / tmax, tmax, tiles


def _cpu_worker(rank: int, nproc: int, nland: int, nsteps: int, warm: int):
    from cable_b200 import lib, synth
    from cable_b200.sharding import shard_grid
    from cable_b200.partition import array_partition, land_to_tile_range
    from oracle.pyoracle import Oracle
    cfg = lib.default_cfg()
    full = synth.make_grid(nland, NAP, seed=synth.SEED)
    tiles_full = synth.make_tiles(full, cfg)
    forcing = synth.Forcing(full, tiles_full, DELS, start_doy=172)
    l0, nl = array_partition(full.nland, nproc, rank)
    t0, t1 = land_to_tile_range(full.cstart, full.cend, l0, nl)
    sets = []
    for k in range(RING):
        forcing.fill(tiles_full, k)
        sets.append({n: np.ascontiguousarray(tiles_full[n][:, t0:t1]) for n in synth.FORCING_FIELDS})
    grid, tiles = shard_grid(full, tiles_full, rank, nproc)
    del tiles_full
    o = Oracle(tiles, cfg, cr_math=False)
    for k in range(warm):
        for n, a in sets[k % RING].items():
            tiles[n][...] = a
        o.cbm(k + 1, DELS)
    t0 = time.perf_counter()
    for k in range(warm, warm + nsteps):
        for n, a in sets[k % RING].items():
            tiles[n][...] = a
        o.cbm(k + 1, DELS)
    return time.perf_counter() - t0, grid.mp


def run_reference(args) -> None:
    """--impl reference: the reference's CPU implementation of the path.  The Fortran cannot be built here
    (no Fortran compiler), so this is the oracle port on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    # the b200 arm's own grid (same seed, same forcing ring, every tile of it each step); a bounded number of steps
    nland_total = args.nland * (args.gpus if args.scaling == "weak" else 1)
    rate, tmax, tiles_step = cpu_oracle_rate(nland_total, args.steps, cores, warm=args.warmup)
    ms = tmax / max(args.steps, 1) * 1e3
    line = {
        "impl": "reference", "metric": "tile-timesteps/sec for cbm()", "value": rate, "unit": "tile-timesteps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        # the b200 arm's workload: same grid, same seed, every tile of it each step
        "config": {"workload": workload_text(nland_total, nland_total * NAP),
                   "global_tiles": nland_total * NAP, "tiles_per_step": tiles_step,
                   "parallelism": f"land-point blocks x{cores} host processes (master_decomp rule)",
                   "sample": f"every step = all {tiles_step} tiles of that grid; {args.steps} steps (bounded), {args.warmup} warm-up"},
        "cpu_baseline": {"value": rate, "unit": "tile-timesteps/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} workers x {tiles_step // cores}+ tiles ({nland_total} land points split by "
                                   f"master_decomp) x {args.steps} steps, C++ restatement of the reference (not the Fortran "
                                   "binary), forcing in memory"},
        "e2e": {"value": rate, "unit": "tile-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist
    from cable_b200 import lib, synth
    from cable_b200.cbm import CableB200
    from cable_b200.registry import FIELDS, ROLE, FLAG

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cable_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"       # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, W = args.steps, max(args.warmup, 3)

    # ---- this rank's shard: a contiguous block of land points (no halo, no inter-GPU traffic in a step)
    from cable_b200.partition import array_partition, land_to_tile_range
    from cable_b200.sharding import shard_grid, shard_grid_points, interleaved_land_points
    cfg = lib.default_cfg()
    cfg.n_forcing_slots = RING
    cfg.output_level = 1
    strong = args.scaling == "strong"
    if strong:
        # ONE grid (seed = SEED) cut by the reference's decomposition rule; every rank generates the whole synthetic grid
        # and forcing on its host (cheap) and keeps its block -- the master's scatter is outside the timed region anyway
        full = synth.make_grid(args.nland, NAP, seed=synth.SEED)
        tiles_full = synth.make_tiles(full, cfg)
        if args.decomp == "block" or world == 1:      # the reference's rule: contiguous blocks, sizes differ by <= 1
            l0, nland = array_partition(full.nland, world, rank)
            land_idx = np.arange(l0, l0 + nland)
        else:                                         # chunks of 64 land points dealt round-robin (load balance)
            land_idx = interleaved_land_points(full.nland, world, rank)
            nland = int(land_idx.size)
        grid, tiles, tile_idx = shard_grid_points(full, tiles_full, land_idx)
        nland_total, mp_total = full.nland, full.mp
    else:
        full = synth.make_grid(args.nland, NAP, seed=synth.SEED + rank)
        tiles_full = synth.make_tiles(full, cfg)
        nland = full.nland
        land_idx, tile_idx = np.arange(full.nland), np.arange(full.mp)
        grid, tiles = full, tiles_full
        nland_total, mp_total = full.nland * world, full.mp * world
    mp = grid.mp
    forcing = synth.Forcing(full, tiles_full, DELS, start_doy=172)      # generated on the whole grid, sliced per rank

    # pinned host buffers for everything that moves per step (forcing ring, state, STAR diagnostics)
    def pinned_like(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    keep = []
    for f in FIELDS:
        if f.flags & FLAG["HOSTONLY"]:
            continue
        if f.role == ROLE["STATE"] or (f.role == ROLE["DIAG"] and f.flags & FLAG["STAR"]):
            t = pinned_like(tiles[f.name]); keep.append(t); tiles[f.name] = t.numpy()
    fsets = []
    for k in range(RING):
        forcing.fill(tiles_full, k)
        s = {}
        for n in synth.FORCING_FIELDS:
            t = pinned_like(np.ascontiguousarray(tiles_full[n][:, tile_idx])); keep.append(t); s[n] = t.numpy()
        fsets.append(s)
    for n in synth.FORCING_FIELDS:
        tiles[n] = fsets[0][n].copy()
    if strong:
        del tiles_full
    state0 = {f.name: tiles[f.name].copy() for f in FIELDS if f.role == ROLE["STATE"]}

    h = CableB200(mp, cfg, device=local)
    h.bind(tiles)
    h.upload_params()
    h.upload_state()
    # forcing ring resident in HBM
    for k in range(RING):
        h.bind(fsets[k]); h.set_forcing_async(k)
    h.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # output-interval diagnostics: device-side patch -> grid-cell reduction of the output module's rows, gathered to rank 0
    # INSIDE the library (ncclSend / ncclRecv of uneven land-point blocks, cable_b200_output_gather_async); torch only
    # carries the 128-byte NCCL id from rank 0 to the others, as the Fortran driver's MPI_Bcast would
    if not strong:
        counts = np.asarray([nland] * world, np.int32)
    elif args.decomp == "block" or world == 1:
        counts = np.asarray([array_partition(nland_total, world, r)[1] for r in range(world)], np.int32)
    else:
        counts = np.asarray([interleaved_land_points(nland_total, world, r).size for r in range(world)], np.int32)
    def new_comm_id():
        """a fresh NCCL id per communicator (one per handle), created on rank 0 and broadcast"""
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(CableB200.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, src=0)
        return bytes(idt.cpu().numpy().tobytes())
    gathered_host = torch.zeros((len(OUTPUT_ROWS), int(counts.sum())), dtype=torch.float32, pin_memory=True) if rank == 0 else None

    def attach_driver(hh):
        hh.driver_init(grid.cstart, grid.cend, grid.patchfrac, grid.lat[grid.tile2land])
        hh.output_plan(OUTPUT_ROWS)
        if world > 1:
            hh.comm_init(new_comm_id(), rank, world)

    def gather_diags():
        h.output_gather_async(0, gathered_host.numpy() if rank == 0 else None, counts)
        h.output_wait()
        return [gathered_host] if rank == 0 else None

    attach_driver(h)

    # ---- (1) device-resident rate ----------------------------------------------------------------------------
    for k in range(W):
        h.step(k + 1, DELS, k % RING)
    h.sync()
    gather_diags()                       # warm-up: NCCL communicator set-up must not land in the timed region
    h.reset_counters()
    h.profile(True)
    sampler = ClockSampler(local); sampler.start()
    time.sleep(0.25)
    barrier()
    # Timing rule: inputs larger than L2, or L2 flushed between timed iterations.  A rank's resident working set
    # (parameters + state + diagnostics + the 8-slot forcing ring ~ 1.7 KB per tile) exceeds the 126 MB L2 only for
    # shards above ~150 k tiles; below 2 x L2 every timed step is preceded by a 256 MB memset and the steps are timed
    # one by one with the library's CUDA events on its compute stream (the flush is outside the events), summed.
    ws_bytes = mp * 1700
    flush = ws_bytes < 2 * 126e6 and not args.no_flush
    fl = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush else None
    t0 = time.perf_counter()
    for k in range(W, W + K):
        if flush:
            h.sync(); fl.zero_(); torch.cuda.synchronize()
        h.step(k + 1, DELS, k % RING)
    h.sync()
    tg0 = time.perf_counter()
    gathered = gather_diags()
    barrier()
    t_gather = time.perf_counter() - tg0
    t_wall = time.perf_counter() - t0
    clocks = sampler.stop()
    ctr = h.counters()
    h.profile(False)
    kern_ms = ctr.kernel_ms / max(ctr.kernel_ms_count, 1)
    # flushed: sum of the K event-timed steps + the gather; otherwise the wall clock between the two barriers
    t_res = max_over_ranks(ctr.kernel_ms * 1e-3 + t_gather if flush else t_wall)
    t_wall = max_over_ranks(t_wall)
    per_rank_ms = [kern_ms]
    if world > 1:
        tt = torch.tensor([kern_ms], device="cuda", dtype=torch.float64)
        allk = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allk, tt)
        per_rank_ms = [float(x.item()) for x in allk]
    launches = int(ctr.kernel_launches)
    value = mp_total * K / t_res
    finite = bool(torch.isfinite(gathered[0]).all().item()) if rank == 0 else True

    # ---- (2) end-to-end through the C ABI with HOST buffers: the offline driver's time loop --------------------------
    # Per step, exactly what cable_serial does around CALL cbm (cable_serial.F90:566-746), every stage through the
    # library: one time slice of per-land-point met (pinned host memory, as read from the met file) goes H2D and is
    # expanded to the tiles (get_met_data), cbm runs, the post-step statements run (runoff*dels, sumcflux, mass and
    # energy balance), and the output module's rows are reduced patch -> grid cell and come back D2H EVERY step
    # (output%averaging='all', output%patch=.FALSE.: cable.nml defaults).  The D2H of step k-1 overlaps step k.
    def fresh_handle():
        """A new handle on the initial state: the first-call initialisation of soil_snow (gammzz, cbl_soilsnow_main.F90:92-96)
        belongs to a handle's first step, so a run that starts over starts on a new handle, like a new process would."""
        nonlocal h
        h.close()
        for n, a in state0.items():
            tiles[n][...] = a
        h = CableB200(mp, cfg, device=local)
        h.bind(tiles); h.upload_params(); h.upload_state()

    fresh_handle()
    Ke = max(3, min(K, args.e2e_steps))
    h.driver_init(grid.cstart, grid.cend, grid.patchfrac, grid.lat[grid.tile2land])
    if world > 1:
        h.comm_init(new_comm_id(), rank, world)
    conv = lib.MetConvert(tair_offset=0.0, psurf_scale=0.01, rainf_scale=DELS, co2_scale=1.0e-6, snowf_from_tair=1)
    slices = []
    for k in range(RING):
        t = torch.empty((len(lib.MET_ROWS), nland), dtype=torch.float32, pin_memory=True)
        t.numpy()[...] = forcing.land_slice(k)[:, land_idx]; keep.append(t); slices.append(t.numpy())
    rows = OUTPUT_ROWS
    h.output_plan(rows)
    # every step's output block goes to rank 0 (the reference master receives every worker's fields every step,
    # cable_mpimaster.F90:8066-8072): [rows, all land points] on rank 0, nothing on the others
    nout = int(counts.sum()) if world > 1 else nland
    outs = [torch.zeros((len(rows), nout), dtype=torch.float32, pin_memory=True) for _ in range(2)] if (rank == 0 or world == 1) else [None, None]
    tiles["veg_vlai"][0] = forcing.lai(0)[tile_idx]
    h.upload_lai()
    checksum = 0.0

    def driver_step(k):
        nonlocal checksum
        h.set_met_async(k % RING, slices[k % RING], conv)       # H2D met slice + tile expansion + sinbet
        h.step(k + 1, DELS, k % RING)                           # cbm
        h.post_step(k + 1, 1, DELS)                             # runoff*dels, sumcflux, mass/energy balance
        h.output_wait()                                         # step k-1's output block is now on the host
        if k > 0 and outs[0] is not None:
            checksum += float(outs[(k - 1) % 2].numpy()[0, ::997].sum())     # the host consumes it
        if world > 1:                                           # reduce -> NCCL gather to rank 0 -> D2H there
            h.output_gather_async(0, outs[k % 2].numpy() if rank == 0 else None, counts)
        else:
            h.output_fetch_async(outs[k % 2].numpy())           # reduce -> D2H of this step's rows

    for k in range(3):
        driver_step(k)
    h.sync()
    h.reset_counters()
    barrier()
    t0 = time.perf_counter()
    for k in range(3, 3 + Ke):
        driver_step(k)
    h.output_wait()
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    ce = h.counters()
    e2e = mp_total * Ke / t_e2e
    h2d_step, d2h_step = ce.h2d_bytes / Ke, ce.d2h_bytes / Ke
    e2e_launches = int(ce.kernel_launches)
    last = outs[(3 + Ke - 1) % 2].numpy() if outs[0] is not None else np.zeros((len(rows), 1), np.float32)
    bad_rows = [f"{rows[r][0]}[{rows[r][1]}]" for r in range(len(rows)) if not np.isfinite(last[r]).all()]
    e2e_finite = (not bad_rows) and bool(np.isfinite(checksum))
    if bad_rows:
        print("non-finite output rows:", bad_rows, file=sys.stderr)

    # ---- (2b) the unchanged-caller drop-in: cable_b200_cbm() per step, host arrays in, host arrays out.  First with the
    # output mask an unchanged serialdrv needs (what its output module and balance checks read), then mirroring every
    # prognostic + driver-visible array (the most conservative mode)
    def dropin_leg(mask):
        fresh_handle()
        if mask:
            h.set_output_mask(mask)
        Km = max(3, min(Ke, 12))
        # warm-up: one call per forcing buffer of the ring -- the library keys its CUDA graph of the pipelined call on the bound
        # host pointers (a Fortran caller has ONE set of met arrays, hence one graph; this loop rotates RING sets)
        for k in range(RING):
            h.bind(fsets[k % RING]); h.cbm(k + 1, DELS)
        h.reset_counters()
        barrier()
        t0 = time.perf_counter()
        for k in range(RING, RING + Km):
            h.bind(fsets[k % RING])          # the driver fills met%* for this step (buffers already pinned)
            h.cbm(k + 1, DELS)               # H2D forcing + kernel + D2H of the selection + sync
        barrier()
        t_mir = max_over_ranks(time.perf_counter() - t0)
        cm = h.counters()
        return {"value": mp_total * Km / t_mir, "unit": "tile-timesteps/s", "h2d_bytes_per_step": cm.h2d_bytes / Km,
                "d2h_bytes_per_step": cm.d2h_bytes / Km, "steps": Km}
    mirror = dropin_leg(DRIVER_READS)
    mirror["api"] = (f"cable_b200_cbm + cable_b200_set_output_mask({len(DRIVER_READS)} fields: what cable_diagnostics registers and "
                     "mass_balance / energy_balance read); the rest of the state stays on the device")
    mirror_all = dropin_leg(None)
    mirror_all["api"] = "cable_b200_cbm, output_level=1, no mask: every prognostic + driver-visible array mirrored to the host each step"

    # ---- (2c) BASELINE config 5: the same grid with CASA-CNP biogeochemistry (icycle = 3, dynamic allocation): per step
    # cbm -> bgcdriver (daily accumulation; biogeochem for every tile at the end of each model day = every 8th step) ->
    # sumcflux / balances, all device-resident (cable_serial.F90:594-715)
    casa_leg = None
    if not args.no_casa:
        from cable_b200 import casa as casa_mod
        cfg5 = lib.default_cfg(); cfg5.n_forcing_slots = RING; cfg5.output_level = 1; cfg5.icycle = 3
        ccfg = casa_mod.default_cfg(); ccfg.icycle = 3; ccfg.lalloc = 1
        h.close()
        for n, a in state0.items():
            tiles[n][...] = a
        h = CableB200(mp, cfg5, device=local)
        h.bind(tiles); h.upload_params(); h.upload_state()
        for k in range(RING):
            h.bind(fsets[k]); h.set_forcing_async(k)
        h.driver_init(grid.cstart, grid.cend, grid.patchfrac, grid.lat[grid.tile2land])
        cs = casa_mod.Casa(h, ccfg)
        A = casa_mod.synth_casa(grid, tiles, ccfg, seed=31)
        silt, clay = casa_mod.soil_texture(tiles)
        cs.bind(A, silt, clay); cs.upload()
        ktauday = int(86400 / DELS)
        def casa_step(k):
            h.step(k + 1, DELS, k % RING)
            cs.bgcdriver(k + 1, 1, 1 << 30, DELS, ktauday, 172 + k // ktauday)
            h.post_step(k + 1, 1, DELS)
        for k in range(ktauday):
            casa_step(k)
        h.sync(); h.reset_counters(); barrier()
        Kc = 2 * ktauday
        t0 = time.perf_counter()
        for k in range(ktauday, ktauday + Kc):
            casa_step(k)
        h.sync(); barrier()
        t_c = max_over_ranks(time.perf_counter() - t0)
        cc = h.counters()
        cs.download()
        npp = A["casaflux_cnpp"][0]
        casa_leg = {"value": mp_total * Kc / t_c, "unit": "tile-timesteps/s", "ms_per_step": t_c / Kc * 1e3, "steps": Kc,
                    "gpu_launches": int(cc.kernel_launches), "biogeochem_calls": Kc // ktauday,
                    "outputs_finite": bool(np.isfinite(npp).all() and np.isfinite(A["casapool_cplant"]).all()),
                    "workload": "BASELINE config 5 on the same grid: cbm + bgcdriver (icycle=3 C+N+P, LALLOC=1; biogeochem for every "
                                "tile once per model day = every 8th step) + sumcflux/balances, device-resident; groundwater "
                                "(gw_model) is rejected by the reference snapshot itself (cable_canopy.F90:463-464)"}

    # ---- (3) CPU baseline on this box's host cores (rank 0, N=1 only) -----------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        nst = max(8, min(160, int(15.0 * 3.0e6 / mp_total)))      # ~10-20 s of CPU work on the same grid
        rate, tmax, tiles_step = cpu_oracle_rate(nland_total, nst, cores)
        cpu = {"value": rate, "unit": "tile-timesteps/s", "cores": cores, "kind": "port",
               "sample": f"the same {nland_total}-point grid ({tiles_step} tiles per step) split over {cores} workers by "
                         f"master_decomp x {nst} steps ({tmax:.1f} s), C++ restatement of the reference (not the Fortran "
                         "binary), forcing in memory"}

    if rank == 0:
        peak, which = measured_peak_hbm()
        achieved = ALGO_BYTES_PER_TILE_STEP * mp / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
        prof = profiled_metrics()
        traffic = None
        if prof and prof.get("tiles") == mp:
            traffic = prof["traffic_bytes_per_step"]
        line = {
            "metric": "tile-timesteps/sec for cbm()", "value": value, "unit": "tile-timesteps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": t_res / K * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": workload_text(nland_total, mp_total) if strong else
                                   f"weak scaling: every GPU its own grid of {nland} land points x {NAP} tiles = {mp} tiles, "
                                   f"dels={int(DELS)}s, leuning/standard/HDM/icycle=0 (cable.nml)",
                       "tiles_per_gpu": mp, "global_tiles": mp_total, "land_points_per_gpu": nland,
                       "parallelism": (f"contiguous land-point blocks x{world} (master_decomp rule, cable_mpimaster.F90:1428-1463)"
                                       if (args.decomp == "block" or world == 1 or not strong) else
                                       f"chunks of 64 land points dealt round-robin to {world} GPUs (load balance; tiles of a land point "
                                       "stay together; --decomp block = the reference's contiguous master_decomp blocks)"),
                       "decomp": args.decomp if (strong and world > 1) else "block",
                              "l2": (f"resident set of a rank ~ {ws_bytes / 1e6:.0f} MB (1.7 KB/tile incl. the 8-slot forcing ring) "
                              + ("< 2 x the 126 MB L2: L2 flushed (256 MB memset) before every timed step, steps timed one by "
                                 "one with CUDA events on the library's compute stream and summed (+ the gather)"
                                 if flush else "> 2 x the 126 MB L2: inputs larger than L2, no flush; wall clock between barriers")),
                       "timing": "events+flush" if flush else "wall",
                       "kernel_ms_per_rank": per_rank_ms,
                       "wall_ms_per_step_incl_flush": t_wall / K * 1e3,
                       "forcing_ring_steps": RING, "outputs_finite": finite},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": which, "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_tile_step": ALGO_BYTES_PER_TILE_STEP,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_TILE_STEP * mp,
                         "launch": "one step over all tiles = chunk chains on four streams, no join between steps (a chunk = one round of "
                                   "148 x 640 tiles; 310 000 tiles = 3 rounds + a 25 840-tile remainder): per chunk kernel A (surface+canopy, "
                                   "CBL_FASTDIV build, then the ordinary build over the blocks it handed back -- normally none) -> kernel B "
                                   "(soil/snow/carbon); kernel_ms = the timed region's first launch to the last chain's end, CUDA events on "
                                   "the library's streams, / steps",
                         "ncu": prof,
                         "pipe": pipe_roof(prof, value / world, clocks),
                         "issue": issue_roof(prof, mp, kern_ms, clocks),
                         "note": "instruction-issue / latency-bound step (fp64 islands, ~300 correctly rounded transcendentals "
                                 "and 4 x (<=20) data-dependent iterations per tile-step): issue-active 48 %, FP64 pipe 24 %, 22.8 of 32 "
                                 "lanes active, DRAM 5-20 % in the ncu capture (profiles/r02_cbm_kernels_ncu_summary.txt); the HBM fraction "
                                 "is reported because it is the official denominator, `issue` is the roof that binds.  traffic > "
                                 "algorithmic bytes: driver-visible diagnostics (336 B/tile at output_level=1), the A->B exchange "
                                 "(132 B/tile) and write-backs of kernel A's ~1 KB/thread of spill slots"},
            "e2e": {"value": e2e, "unit": "tile-timesteps/s", "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                    "steps": Ke, "gpu_launches": e2e_launches, "outputs_finite": e2e_finite, "output_rows": len(rows),
                    "api": "offline driver loop through the C ABI, host buffers: cable_b200_set_met_async (met slice H2D + "
                           "tile expansion) -> cable_b200_step -> cable_b200_post_step -> cable_b200_output_fetch_async "
                           "(grid-cell output rows D2H every step, output%averaging='all')"},
            "e2e_dropin_mirror": mirror,
            "config5_casa_cnp": casa_leg,
            "e2e_dropin_mirror_all": mirror_all,
            "gpu_launches": launches,
            "clocks": clocks,
            "dryleaf_soft_warnings": int(ctr.n_dryleaf_warn),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    h.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nland", type=int, default=NLAND, help="land points of the grid (strong) / per GPU (weak); "
                    "62000 = BASELINE configs[2], 250000 = configs[3]")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--decomp", default="interleaved", choices=["interleaved", "block"],
                    help="strong scaling: deal 64-point chunks round-robin (default) or the reference's contiguous blocks")
    ap.add_argument("--no-flush", action="store_true", help="small shards: do not flush L2 between timed steps")
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-casa", action="store_true", help="skip the config-5 (CASA-CNP) leg")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 on their own (NCCL prints its
    # version banner there at NCCL_DEBUG=VERSION and above) are sent to stderr for the whole run, and print() gets the
    # saved descriptor back
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(saved, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
