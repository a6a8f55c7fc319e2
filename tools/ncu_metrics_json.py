"""ncu report -> profiles/<tag>_ncu_metrics.json (what bench.py attaches to its roofline object).
usage: python tools/ncu_metrics_json.py rep.ncu-rep profiles/r01_ncu_metrics.json"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
def col(r, name):
    try: return float(r[hdr.index(name)].replace(",", ""))
    except Exception: return None
U = hdr.index
def flop_counts(r):
    cyc = col(r, "smsp__cycles_elapsed.avg")
    n = {}
    for op in ("dfma", "dadd", "dmul", "ffma", "fadd", "fmul"):
        rate = col(r, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed")
        n[op] = (rate or 0.0) * (cyc or 0.0)
    return {"thread_inst": {k: round(v) for k, v in n.items()},
            "flops_fp64": 2 * n["dfma"] + n["dadd"] + n["dmul"], "flops_fp32": 2 * n["ffma"] + n["fadd"] + n["fmul"],
            "sm_cycles_elapsed": cyc}
kernels = []
for r in data:
    unit = {n: rows[1][U(n)] for n in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    tscale = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}
    kernels.append({
        "kernel": r[U("Kernel Name")][:48],
        "duration_ms": col(r, "gpu__time_duration.sum") * tscale.get(unit["gpu__time_duration.sum"], 1e-3) * 1e3,
        "dram_bytes": col(r, "dram__bytes_read.sum") * scale[unit["dram__bytes_read.sum"]] + col(r, "dram__bytes_write.sum") * scale[unit["dram__bytes_write.sum"]],
        "issue_active_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "pipe_fp64_pct": col(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "pipe_fma_fp32_pct": col(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "pipe_alu_pct": col(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "pipe_xu_pct": col(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "dram_pct_of_peak": col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "warp_inst_executed": col(r, "smsp__inst_executed.sum"),
        "active_threads_per_warp_inst": col(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "icache_hit_pct": col(r, "sm__icc_request_hit_rate.pct"),
        # executed floating-point operations, counted by the hardware: predicated-on thread instructions by opcode
        # (rate summed over the SM sub-partitions x elapsed cycles); FMA = 2 flops
        **flop_counts(r),
        "registers_per_thread": col(r, "launch__registers_per_thread"),
        "block_size": col(r, "launch__block_size"), "grid_size": col(r, "launch__grid_size"),
    })
TILES = 310000
for k in kernels:
    k["flops_fp64_per_tile_step"] = k["flops_fp64"] / TILES
    k["flops_fp32_per_tile_step"] = k["flops_fp32"] / TILES
json.dump({"source": rep.split("/")[-1], "command": "ncu --set full --clock-control none --import-source on -k regex:cbm_kernel -s 96 -c 12 python tools/quick_perf.py 62000 12 (tools/gpu_profile.sh): the twelve launches of one resident step = four chunk chains (three rounds of 148 x 640 tiles, then the 25 840-tile remainder as 128-thread blocks), each chain: kernel A CBL_FASTDIV build, ordinary kernel A over the blocks it handed back (none here), kernel B.  Under ncu the launches are serialised; in the timed step the chains overlap on four streams",
           "tiles": TILES,
           "flops_per_tile_step": {"fp64": sum(k["flops_fp64_per_tile_step"] for k in kernels), "fp32": sum(k["flops_fp32_per_tile_step"] for k in kernels),
                                   "how": "executed, counted by ncu (smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul,ffma,fadd,fmul}_pred_on; FMA = 2), "
                                          "all twelve launches of the step; includes the fp64 evaluation of the correctly rounded fp32 intrinsics and the IEEE divide / square-root sequences"},
           "kernels": kernels}, open(out, "w"), indent=1)
print(open(out).read()[:1500])
