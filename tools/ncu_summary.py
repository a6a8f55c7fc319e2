"""Summarise an .ncu-rep: python tools/ncu_summary.py rep [--source N]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
for r in data:
    print("kernel:", r[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f"  {w:72s} {r[i]:>18s} {units[i]}")
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try: v = float(r[i])
            except: continue
            if v > 0.05: print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:8.3f}")
if len(sys.argv) > 2 and sys.argv[2] == '--source':
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    print(src[:200] if len(src) < 300 else "")
    rows = list(csv.reader(io.StringIO(src)))
    # find header row
    for k, r in enumerate(rows):
        if 'Source' in r and any('Sampl' in c for c in r):
            h = r; body = rows[k + 1:]; break
    else:
        print("no source table"); sys.exit()
    isrc = h.index('Source'); isamp = [i for i, c in enumerate(h) if c.startswith('# Samples') or c == 'Warp Stall Sampling (All Samples)'][0]
    iexec = [i for i, c in enumerate(h) if c == 'Instructions Executed']
    tot = 0; items = []
    for r in body:
        try: s = float(r[isamp])
        except: continue
        tot += s; items.append((s, r[0] if r[0] else '', r[isrc][:110], r[iexec[0]] if iexec else ''))
    items.sort(reverse=True)
    print("total samples", tot)
    for s, ln, txt, ex in items[:n]: print(f"{s / tot * 100:6.2f}%  L{ln:>5s} ex={ex:>10s} {txt}")
