"""Join ncu per-SASS-instruction samples with nvdisasm -g line info (same binary) and aggregate by source line.
usage: python tools/sass_hotspots.py rep.ncu-rep lib.so [topN]"""
import csv, io, re, subprocess, sys, collections, os, tempfile, glob
rep, so = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
# the library holds one cubin per translation unit (ordinary kernels; kernel A's CBL_FASTDIV build): disassemble all
dis = "\n".join(subprocess.run(["nvdisasm", "-g", "-c", c], capture_output=True, text=True).stdout for c in sorted(glob.glob(tmp + "/*.cubin")))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(sass)))
# the page concatenates kernels: "Kernel Name" row, header row, instruction rows ...
starts = [k for k, r in enumerate(allrows) if r and r[0] == "Kernel Name"]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
k0 = starts[which]; k1 = starts[which + 1] if which + 1 < len(starts) else len(allrows)
rows = allrows[k0:k1]
kname = rows[0][1]
hdr = rows[1]; body = rows[2:]
i_s, i_ex, i_ni, i_src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("stall_no_inst"), hdr.index("Source")
# kernel mangled: pick the .text section whose instruction count matches
secs = {}; cur = None; line = None
for l in dis.splitlines():
    if l.startswith(".text."):
        cur = l.strip(); secs[cur] = []; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur and re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): secs[cur].append(line)
match = [k for k, v in secs.items() if len(v) == len(body)]
if not match:
    print("no section with", len(body), "instructions;", {k: len(v) for k, v in secs.items()}); sys.exit(1)
lines = secs[match[0]]
samp = collections.Counter(); ex = collections.Counter(); noinst = collections.Counter(); ninst = collections.Counter()
for r, ln in zip(body, lines):
    s = float(r[i_s] or 0); samp[ln] += s; ex[ln] += float(r[i_ex] or 0); noinst[ln] += float(r[i_ni] or 0); ninst[ln] += 1
tot = sum(samp.values()); totex = sum(ex.values())
print(f"{kname[:60]}: {len(body)} SASS instr, samples {tot:.0f}, warp-instr executed {totex:.3g}")
def srcline(f, n):
    for d in ("cable_b200/csrc", "include"):
        p = os.path.join(d, f)
        if os.path.exists(p):
            L = open(p).read().splitlines(); return L[n - 1].strip()[:90] if n <= len(L) else ""
    return ""
print("--- by source line (samples%, exec%, no_inst share, #sass)")
for ln, s in samp.most_common(top):
    print(f"{s / tot * 100:5.1f}% ex {ex[ln] / totex * 100:5.1f}% ni {noinst[ln] / max(s, 1) * 100:4.0f}% n={ninst[ln]:4d} {ln[0] if ln else '?'}:{ln[1] if ln else 0:4d}  {srcline(*ln) if ln else ''}")
# by 'function' = file + nearest preceding CBL_DEV/CBL_NOINLINE definition
def fn_of(f, n):
    for d in ("cable_b200/csrc", "include"):
        p = os.path.join(d, f)
        if os.path.exists(p):
            L = open(p).read().splitlines()
            for k in range(min(n, len(L)) - 1, -1, -1):
                m = re.match(r'\s*(?:template.*)?(?:CBL_DEV|CBL_NOINLINE|__global__)\s+[\w:<> ]*?(\w+)\(', L[k])
                if m: return m.group(1)
    return f
byfn = collections.Counter(); byfn_ex = collections.Counter(); byfn_n = collections.Counter()
for ln in samp:
    if ln is None: continue
    fn = fn_of(*ln); byfn[fn] += samp[ln]; byfn_ex[fn] += ex[ln]; byfn_n[fn] += ninst[ln]
print("--- by function")
for fn, s in byfn.most_common(30):
    print(f"{s / tot * 100:5.1f}% ex {byfn_ex[fn] / totex * 100:5.1f}% sass={byfn_n[fn]:5d}  {fn}")
