"""Write compact SASS listings of the shipped step kernels into profiles/ (instruction text only, one per line) and a
mnemonic histogram.  usage: python tools/make_sass_listing.py [lib.so]"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "cable_b200", "libcable_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
tag_round = sys.argv[2] if len(sys.argv) > 2 else "r02"
want = {"kernelA_fast_cbm_kernel_1_640_1_1": "4cblf10cbm_kernelILi1ELi640ELi1ELi1ELi0E", "kernelA_small_fast_cbm_kernel_1_128_3_1": "4cblf10cbm_kernelILi1ELi128ELi3ELi1ELi0E",
        "kernelB_cbm_kernel_2_384_2_1": "3cbl10cbm_kernelILi2ELi384ELi2ELi1ELi0E", "casa_kernels": None, "driver_kernels": None}
cur, out = None, collections.defaultdict(list)
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1); cur = None
        for tag, key in want.items():
            if key and key in fn: cur = tag
        if cur is None and any(k in fn for k in ("met_expand", "post_step", "aggregate_kernel", "output_reduce", "grid_reduce")): cur = "driver_kernels"
        if cur is None and "casa" in fn: cur = "casa_kernels"
        if cur: out[cur].append(f"// Function : {fn}")
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", line)
    if m and cur: out[cur].append(f"/*{m.group(1)}*/ {m.group(2).strip()} ;")
for tag, lines in out.items():
    path = os.path.join(root, "profiles", f"{tag_round}_sass_{tag}.txt")
    ops = collections.Counter()
    for l in lines:
        m = re.match(r"/\*[0-9a-f]+\*/ (?:@!?U?P\d+ )?([A-Z0-9_]+)", l)
        if m: ops[m.group(1)] += 1
    with open(path, "w") as fh:
        fh.write(f"// cuobjdump -sass {os.path.basename(so)} (sm_100a), instruction text only; {sum(ops.values())} instructions\n")
        fh.write("// mnemonic histogram: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(24)) + "\n")
        fh.write("\n".join(lines) + "\n")
    print(path, sum(ops.values()), "instructions;", "DFMA", ops["DFMA"], "DMUL", ops["DMUL"], "DADD", ops["DADD"], "FFMA", ops["FFMA"], "MUFU", ops["MUFU"], "BAR", ops["BAR"], "LDG", ops["LDG"], "STG", ops["STG"], "LDL", ops["LDL"], "STL", ops["STL"])
