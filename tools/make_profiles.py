#!/usr/bin/env python
"""Turns gpurun_out/{<tag>_launches.csv, <tag>_conv_umma.ncu-rep} into the tracked summaries under profiles/.
usage: tools/make_profiles.py [tag] [command line that was profiled]"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles"); os.makedirs(P, exist_ok=True)
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
cmdline = sys.argv[2] if len(sys.argv) > 2 else "python bench.py --steps 2 --warmup 3 --batch 32768 --no-cpu-baseline --no-e2e"

# 1. launch list: per-kernel totals and shares
rows = list(csv.DictReader(l for l in open(os.path.join(G, tag + "_launches.csv")) if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
    k = r["Kernel Name"].split("(")[0]
    a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]]); a[0] += 1; a[1] += float(r["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, tag + "_launches_summary.md"), "w") as f:
    f.write("# ncu launch list (%s): `ncu --metrics gpu__time_duration.sum --clock-control none` over\n" % tag)
    f.write("`%s` (cold-cache,\nserialised: compare SHARES with bench.py's live per-kernel times, not absolutes)\n\n" % cmdline)
    f.write("| kernel | launches | grid | block | total us | share |\n|---|---|---|---|---|---|\n")
    for k, v in agg.items():
        f.write("| `%s` | %d | %s | %s | %.1f | %.3f |\n" % (k, v[0], v[2], v[3], v[1], v[1] / tot))
    conv = sum(v[1] for k, v in agg.items() if "conv_umma_kernel" in k and "<1, 3" not in k)
    f.write("\nshare of the eight 64->64 convolution launches (the `roofline` kernel family): %.3f\n" % (conv / tot))
import shutil
shutil.copy(os.path.join(G, tag + "_launches.csv"), os.path.join(P, tag + "_launches.csv"))

# 2. full capture: selected metrics per profiled launch
out = subprocess.run(["ncu", "-i", os.path.join(G, tag + "_conv_umma.ncu-rep"), "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rr = list(csv.reader(out.splitlines())); hdr, units, data = rr[0], rr[1], rr[2:]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_uniform.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum"]
ki = hdr.index("Kernel Name")
with open(os.path.join(P, tag + "_conv_umma_full.md"), "w") as f:
    f.write("# ncu --set full (%s), conv_umma_kernel launches of one forward pass (32768 candidates, C=26)\n\n" % tag)
    f.write("`ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s 10 -c 9`; per launch.\n")
    f.write("`traffic` for bench.py = dram__bytes_read.sum + dram__bytes_write.sum.\n\n")
    for r in data:
        f.write("## `%s`\n\n| metric | value | unit |\n|---|---|---|\n" % r[ki].split("(")[0])
        for w in want:
            if w in hdr:
                i = hdr.index(w); f.write("| %s | %s | %s |\n" % (w, r[i], units[i]))
        f.write("\n")
# 3. DRAM traffic of the conv family (bench.py roofline.traffic)
ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
def tobytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
fam = [r for r in data if "conv_umma_kernel" in r[ki] and "<1, 3" not in r[ki]]
tr = sum(tobytes(r[ir], units[ir]) + tobytes(r[iw], units[iw]) for r in fam)
json.dump({"kernel_family": "conv_umma_kernel (eight 64->64 convolutions)", "launches": len(fam), "candidates_per_launch": 32768,
           "dram_bytes_total": tr, "dram_bytes_per_launch": tr / max(len(fam), 1), "dram_bytes_per_candidate": tr / 32768,
           "source": "profiles/%s_conv_umma_full.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % tag},
          open(os.path.join(P, tag + "_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, tag + "_launches_summary.md")).read())
