"""summarise `ncu --set full` reports for profiles/: per kernel the duration, DRAM bytes, L2 hit rate, issue utilisation and the
top stall reasons, plus one machine-readable `traffic <kernel> <bytes per launch>` line per kernel (read by bench.py):
   python tools/ncu_summary.py <out.txt> <report.ncu-rep> [<report2.ncu-rep> ...]"""
import csv
import io
import subprocess
import sys

out_path, reps = sys.argv[1], sys.argv[2:]
lines, traffic = [], {}
for rep in reps:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}

    def val(r, name, scale_to=None):
        if name not in col:
            return float("nan")
        v = float(r[col[name]].replace(",", "")) if r[col[name]] else float("nan")
        u = units[col[name]]
        if scale_to == "bytes":
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        if scale_to == "us":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
        return v

    lines.append("== %s" % rep)
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        rd, wr = val(r, "dram__bytes_read.sum", "bytes"), val(r, "dram__bytes_write.sum", "bytes")
        # machine-readable name: no "void ", no template arguments, no anonymous-namespace prefix; template instances of one
        # kernel (the two phases of solverIterateKernel) add up: bytes per STEP of that kernel
        key = name.replace("void ", "").split("<")[0].split("::")[-1].strip()
        traffic[key] = traffic.get(key, 0.0) + rd + wr
        stalls = sorted(((val(r, n), n.split("issue_stalled_")[1].split("_per")[0]) for n in hdr
                         if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")), reverse=True)[:4]
        lines.append("%-28s %9.1f us  dram rd %8.1f MB wr %8.1f MB  L2 hit %5.1f%%  L1 hit %5.1f%%  sm throughput %5.1f%%  regs %3d  lanes/inst %4.1f  stalls: %s" % (
            name, val(r, "gpu__time_duration.sum", "us"), rd / 1e6, wr / 1e6, val(r, "lts__t_sector_hit_rate.pct"), val(r, "l1tex__t_sector_hit_rate.pct"),
            val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), int(val(r, "launch__registers_per_thread")),
            val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"), ", ".join("%s %.1f" % (n, v) for v, n in stalls)))
with open(out_path, "w") as f:
    f.write("# ncu --set full --clock-control none, one launch per kernel at the settled bench scene (times are cold-cache, serialised)\n")
    f.write("\n".join(lines) + "\n")
    for k, v in traffic.items():
        f.write("traffic %s %.0f\n" % (k, v))
print(open(out_path).read())
