"""Summarises an `ncu --csv --log-file` launch list (gpu__time_duration.sum [+ dram bytes]) per kernel:
launch count, total / mean duration, share of the captured GPU time, DRAM bytes per launch.
usage: ncu_launch_summary.py launches.csv out.txt [traffic.json]"""
import csv, io, json, re, sys, collections
src, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(src, errors="replace") if l.startswith('"')]
rows = list(csv.reader(io.StringIO("".join(lines))))
hdr = rows[0]
iname, imetric, ival, iunit, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= ival:
        continue
    k = (r[iid], re.sub(r"\(.*", "", r[iname]))
    v = float(r[ival].replace(",", "") or 0)
    u = r[iunit]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    per.setdefault(k, {})[r[imetric]] = v * scale
agg = collections.OrderedDict()
for (i, name), m in per.items():
    a = agg.setdefault(name, {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0, "max_ms": 0.0})
    a["n"] += 1
    d = m.get("gpu__time_duration.sum", 0.0)
    a["ms"] += d
    a["max_ms"] = max(a["max_ms"], d)
    a["rd"] += m.get("dram__bytes_read.sum", 0.0)
    a["wr"] += m.get("dram__bytes_write.sum", 0.0)
tot = sum(a["ms"] for a in agg.values())
with open(out, "w") as f:
    f.write("# %s: %d launches captured, %.1f ms of GPU time (kernels serialised and cold-cache under ncu:\n"
            "# shares are meaningful, absolute times are not bench numbers)\n" % (src, len(per), tot))
    f.write("%-28s %6s %10s %9s %9s %7s %14s %14s\n" % ("kernel", "n", "total ms", "mean ms", "max ms", "share", "dram rd B/launch", "dram wr B/launch"))
    for name, a in sorted(agg.items(), key=lambda x: -x[1]["ms"]):
        f.write("%-28s %6d %10.2f %9.3f %9.2f %6.1f%% %14.0f %14.0f\n" % (name[:28], a["n"], a["ms"], a["ms"] / a["n"], a["max_ms"],
                                                                      100 * a["ms"] / tot, a["rd"] / a["n"], a["wr"] / a["n"]))
print(open(out).read())
if len(sys.argv) > 3:
    a = agg.get("zb::k_iterate") or agg.get("k_iterate")
    if a:
        json.dump({"k_iterate_bytes_per_launch": (a["rd"] + a["wr"]) / a["n"], "k_iterate_launches_captured": a["n"],
                   "dram_read_bytes_per_launch": a["rd"] / a["n"], "dram_write_bytes_per_launch": a["wr"] / a["n"],
                   "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the k_iterate launches of one bench.py step "
                             "(profiles/r2_launches.txt)"}, open(sys.argv[3], "w"), indent=1)
