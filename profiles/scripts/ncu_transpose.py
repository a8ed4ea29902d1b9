"""Transpose `ncu -i X.ncu-rep --page raw --csv` into one row per metric (columns = captured launches),
keeping the metric families the round summaries quote.  usage: ncu_transpose.py raw.csv out.csv"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = re.compile(r"^(Kernel Name|Block Size|Grid Size|dram__|gpu__|launch__|lts__t_sector|lts__t_bytes|l1tex__t_sector_hit|l1tex__data_bank|"
                  r"sm__throughput|sm__warps_active|sm__cycles_active|smsp__pcsamp_warps_issue_stalled|smsp__inst_executed\.sum|"
                  r"smsp__cycles_active|sm__inst_executed_pipe_(fp64|lsu|uniform)|l1tex__throughput|lts__throughput)")
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + ["launch_%d" % i for i in range(len(data))])
    for i, h in enumerate(hdr):
        if keep.match(h) and any(r[i] != "" for r in data):
            w.writerow([h, units[i]] + [r[i] for r in data])
