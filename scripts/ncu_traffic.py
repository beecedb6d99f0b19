"""Extract per-launch DRAM traffic of one kernel from an .ncu-rep (read here, no GPU needed) into profiles/ncu_traffic.json,
the file bench.py's roofline.traffic is filled from.
    python scripts/ncu_traffic.py <rep> <kernel substring> <grid> <queries> <hchoice> <tag>"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(rep, kernel, grid, queries, hchoice, tag):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if kernel not in d.get("Kernel Name", ""):
            continue
        def val(k):
            v, u = float(d[k].replace(",", "")), units[hdr.index(k)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        recs.append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    if not recs:
        raise SystemExit("no launch of %s in %s" % (kernel, rep))
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        data = json.load(open(path))
    except Exception:
        data = []
    data = [x for x in data if not (x["kernel"] == kernel and x["grid"] == grid and x["queries"] == queries and x["hchoice"] == hchoice)]
    data.append({"kernel": kernel, "grid": grid, "queries": queries, "hchoice": hchoice, "launches_in_capture": len(recs),
                 "dram_bytes_per_launch": sum(recs) / len(recs), "source": tag})
    json.dump(data, open(path, "w"), indent=1)
    print(data[-1])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6])
