"""profiles/ncu_traffic.json from an `ncu --set full` report of one step (read here, no GPU needed):
per kernel the DRAM bytes, duration, L2 hit rate, issue-slot and warp occupancy that bench.py quotes in
roofline.traffic / roofline.ncu.   python tools/ncu_traffic.py <rep.ncu-rep> <cfg_key e.g. cfg1_S2> [out.json]"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0,
        "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def main(path, key, out="profiles/ncu_traffic.json"):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}

    def val(r, name):
        return float(r[col[name]].replace(",", "")) * UNIT.get(units[col[name]], 1.0)

    entry = {}
    for r in body:
        name = r[col["Kernel Name"]].split("(")[0].replace("rpool::", "")
        if name in entry:
            continue                      # first launch of each kernel
        entry[name] = {
            "read": val(r, "dram__bytes_read.sum"), "write": val(r, "dram__bytes_write.sum"),
            "us": val(r, "gpu__time_duration.sum"),
            "l2_hit_pct": float(r[col["lts__t_sector_hit_rate.pct"]]),
            "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
            "warps_active_pct": float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
        }
    try:
        with open(out) as f:
            doc = json.load(f)
    except Exception:  # noqa: BLE001
        doc = {}
    doc["source"] = "ncu --set full --clock-control none, one launch of each kernel (tools/ncu_traffic.py over %s)" % path
    doc["unit"] = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)"
    doc["note"] = ("per-launch values of one ncu --set full capture (cold-ish caches, serialised launches): time and L2 "
                   "hit rate are for share/ratio checks, the bench line's own CUDA-event timings are the throughput numbers")
    doc[key] = entry
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])
