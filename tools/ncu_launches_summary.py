"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per kernel family the launch count, total device time, share of the captured window and DRAM bytes per launch.

  python tools/ncu_launches_summary.py gpurun_out/launches.csv [--json profiles/rNN_traffic.json]
"""
import collections
import csv
import io
import json
import sys


def family(name: str) -> str:
    n = name.split("(")[0].replace("void ", "")
    for pre in ("ser::<unnamed>::", "ser::", "at::native::", "at::"):
        n = n.replace(pre, "")
    return n.split("<")[0].strip()


def main():
    path = sys.argv[1]
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    agg = collections.defaultdict(lambda: dict(launches=0, us=0.0, rd=0.0, wr=0.0))
    for d in per.values():
        a = agg[family(d["name"])]
        a["launches"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["us"] for a in agg.values()) or 1.0
    out = {}
    print(f"{len(per)} launches, {tot:.1f} us of device time (cold-cache, serialised: compare shares)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        out[k] = dict(launches=a["launches"], us=round(a["us"], 1), share=round(a["us"] / tot, 4),
                      dram_bytes_per_launch=round((a["rd"] + a["wr"]) / a["launches"], 1),
                      dram_read_bytes=a["rd"], dram_write_bytes=a["wr"])
        print(f"{k:34s} n={a['launches']:4d} {a['us']:9.1f} us share={a['us']/tot*100:5.1f}%  "
              f"DRAM rd={a['rd']/1e6:8.1f} MB wr={a['wr']/1e6:8.1f} MB  per-launch={(a['rd']+a['wr'])/a['launches']/1e6:7.2f} MB")
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(dict(source=path, launches=len(per), total_us=round(tot, 1), kernels=out), f, indent=1)


if __name__ == "__main__":
    main()
