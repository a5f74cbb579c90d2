"""Summaries of the ncu CSV logs that scripts/gpu_final.sh brings back (profiles/README.md).

    python scripts/summarize_ncu.py launches gpurun_out/final_launches.csv   > profiles/<round>_launch_list_summary.txt
    python scripts/summarize_ncu.py kernels  gpurun_out/final_kernels_ncu.csv > profiles/<round>_tuned_kernels_metrics.txt
"""
import collections
import csv
import re
import sys


def rows(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*$", "", name)
    if name.startswith("at::") or name.startswith("cutlass") or "elementwise" in name or name.startswith("cub::"):
        return "torch: " + re.sub(r"<.*$", "", name)[:60]
    return re.sub(r"<.*$", "", name)


def launches(path):
    per = collections.OrderedDict()
    for r in rows(path):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1.0)
        d = per.setdefault(short(r["Kernel Name"]), [0, 0.0])
        d[0] += 1
        d[1] += ns / 1e6
    tot = sum(v[1] for v in per.values())
    n = sum(v[0] for v in per.values())
    ours = {k: v for k, v in per.items() if not k.startswith("torch:") and not k.startswith("nccl")}
    print(f"# total {tot:.2f} ms over {n} launches")
    print(f"# passion_b200 kernels: {sum(v[0] for v in ours.values())} launches, {sum(v[1] for v in ours.values()):.2f} ms = "
          f"{100 * sum(v[1] for v in ours.values()) / tot:.1f}% of the step's kernel time; the rest is torch glue")
    print("kernel | launches | total ms | share")
    for k, (c, ms) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        if ms / tot < 0.001:
            continue
        print(f"{k} | {c} | {ms:.3f} | {100 * ms / tot:.1f}%")


def kernels(path):
    byid = collections.OrderedDict()
    for r in rows(path):
        d = byid.setdefault(r["ID"], {"name": short(r["Kernel Name"]), "full": r["Kernel Name"], "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
        d[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
    for i, d in byid.items():
        tmpl = re.search(r"<(.*?)>\(", d["full"])
        print(f"== {d['name']}<{tmpl.group(1) if tmpl else ''}>  grid {d['grid']} block {d['block']}")
        for k, v in d.items():
            if isinstance(v, tuple):
                print(f"  {k:70s} {v[0]:>16s} {v[1]}")


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels}[sys.argv[1]](sys.argv[2])
