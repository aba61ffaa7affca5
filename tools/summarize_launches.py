"""`python tools/summarize_launches.py gpurun_out/TAG_launches.csv profiles/TAG_launch_summary.md "title"`: the launch list of
ONE warm Gibbs iteration (tools/profile_iteration.sh: `ncu --profile-from-start off --metrics gpu__time_duration.sum`, the
iteration bracketed by cudaProfilerStart/Stop inside the library) as a table of kernels with their share of the iteration."""
import collections
import csv
import sys


def main():
    src, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > 5]
    ki, vi, gi, bi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size"), hdr.index("Metric Unit")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
    items = [(r[ki].split("(")[0].replace("void ", "").replace("bsfg::", ""), r[gi], r[bi], float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-3)) for r in data]
    total = sum(x[3] for x in items)
    agg = collections.OrderedDict()
    for name, grid, block, us in items:
        a = agg.setdefault((name, grid, block), [0, 0.0])
        a[0] += 1; a[1] += us
    lines = [f"# {title or src}", "",
             f"Source: `{src.split('/')[-1]}` -- `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` over",
             "`python bench.py --steps 2 --warmup 3` with `BSFG_PROFILE_ITER=4` (the library brackets its 5th Gibbs iteration with",
             "cudaProfilerStart/Stop; `tools/profile_iteration.sh`).  Per-launch times under ncu are serialised and cold-cache: compare the",
             "SHARES with the CUDA-event phase times of `bench.py`, not the absolutes.", "",
             f"{len(items)} launches, {total / 1e3:.3f} ms in total.", "",
             "| kernel | grid | block | launches | total us | share |", "|---|---|---|---|---|---|"]
    for (name, grid, block), (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{name}` | {grid} | {block} | {cnt} | {us:.1f} | {100 * us / total:.1f}% |")
    by = collections.Counter()
    for (name, _, _), (cnt, us) in agg.items():
        by[name] += us
    lines += ["", "By kernel:", "", "| kernel | total us | share |", "|---|---|---|"]
    for name, us in by.most_common():
        lines.append(f"| `{name}` | {us:.1f} | {100 * us / total:.1f}% |")
    lines += ["", "In launch order:", "", "| # | kernel | grid | us |", "|---|---|---|---|"]
    for i, (name, grid, block, us) in enumerate(items):
        lines.append(f"| {i} | `{name}` | {grid} | {us:.1f} |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
