"""One table per `ncu --set full` capture: `python tools/summarize_ncu.py gpurun_out/prof_X.ncu-rep|X_raw.csv profiles/rNN_ncu_X.md "title"`.
Per launch: duration, DRAM bytes read / written, achieved GB/s against the measured HBM copy rate (MEASURED_PEAKS.json),
occupancy, registers, FP64-pipe and issue utilisation, shared-memory bank conflicts, the top stall reasons."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm = 6553.9
    if rep.endswith(".csv"):        # already exported on the GPU box (`ncu -i X.ncu-rep --page raw --csv`)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, default=0.0):
        try:
            return float(r[col[name]].replace(",", ""))
        except Exception:
            return default

    def to_bytes(r, name):
        u = units[col[name]].lower()
        return val(r, name) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    def to_ms(r, name):
        u = units[col[name]].lower()
        return val(r, name) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(u, 1)

    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    lines = [f"# {title or os.path.basename(rep)}", "",
             f"Source: `{os.path.basename(rep)}` (`ncu --set full --clock-control none`, one warm Gibbs iteration of `bench.py --steps 2 --warmup 3`, C4 on one B200).",
             f"GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration; % = of the measured HBM copy rate {hbm:.1f} GB/s (MEASURED_PEAKS.json).",
             "Durations under ncu are cold-cache and serialised; the CUDA-event phase times of `bench.py` are the timing of record.", "",
             "| kernel | grid x block | ms | DRAM read MB | DRAM write MB | GB/s | % HBM | warps active % | regs | FP64 pipe % | issue active % | smem conflicts | top stalls (warps per issue) |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        ms = to_ms(r, "gpu__time_duration.sum")
        rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
        gbs = (rd + wr) / (ms * 1e-3) * 1e-9 if ms > 0 else 0.0
        stalls = sorted(((val(r, h), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall_cols), reverse=True)
        top = ", ".join(f"{n} {v:.2f}" for v, n in stalls if n != "selected")[:80]
        top = ", ".join(f"{n} {v:.2f}" for v, n in [s for s in stalls if s[1] != "selected"][:3])
        lines.append(f"| `{name}` | {r[col['Grid Size']]} x {r[col['Block Size']]} | {ms:.3f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | {100 * gbs / hbm:.1f} | "
                     f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {int(val(r, 'launch__registers_per_thread'))} | "
                     f"{val(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f} | "
                     f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                     f"{int(val(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'))} | {top} |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
