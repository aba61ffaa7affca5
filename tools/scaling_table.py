"""`python tools/scaling_table.py profiles/rXX_bench_c4_n1.json ... > profiles/r02_scaling.md`: one table of the per-N bench lines
(value, ms per iteration, efficiency against the first file, e2e and its parts, CUDA-event phases of the last timed iteration)."""
import json
import sys


def load(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


def main():
    lines = [load(p) for p in sys.argv[1:]]
    base = lines[0]
    print("# r02 -- strong scaling of the C4 sweep (n=5000, p=20000, k=100) on 1 / 2 / 4 / 8 B200\n")
    print("Sources: " + ", ".join(f"`{p}`" for p in sys.argv[1:]) + " (`bench.py --gpus N --steps 20 --warmup 3`; N > 1 under torchrun, one rank per GPU,")
    print("NCCL over NVLink; `value` = 20 iterations inside one `bsfg_sweep` call, CUDA events, max over ranks).  Efficiency = value(N) / (N x value(1)).\n")
    print("| GPUs | iter/s | ms / iteration | efficiency | e2e iter/s | e2e efficiency | set_state ms | sweep(1) ms | get_state ms | launches / iteration |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for d in lines:
        n = d["n_gpus"]; e = d["e2e"]; parts = e.get("ms_per_step_parts", {})
        print(f"| {n} | {d['value']:.1f} | {d['ms_per_step']:.3f} | {d['value'] / (n * base['value']):.2f} | {e['value']:.1f} | "
              f"{e['value'] / (n * base['e2e']['value']):.2f} | {parts.get('set_state_h2d', 0):.2f} | {parts.get('sweep', 0):.2f} | "
              f"{parts.get('get_state_d2h', 0):.2f} | {d['gpu_launches'] / d['steps']:.0f} |")
    names = [k for k in base["phases_ms_last_iter"] if k != "update_k"]
    print("\nCUDA-event phases of the last timed iteration (ms); `ideal` = the 1-GPU phase / N.  (In bench lines taken before the library ran")
    print("its phase-timed iteration on one stream, h2, F_a and delta were on the side stream and read ~0.)\n")
    print("| phase | " + " | ".join(f"N={d['n_gpus']}" + (f" (ideal {1})" if False else "") for d in lines) + " | N=8 ideal |")
    print("|---|" + "---|" * (len(lines) + 1))
    for nm in names:
        row = [f"{d['phases_ms_last_iter'][nm]:.3f}" for d in lines]
        print(f"| {nm} | " + " | ".join(row) + f" | {base['phases_ms_last_iter'][nm] / 8:.3f} |")
    print("| **sum** | " + " | ".join(f"{sum(v for k, v in d['phases_ms_last_iter'].items()):.3f}" for d in lines) + f" | {base['ms_per_step'] / 8:.3f} |")


if __name__ == "__main__":
    main()
