"""profiles/ncu_traffic.json from `ncu --page raw --csv` exports: per kernel the DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum) of ONE launch and the env-steps that launch processed -- bench.py scales it to its own launch size
for `roofline.traffic`.

    python tools/ncu_traffic.py <raw.csv> <kernel key> <env-steps per launch> [<raw.csv> <key> <units> ...]
"""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def dram_bytes(path):
    rows = list(csv.reader(open(path)))
    h, u, v = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = h.index(name)
        tot += float(v[i].replace(",", "")) * UNIT[u[i]]
    t = h.index("gpu__time_duration.sum")
    tens = h.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" in h else None
    return tot, float(v[t].replace(",", "")), u[t], (float(v[tens]) if tens is not None else None)


out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
args = sys.argv[1:]
for i in range(0, len(args), 3):
    path, key, units = args[i], args[i + 1], float(args[i + 2])
    b, dur, du, tens = dram_bytes(path)
    out[key] = {"dram_bytes_per_launch": b, "env_steps_per_launch": units, "dram_bytes_per_env_step": b / units,
                "ncu_duration": f"{dur} {du} (under ncu: cold caches, serialised)", "tensor_pipe_active_pct_of_elapsed": tens,
                "source": os.path.relpath(path, ROOT)}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
