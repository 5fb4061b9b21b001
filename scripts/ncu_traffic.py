"""Extract DRAM traffic and executed FP64 flops per launch from an .ncu-rep into profiles/r01_traffic.json"""
import csv, json, subprocess, sys
rep, key, windows, src = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
def val(name, scale_units={"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}):
    v, u = d[name]
    return float(v) * scale_units.get(u, 1)
cyc = float(d["sm__cycles_elapsed.max"][0])
flop = (2 * float(d["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"][0])
        + float(d["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"][0])
        + float(d["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"][0])) * cyc
ms = float(d["gpu__time_duration.sum"][0]) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}[d["gpu__time_duration.sum"][1]]
p = sys.argv[5] if len(sys.argv) > 5 else "profiles/r01_traffic.json"
j = json.load(open(p)) if __import__("os").path.exists(p) else {}
j[key] = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
          "windows_per_launch": windows, "kernel_ms_under_ncu": ms, "fp64_flop_per_launch": flop,
          "fp64_pipe_active_pct": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0]),
          "source": src,
          "fp64_flop_source": "ncu smsp__sass_thread_inst_executed_op_{dfma x2,dadd,dmul}_pred_on.sum.per_cycle_elapsed x sm__cycles_elapsed.max (Newton steps of sqrt and division included)"}
json.dump(j, open(p, "w"), indent=1)
print(key, j[key])
