"""Summarises an `ncu --set full` capture of the step kernels for profiles/ (run here, no GPU needed):

    python tools/ncu_extract.py gpurun_out/r02_v9_step.ncu-rep profiles/r02_v9_kernels_raw_summary.txt "header text" [--json]

Picks the LAST captured launch of rk45_init_kernel / rk45_attempt_kernel / head_kernel, prints the metrics the DESIGN /
VERDICT discussion uses side by side and, with --json, rewrites profiles/ncu_attempt_kernel.json (the numbers bench.py
reports but cannot measure itself: DRAM traffic per launch and executed FP64 flops per env-step)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = ("rk45_init_kernel", "rk45_attempt_kernel", "head_kernel")
METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "smsp__cycles_elapsed.avg",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
]
UNIT_SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def main():
    rep, out_txt, header = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    last = {}
    for r in data:
        for k in KERNELS:
            if k in r[ix["Kernel Name"]]:
                last[k] = r
    missing = [k for k in KERNELS if k not in last]
    if missing:
        raise SystemExit("kernels not in the capture: %s" % missing)
    n_envs = 65536

    def val(k, m):
        i = ix.get(m)
        if i is None or last[k][i] in ("", "n/a"):
            return None
        return float(last[k][i].replace(",", ""))

    lines = [header, "%-78s %-16s %s" % ("Kernel Name", "", " | ".join(last[k][ix["Kernel Name"]][:44] for k in KERNELS))]
    for m in METRICS:
        if m not in ix:
            continue
        vs = [val(k, m) for k in KERNELS]
        lines.append("%-78s %-16s %s" % (m, units[ix[m]], " | ".join("n/a" if v is None else ("%.6f" % v if v != int(v) else "%d" % v) for v in vs)))
    flops, dram, us = {}, {}, {}
    for k in KERNELS:
        cyc = val(k, "smsp__cycles_elapsed.avg")
        ops = [val(k, "smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % o) for o in ("dfma", "dmul", "dadd")]
        if cyc is not None and None not in ops:
            flops[k] = (2 * ops[0] + ops[1] + ops[2]) * cyc / n_envs
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += val(k, m) * UNIT_SCALE.get(units[ix[m]], 1.0)
        dram[k] = b
        us[k] = val(k, "gpu__time_duration.sum") * UNIT_SCALE.get(units[ix["gpu__time_duration.sum"]], 1.0)
    lines.append("executed FP64 flops per env-step (2 DFMA + DMUL + DADD thread instructions / %d envs): %s -> total %.0f" % (
        n_envs, {k: round(v) for k, v in flops.items()}, sum(flops.values())))
    lines.append("DRAM bytes per launch: %s" % {k: round(v) for k, v in dram.items()})
    open(out_txt, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if "--json" in sys.argv:
        js = {"source": "%s (ncu --set full --clock-control none, bench.py --steps 5 --warmup 3 --no-e2e --no-extra, the last "
                        "captured launch of each kernel of a step, 65536 envs); written by tools/ncu_extract.py" % os.path.relpath(out_txt, ROOT),
              "attempt_kernel_dram_bytes_per_launch": dram["rk45_attempt_kernel"], "dram_bytes_per_launch": dram,
              "executed_fp64_flop_per_env_step": sum(flops.values()), "executed_fp64_flop_per_env_step_by_kernel": flops,
              "kernel_us_under_ncu": us,
              "note": "traffic = dram__bytes_read.sum + dram__bytes_write.sum of one rk45_attempt_kernel launch; executed flops = "
                      "(2*DFMA + DMUL + DADD thread instructions, smsp__sass_thread_inst_executed_op_*_pred_on) of the three "
                      "kernels / 65536 envs; bench.py reads this file, nothing is typed in"}
        json.dump(js, open(os.path.join(ROOT, "profiles", "ncu_attempt_kernel.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
