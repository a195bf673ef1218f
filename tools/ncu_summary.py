"""Summarises an .ncu-rep (read here on the CPU box with `ncu -i`) into profiles/<name>.md + .json.

usage: python tools/ncu_summary.py gpurun_out/march_c3_r1.ncu-rep profiles/r1_march_c3 [rk4_steps_per_launch]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    steps = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    results = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[k] = v * UNIT_SCALE.get(units[i], 1.0) if units[i] in UNIT_SCALE else v
                d.setdefault("_units", {})[k] = "bytes" if "byte" in units[i] else ("s" if units[i] in ("ns", "us", "ms", "s") else units[i])
        cyc = d.get("sm__cycles_elapsed.max")
        fl = sum(d.get(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed", 0.0) * w
                 for o, w in (("ffma", 2), ("fmul", 1), ("fadd", 1)))
        if cyc and fl:
            d["fp32_flops_executed"] = fl * cyc            # FADD + FMUL + 2 FFMA, thread level, predicated-on
            d["fp32_tflops"] = fl * cyc / d["gpu__time_duration.sum"] / 1e12
            if steps:
                d["fp32_flops_per_rk4_step"] = fl * cyc / steps
                d["warp_inst_per_rk4_step"] = d["smsp__inst_executed.sum"] * 32 / steps / 32 * 1.0
                d["warp_inst_per_warp_step"] = d["smsp__inst_executed.sum"] / (steps / 32)
        if "dram__bytes_read.sum" in d:
            d["dram_traffic_bytes"] = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
            d["dram_gbs"] = d["dram_traffic_bytes"] / d["gpu__time_duration.sum"] / 1e9
        results.append(d)
    json.dump(results, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu summary of `{rep}`\n\n`ncu --set full --clock-control none --import-source on`; read with `ncu -i ... --page raw --csv`.\n"
                "Times under the profiler are serialised / cold-cache: use shares and per-launch counters, not absolutes.\n\n")
        for d in results:
            f.write(f"## {d['kernel']}\n\n| metric | value |\n|---|---|\n")
            for k, v in d.items():
                if k.startswith("_") or k == "kernel":
                    continue
                f.write(f"| {k} | {v:,.6g} {d.get('_units', {}).get(k, '')} |\n")
            f.write("\n")
    print(json.dumps([{k: v for k, v in d.items() if not k.startswith('_')} for d in results], indent=1)[:3000])


if __name__ == "__main__":
    main()
