"""Summarises an .ncu-rep (read here on the CPU box with `ncu -i`) into profiles/<name>.md + .json.

usage: python tools/ncu_summary.py gpurun_out/march_c3_r1.ncu-rep profiles/r1_march_c3 [rk4_steps_per_launch]
                                   [--cubin-hash H] [--source-hash H] [--note TEXT]
--cubin-hash: FNV-1a of the NVRTC cubin the capture profiled (tools/render_once.py prints it; linon_march_cubin_hash);
--source-hash: bench.static_source_hash() of the ahead-of-time kernels.  bench.py refuses a summary whose identity is not
the running kernel's.
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


KEYS += [
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_bytes.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.per_cycle_active", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
    "smsp__thread_inst_executed.sum", "smsp__inst_executed_op_branch.sum",
]


def main():
    args = sys.argv[1:]
    extra = {}
    for flag, key in (("--cubin-hash", "cubin_hash"), ("--source-hash", "source_hash"), ("--note", "note")):
        if flag in args:
            i = args.index(flag)
            extra[key] = args[i + 1]
            del args[i:i + 2]
    rep, out = args[0], args[1]
    steps = float(args[2]) if len(args) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    results = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")], **extra}
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
                f.write(f"| {k} | {v:,.6g} {d.get('_units', {}).get(k, '')} |\n" if isinstance(v, (int, float)) else f"| {k} | {v} |\n")
            f.write("\n")
    print(json.dumps([{k: v for k, v in d.items() if not k.startswith('_')} for d in results], indent=1)[:3000])


if __name__ == "__main__":
    main()
