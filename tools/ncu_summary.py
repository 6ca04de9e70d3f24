#!/usr/bin/env python3
"""Turn an .ncu-rep (one kernel, `--set full`) into the short summary kept under profiles/.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_ldpc [frames [links_per_frame mean_iters]]  -> .md and .json
With frames, links per frame and the mean iteration count of the profiled launch it also derives what bench.py
needs for its on-chip roofline: ALU-pipe warp instructions and shared-memory wavefronts per edge update."""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "launch__occupancy_limit_registers": "occupancy_limit_registers(blocks)",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_shared_mem(blocks)",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed": "shared_pipe_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__cycles_active.sum": "sm_cycles_active_sum",
    "sm__cycles_elapsed.avg.per_second": "sm_clock_hz",
}
UNIT = {"Ghz": 1e9, "Mhz": 1e6, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0,
        "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    res = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
    stalls = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            res[KEYS[h]] = x * UNIT[u] if u in UNIT else x
            if u in UNIT:
                res[KEYS[h] + "_unit"] = "s" if u in ("ns", "us", "ms", "s", "second") else "B"
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v)
    res["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    if "dram_read" in res and "dram_write" in res:
        res["dram_traffic_bytes"] = res["dram_read"] + res["dram_write"]
    if len(sys.argv) > 3:
        res["frames_in_launch"] = int(sys.argv[3])
        res["dram_traffic_bytes_per_frame"] = res.get("dram_traffic_bytes", 0.0) / int(sys.argv[3])
    if len(sys.argv) > 5:
        links, iters = int(sys.argv[4]), float(sys.argv[5])
        edge_updates = res["frames_in_launch"] * links * iters
        res["links_per_frame"], res["mean_iters"], res["edge_updates_in_launch"] = links, iters, edge_updates
        # pipe_alu: per cent of the ALU pipe's peak (2 warp instructions per clock per SM, profiles/r02_onchip_peaks.json)
        alu_instr = res["alu_pipe_pct"] / 100.0 * 2.0 * res["sm_cycles_active_sum"]
        res["alu_warp_instr_in_launch"] = alu_instr
        res["alu_warp_instr_per_edge_update"] = alu_instr / edge_updates
        res["warp_instr_per_edge_update"] = res["warp_instructions"] / edge_updates
        res["smem_wavefronts_per_edge_update"] = res["shared_wavefronts"] / edge_updates
    for k in ("sm_clock_hz",):
        if k in res and res.get(k + "_unit"):
            res.pop(k + "_unit")
    json.dump(res, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write("# ncu --set full summary: %s\n\nsource: `%s` (one launch, `--clock-control none`)\n\n| metric | value |\n|---|---|\n" % (res["kernel"][:80], rep))
        for k, v in res.items():
            if k in ("kernel", "stall_warps_per_issue") or k.endswith("_unit"):
                continue
            f.write("| %s | %s |\n" % (k, ("%.4g" % v) if isinstance(v, float) else v))
        f.write("\nstall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % kv for kv in res["stall_warps_per_issue"].items()) + "\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
