"""Summarise an `ncu --set full` report (read on the CPU box) into a small JSON for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls N] > profiles/rNN_<kernel>.json

Per profiled launch: duration, achieved clocks, tensor / xu / fma / alu pipe utilisation, issue-slot utilisation, DRAM and L2
bytes, registers, shared memory, warp-stall reasons (summed over the source page when --stalls is given, top-N SASS lines).
"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "sm__cycles_elapsed.avg": "sm_cycles",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_active_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed": "xu_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed": "fma_pipe_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed": "alu_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed": "lsu_pipe_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__bytes_read.sum.per_second": "dram_read_rate",
    "dram__bytes_write.sum.per_second": "dram_write_rate",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "launch__registers_per_thread": "registers",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__waves_per_multiprocessor": "waves",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__inst_executed.sum": "warp_instructions",
}


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    start = out.find('"ID"') if name == "raw" else 0
    return list(csv.reader(io.StringIO(out[start:])))


def main():
    rep = sys.argv[1]
    nstall = int(sys.argv[sys.argv.index("--stalls") + 1]) if "--stalls" in sys.argv else 0
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:120]}
        for h, u, v in zip(hdr, units, r):
            if h in WANT:
                try:
                    d[WANT[h]] = {"value": float(v.replace(",", "")), "unit": u}
                except ValueError:
                    d[WANT[h]] = {"value": v, "unit": u}
            elif "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    d.setdefault("stall_cycles_per_issue", {})[h.split("issue_stalled_")[1].split("_per_issue")[0]] = round(float(v), 3)
                except ValueError:
                    pass
        if "stall_cycles_per_issue" in d:
            d["stall_cycles_per_issue"] = dict(sorted(d["stall_cycles_per_issue"].items(), key=lambda kv: -kv[1])[:8])
        launches.append(d)
    out = {"report": rep, "launches": launches}
    if nstall:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        lines = list(csv.reader(io.StringIO(src[src.find('"#"') if '"#"' in src else 0:])))
        if lines:
            h = lines[0]
            try:
                ci = h.index("Source")
                cs = next(i for i, x in enumerate(h) if x.startswith("Warp Stall Sampling (All"))
                top = sorted((l for l in lines[1:] if len(l) > cs and l[cs].replace(",", "").isdigit()), key=lambda l: -int(l[cs].replace(",", "")))[:nstall]
                tot = sum(int(l[cs].replace(",", "")) for l in lines[1:] if len(l) > cs and l[cs].replace(",", "").isdigit())
                out["stall_samples_total"] = tot
                out["top_stall_lines"] = [{"sass": l[ci][:100], "samples": int(l[cs].replace(",", ""))} for l in top]
            except (ValueError, StopIteration):
                out["source_page_columns"] = h
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
