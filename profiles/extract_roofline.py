"""Summarise `ncu --set full` reports (gpurun_out/<tag>_<kernel>.ncu-rep) into profiles/<tag>_ncu_summary.json and
write profiles/roofline_capture.json (DRAM traffic of the dominant kernel, read by bench.py).
    python profiles/extract_roofline.py <tag> [<commit>]"""
import csv
import glob
import io
import json
import subprocess
import sys
from os.path import basename, dirname, join, realpath

ROOT = dirname(dirname(realpath(__file__)))
KEYS = {
    "gpu__time_duration.sum": "ms",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed": "l1_writeback_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed": "fma_pipe_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed": "fp64_pipe_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_inst_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "tensor_inst_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "memory_l1_wavefronts_shared": "smem_wavefronts",
    "memory_l1_wavefronts_shared_ideal": "smem_wavefronts_ideal",
}
UNIT_SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def summarise(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {"kernel": vals[hdr.index("Kernel Name")].split("(")[0]}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            out[KEYS[h]] = x * UNIT_SCALE.get(u, 1.0)
    for h, u, v in zip(hdr, units, vals):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            name = h.split("issue_stalled_")[1].split("_per_issue")[0]
            try:
                x = float(v)
            except ValueError:
                continue
            if x >= 0.3:
                out.setdefault("stalls_per_issue", {})[name] = round(x, 2)
    return out


if __name__ == "__main__":
    tag = sys.argv[1]
    commit = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(
        ["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    summary = {"tag": tag, "commit": commit, "kernels": {}}
    for path in sorted(glob.glob(join(ROOT, "gpurun_out", f"{tag}_*.ncu-rep"))):
        name = basename(path)[len(tag) + 1:-len(".ncu-rep")]
        summary["kernels"][name] = summarise(path)
    with open(join(ROOT, "profiles", f"{tag}_ncu_summary.json"), "w") as fh:
        json.dump(summary, fh, indent=1)
    f = summary["kernels"].get("filter")
    if f:
        cap = {"file": f"profiles/{tag}_ncu_summary.json", "commit": commit, "kernel": f["kernel"],
               "launch": "filter launch of the profiled pass", "structures": int(f.get("grid", 0)) // 2,
               "ms": f.get("ms"), "dram_bytes_per_launch": f.get("dram_read_bytes", 0) + f.get("dram_write_bytes", 0)}
        with open(join(ROOT, "profiles", "roofline_capture.json"), "w") as fh:
            json.dump(cap, fh, indent=1)
    print(json.dumps(summary, indent=1))
