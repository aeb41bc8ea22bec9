"""profiles/rNN_k_match_traffic.json from an `ncu --set full` report of the match kernel (dev tool).
usage: python tools/ncu_traffic_json.py gpurun_out/prof_r02.ncu-rep 268435456 "<command the capture was taken from>" > profiles/r02_k_match_traffic.json"""
import csv, io, json, subprocess, sys
rep, n_in, what = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
hdr, units, val = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name)
    v, u = float(val[i].replace(",", "")), units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return v * scale
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
inst = get("smsp__inst_executed.sum")
print(json.dumps({
    "source": what,
    "kernel": val[hdr.index("Kernel Name")].split("(")[0],
    "launch_input_bytes": n_in,
    "dram__bytes_read.sum": rd, "dram__bytes_write.sum": wr,
    "dram_bytes_per_launch": rd + wr, "dram_bytes_per_input_byte": (rd + wr) / n_in,
    "gpu__time_duration_ms": get("gpu__time_duration.sum"),
    "smsp__inst_executed.sum": inst, "warp_inst_per_input_byte": inst / n_in,
    "avg_threads_per_inst": get("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "smsp__issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "sm__throughput_pct": get("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    "shared_wavefronts": get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "lts__t_sector_hit_rate_pct": get("lts__t_sector_hit_rate.pct"),
    "registers_per_thread": get("launch__registers_per_thread"),
    "dynamic_smem_bytes": get("launch__shared_mem_per_block_dynamic"),
}, indent=2))
