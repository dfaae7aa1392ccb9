"""One `ncu --set full` report of the solve kernel -> the per-QP figures bench.py puts next to the roofline
(profiles/ncu_solve_r02.json). usage: ncu_to_json.py <report.ncu-rep> <problems in the profiled launch> <mean iterations> <out.json>"""
import csv, json, subprocess, sys
rep, P, iters, out = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, row = raw[0], raw[1], raw[2]
a = dict(zip(hdr, row)); u = dict(zip(hdr, units))
def val(k):
    v = float(a[k].replace(",", ""))
    unit = u.get(k, "")
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0}.get(unit, 1.0)
    return v * mult
dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
lts = val("lts__t_bytes.sum") if "lts__t_bytes.sum" in a else 32.0 * val("lts__t_sectors.sum")
inst = val("smsp__inst_executed.sum")
res = {"kernel": a.get("Kernel Name"), "problems": P, "mean_iterations": iters,
       "dram_bytes_per_qp": dram / P, "lts_bytes_per_qp": lts / P,
       "issue_active_pct": float(a["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
       "warp_instructions_per_qp": inst / P, "warp_instructions_per_iteration": inst / P / iters,
       "registers_per_thread": int(float(a["launch__registers_per_thread"])),
       "lts_hit_rate_pct": float(a["lts__t_sector_hit_rate.pct"]),
       "lts_throughput_pct_of_peak": float(a.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", "nan")),
       "l1tex_throughput_pct_of_peak": float(a.get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "nan")),
       "shared_pipe_pct_of_peak": float(a.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "nan")),
       "dram_throughput_pct_of_peak": float(a.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "nan")),
       "warps_active_per_sm": float(a.get("sm__warps_active.avg.per_cycle_active", "nan")),
       "duration_ms_under_ncu": val("gpu__time_duration.sum") / 1e6 if u.get("gpu__time_duration.sum") == "nsecond" else float(a["gpu__time_duration.sum"]),
       "source": f"ncu --set full --clock-control none, one launch of {P} C3 problems, report {rep.split('/')[-1]}"}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
