"""Turn one `ncu -i rep --page raw --csv` dump of a kernel into the figures profiles/kernel_metrics.json carries
(bench.py copies them into roofline.traffic / roofline_issue / fp64).  usage: ncu_to_metrics.py raw.csv [sources_per_launch]"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, vals = rows[0], rows[2]
d = dict(zip(hdr, vals))
nsrc = float(sys.argv[2]) if len(sys.argv) > 2 else None


def f(k):
    return float(d[k].replace(",", "")) if k in d and d[k] not in ("", "n/a") else None


def unit(k):
    return rows[1][hdr.index(k)] if k in hdr else ""


def to_bytes(k):
    v, u = f(k), unit(k)
    return None if v is None else v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


dur_us = f("gpu__time_duration.sum") * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(unit("gpu__time_duration.sum"), 1)
cycles = f("sm__cycles_elapsed.max") or f("gpc__cycles_elapsed.max")
rate = lambda op: f("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op) or 0.0  # noqa: E731
flops = (rate("dadd") + rate("dmul") + 2 * rate("dfma")) * cycles if cycles else None
out = {"kernel": d.get("Kernel Name"), "duration_us": dur_us,
       "dram_bytes_per_launch": (to_bytes("dram__bytes_read.sum") or 0) + (to_bytes("dram__bytes_write.sum") or 0),
       "issue_slot_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "fp64_pipe_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
       "lanes_per_instruction": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
       "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
       "warp_instructions": f("smsp__inst_executed.sum"), "registers": f("launch__registers_per_thread"),
       "fp64_flops_per_launch": flops, "l2_hit_pct": f("lts__t_sector_hit_rate.pct")}
if nsrc:
    out["warp_instructions_per_source"] = out["warp_instructions"] / nsrc
print(json.dumps(out, indent=1))
