"""Turn the raw artefacts of one profiling run (gpurun_out/<tag>_launches.csv = `ncu --metrics gpu__time_duration.sum` launch
list of `bench.py --steps 2 --warmup 1`, gpurun_out/<tag>_render.ncu-rep = `ncu --set full` capture of the render launch,
gpurun_out/<tag>_bench.json) into the tracked summaries under profiles/:   python tools/summarize_profiles.py r1h"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__time_duration.sum',
        'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']


def main(tag):
    out, prof = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    p = os.path.join(out, f"{tag}_launches.csv")
    if os.path.isfile(p):
        rows = list(csv.reader(open(p)))
        hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
        h = rows[hi]
        ki, vi = h.index('Kernel Name'), h.index('Metric Value')
        agg = collections.OrderedDict()
        for r in rows[hi + 1:]:
            if len(r) <= vi:
                continue
            try:
                v = float(r[vi].replace(',', ''))
            except ValueError:
                continue
            a = agg.setdefault(re.sub(r'\(.*', '', r[ki]), [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(v[1] for v in agg.values())
        with open(os.path.join(prof, f"{tag}_launch_list_summary.csv"), "w") as f:
            f.write("kernel,launches,total_ms,share\n")
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{k},{v[0]},{v[1] / 1e6:.3f},{v[1] / tot:.4f}\n")
        print("launch list:", len(agg), "kernels; top:", max(agg.items(), key=lambda kv: kv[1][1])[0])
    p = os.path.join(out, f"{tag}_render.ncu-rep")
    if os.path.isfile(p):
        raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, vals = rows[0], rows[1], rows[2]
        lines = [f"Kernel: shade_tc_kernel, render_core launch of 1184 rays x 128 samples (ncu --set full --clock-control none --import-source on "
                 f"-k regex:shade_tc -s 4 -c 1, tools/ncu_forward.py; report gpurun_out/{tag}_render.ncu-rep)"]
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"{w} | {units[i]} | {vals[i]}")
        open(os.path.join(prof, f"{tag}_shade_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
        i, j = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tr = float(vals[i]) * scale[units[i]] + float(vals[j]) * scale[units[j]]
        tp = os.path.join(prof, "traffic.json")
        d = json.load(open(tp))
        d["source"] = f"profiles/{tag}_shade_ncu_summary.txt (ncu --set full, 1184-ray launch: dram__bytes_read.sum + dram__bytes_write.sum)"
        d["dram_bytes_per_ray"] = tr / 1184
        json.dump(d, open(tp, "w"), indent=1)
        print("ncu summary written; DRAM bytes per ray:", tr / 1184)
    p = os.path.join(out, f"{tag}_bench.json")
    if os.path.isfile(p):
        d = json.loads([l for l in open(p) if l.startswith("{")][-1])   # the JSON line (NCCL / build chatter may precede it)
        json.dump(d, open(os.path.join(prof, f"{tag}_bench.json"), "w"), indent=1)
        print("bench:", d["value"], "rays/s; e2e", d["e2e"]["value"], "; roofline frac", d["roofline"]["frac"])


if __name__ == "__main__":
    main(sys.argv[1])
