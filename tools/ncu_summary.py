"""Turns `ncu -i X.ncu-rep --page raw --csv` output into the per-launch summary JSON kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv
    python tools/ncu_summary.py gpurun_out/prof_raw.csv profiles/rN_ncu_full_summary.json [--traffic profiles/ncu_traffic.json]

With --traffic the mean DRAM bytes (read + write) per launch of every kernel family is merged into that file
(bench.py copies the dominant kernel's value into roofline.traffic).
"""
import csv
import json
import re
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def family(name):
    if re.search(r"heads_fused_kernel<(\(bool\))?(0|false)\b", name):
        return "heads_fused_kernel<class>"
    if "heads_fused_kernel" in name:
        return "heads_fused_kernel<box>"
    if "heads_ig_kernel" in name:
        return "heads_ig_kernel<predict>" if re.search(r"IgShape<[^>]*(\(bool\)1|true)>", name) else "heads_ig_kernel<tower>"
    if "heads_dwf_kernel" in name:
        return "heads_dwf_kernel<class>" if re.search(r"heads_dwf_kernel<(\(bool\))?(0|false)\b", name) else "heads_dwf_kernel<box>"
    if "heads_dw_kernel" in name:
        return "heads_dw_kernel<predict>" if re.search(r"DwShape<[^>]*(\(bool\)1|true|, 1)>", name) else "heads_dw_kernel<tower>"
    if "heads_wide_kernel" in name:
        return "heads_wide_kernel<64>" if re.search(r"heads_wide_kernel<(\(int\))?64\b", name) else "heads_wide_kernel<128>"
    m = re.search(r"(\w+_kernel)", name)
    return m.group(1) if m else name


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    header, units, data = rows[start], rows[start + 1], rows[start + 2:]
    out, traffic = [], {}
    for r in data:
        if len(r) != len(header):
            continue
        rec = {}
        for h, un, v in zip(header, units, r):
            if h in KEEP:
                rec[h] = ("%s %s" % (v, un)).strip() if un and h not in ("Kernel Name", "Grid Size", "Block Size") else v
        out.append(rec)
        try:
            i_r, i_w = header.index("dram__bytes_read.sum"), header.index("dram__bytes_write.sum")
            b = float(r[i_r].replace(",", "")) * UNIT.get(units[i_r], 1.0) + float(r[i_w].replace(",", "")) * UNIT.get(units[i_w], 1.0)
            traffic.setdefault(family(rec["Kernel Name"]), []).append(b)
        except (ValueError, KeyError):
            pass
    json.dump(out, open(dst, "w"), indent=1)
    print("%d launches -> %s" % (len(out), dst))
    if "--traffic" in sys.argv:
        path = sys.argv[sys.argv.index("--traffic") + 1]
        try:
            cur = json.load(open(path))
        except (OSError, ValueError):
            cur = {}
        for k, v in traffic.items():
            cur[k] = sum(v) / len(v)
        # per kernel FUNCTION (what bench.py's roofline.traffic looks up): mean over its launches
        fn = {}
        for k, v in traffic.items():
            fn.setdefault(k.split("<")[0], []).extend(v)
        for k, v in fn.items():
            cur.setdefault(k, sum(v) / len(v)) if "<" in "".join(x for x in traffic if x.split("<")[0] == k) else None
            cur[k] = sum(v) / len(v)
        if "--step" in sys.argv:   # the capture holds exactly the launches of one udal_run
            cur["step_total"] = sum(sum(v) for v in traffic.values())
        json.dump(cur, open(path, "w"), indent=1)
        print({k: round(sum(v) / len(v) / 1e9, 3) for k, v in traffic.items()}, "GB per launch ->", path)


if __name__ == "__main__":
    main()
