#!/usr/bin/env python3
"""profiles/<round>_ncu_full.csv -> profiles/ncu_traffic.json (DRAM bytes per launch of each profiled kernel; bench.py reads it
for roofline.traffic).     python profiles/make_traffic.py profiles/r01_v5_ncu_full.csv"""
import csv, json, os, sys

src = sys.argv[1]
rows = list(csv.reader(open(src)))
h = rows[0]
col = {n: i for i, n in enumerate(h)}
names = {"conv1_fused_tc_kernel": "conv1_fused_tc", "conv_mid_tc_kernel": "conv_mid_tc"}
out = {}
for key, short in names.items():
    sel = [r for r in rows[2:] if key in r[0]]
    if not sel:
        continue
    # full sub-batches only: a chunk's last, partial sub-batch and (for conv_mid) the conv3 launches read far less
    most = max(float(r[col["dram__bytes_read.sum"]]) for r in sel)
    sel = [r for r in sel if float(r[col["dram__bytes_read.sum"]]) >= 0.9 * most]
    f = lambda c: sum(float(r[col[c]]) for r in sel) / len(sel)
    out[short] = {
        "dram_bytes_per_launch": (f("dram__bytes_read.sum") + f("dram__bytes_write.sum")) * 1e6,
        "dram_read_bytes_per_launch": f("dram__bytes_read.sum") * 1e6,
        "dram_write_bytes_per_launch": f("dram__bytes_write.sum") * 1e6,
        "launches_profiled": len(sel),
        "avg_duration_us_under_ncu": f("gpu__time_duration.sum"),
        "tensor_pipe_active_pct": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "dram_throughput_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "source": f"{src} (ncu --set full --clock-control none, full sub-batch launches only; ncu flushes the L2 between kernels)",
    }
# conv12_frames_kernel: launches of any frame count, so the figures are kept PER FRAME (argv[2] = the ncu capture of
# tools/ncu_case.py, argv[3] = frames per captured launch); bench.py multiplies by the frames its launches cover
if len(sys.argv) > 3:
    rows2 = list(csv.reader(open(sys.argv[2])))
    col2 = {n: i for i, n in enumerate(rows2[0])}
    frames = int(sys.argv[3])
    sel = [r for r in rows2[2:] if "conv12_frames_kernel<48, 1>" in r[col2["Kernel Name"]]]
    if sel:
        g = lambda c: sum(float(r[col2[c]]) for r in sel) / len(sel)
        out["conv12_frames"] = {
            "dram_bytes_per_frame": (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) * 1e6 / frames,
            "dram_read_bytes_per_frame": g("dram__bytes_read.sum") * 1e6 / frames,
            "dram_write_bytes_per_frame": g("dram__bytes_write.sum") * 1e6 / frames,
            "frames_per_profiled_launch": frames, "launches_profiled": len(sel),
            "avg_duration_us_under_ncu": g("gpu__time_duration.sum"),
            "tensor_pipe_active_pct": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "dram_throughput_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "l2_hit_rate_pct": g("lts__t_sector_hit_rate.pct"),
            "sm_clock_ghz_during_capture": g("sm__cycles_elapsed.avg.per_second"),
            "source": f"{sys.argv[2]} (ncu --set full --clock-control none, 720p, two frames per CTA; ncu flushes the L2 between passes)",
        }
# bench.py's kernel table names conv2/conv3 separately; both are conv_mid_tc_kernel
if "conv_mid_tc" in out:
    out["conv2_tc"] = out["conv_mid_tc"]
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
