#!/usr/bin/env python
"""Turns gpurun_out/*.ncu-rep / launches.csv into the small text summaries kept under profiles/.

    python tools/summarize_profile.py launches gpurun_out/launches.csv profiles/r01_launches.md
    python tools/summarize_profile.py kernel   gpurun_out/attn_prof.ncu-rep profiles/r01_attention_ncu.md
    python tools/summarize_profile.py kernel   gpurun_out/attn_prof.raw.csv profiles/r01_attention_ncu.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        name = d["Kernel Name"].split("(")[0][:90]
        v = float(d["Metric Value"].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0, d["Grid Size"], d["Block Size"]])
        a[0] += 1
        a[1] += v
    # ncu prints the bare function name for some builds (no namespace): match our kernels by name
    mine = ("vltr", "attn::", "paged_attention", "merge_partials", "reshape_and_cache", "rms_",
            "rope_", "act_", "rank_", "gemm_tn", "layernorm_kernel", "embed_kernel", "score_kernel",
            "sample_kernel", "block_copy", "prefill_attention", "tp_fused", "fp8_kernel",
            "add_pos_kernel", "attention_kernel")
    ours = {k: v for k, v in agg.items() if any(m in k for m in mine)}
    tot = sum(v[1] for v in ours.values())
    with open(dst, "w") as f:
        f.write(f"# launch list ({src}): `ncu --metrics gpu__time_duration.sum --clock-control none`\n\n")
        f.write("Per-launch times are cold-cache and serialised; compare SHARES.  Library kernels\n"
                "(torch fills/copies of the benchmark set-up) are listed last and excluded from the shares.\n\n")
        f.write("| kernel | launches | total us | share of our kernels | avg us | grid | block |\n|---|---|---|---|---|---|---|\n")
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]/1e3:.1f} | {v[1]/tot*100:.1f}% | {v[1]/v[0]/1e3:.1f} | {v[2]} | {v[3]} |\n")
        f.write("\nother (set-up, not ours):\n\n")
        for k, v in agg.items():
            if k not in ours:
                f.write(f"- `{k}` x{v[0]}: {v[1]/1e3:.1f} us\n")
    print("wrote", dst)


def kernel(src, dst):
    if src.endswith(".csv"):      # already exported on the GPU box (tools/gpu_session.sh export_rep)
        out = open(src).read()
    else:
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"## `{d['Kernel Name'][:120]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            f.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
