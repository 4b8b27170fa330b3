"""Print the headline metrics + stall breakdown of an `ncu --page raw --csv` export.  usage: ncu_summary.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum']
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print(d.get('Kernel Name', '?')[:90])
    for k in KEYS:
        if k in d:
            print(f"   {k:78s} {d[k]:>16s} {rows[1][hdr.index(k)]}")
    out = []
    for k, v in d.items():
        if k.startswith('smsp__pcsamp_warps_issue_stalled') and not k.endswith('not_issued'):
            try:
                out.append((float(v), k))
            except ValueError:
                pass
    tot = sum(v for v, _ in out) or 1
    print("   warp-state samples:")
    for v, k in sorted(out, reverse=True)[:9]:
        print(f"      {100 * v / tot:5.1f}%  {k.replace('smsp__pcsamp_warps_issue_stalled_', '')}")

# --traffic <workload:envs> <traffic.json> <source label>: record dram__bytes_read.sum + dram__bytes_write.sum of the step
# kernel's launch in profiles/traffic.json, where bench.py reads `roofline.traffic` from (no literals in bench.py)
if "--traffic" in sys.argv:
    import json
    import os

    i = sys.argv.index("--traffic")
    key, path, label = sys.argv[i + 1], sys.argv[i + 2], sys.argv[i + 3]
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        if "mcb_env_kernel<(int)0>" not in d.get("Kernel Name", "") and "mcb_env_kernel<0>" not in d.get("Kernel Name", ""):
            continue
        rd = float(d["dram__bytes_read.sum"]) * unit[rows[1][hdr.index("dram__bytes_read.sum")]]
        wr = float(d["dram__bytes_write.sum"]) * unit[rows[1][hdr.index("dram__bytes_write.sum")]]
        table = json.load(open(path)) if os.path.exists(path) else {}
        table[key] = {"bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "source": label,
                      "kernel_ms": float(d["gpu__time_duration.sum"])}
        json.dump(table, open(path, "w"), indent=1)
        print("traffic", key, table[key])
        break
