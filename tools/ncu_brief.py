"""brief per-kernel summary of an .ncu-rep: python tools/ncu_brief.py gpurun_out/prof.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__inst_executed_op_global_st.sum', 'smsp__inst_executed_op_global_ld.sum', 'smsp__inst_executed_op_ldgsts.sum']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=====", d['Kernel Name'][:90])
    for k in keys:
        if k in d and d[k] != '': print(f"   {k:75s} {d[k]}")
    st = []
    for k in hdr:
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio'):
            try: v = float(d[k])
            except ValueError: continue
            if v > 0.2: st.append((v, k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
    print("   stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)))
