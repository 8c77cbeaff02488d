#!/usr/bin/env python
"""Summarise bench json files and an optional ncu report of a gpurun_out/<tag> directory."""
import csv, glob, json, subprocess, sys, os
d = sys.argv[1]
for f in ([] if d.endswith('.ncu-rep') else sorted(glob.glob(os.path.join(d, 'bench*.json')))):
    try:
        j = json.load(open(f))
    except Exception as e:
        print(os.path.basename(f), 'ERR', e); continue
    r = j['roofline']
    print(os.path.basename(f), 'rays/s %.3g' % j['value'], 'ms %.2f' % j['ms_per_step'], 'e2e %.3g' % j['e2e']['value'],
          {k: round(v, 2) for k, v in r['stage_ms_per_step'].items()}, 'frac %.2f' % r['frac'],
          {k: int(v) for k, v in j['config']['per_step_counts'].items()}, 'err %.2e' % j['max_abs_err_vs_oracle_512rays'])
reps = [d] if d.endswith('.ncu-rep') else [os.path.join(d, 'prof.ncu-rep')]
for rep in reps:
  if os.path.exists(rep):
      out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
      rows = list(csv.reader(out.splitlines()))
      hdr = rows[0]
      want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
              'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
              'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
              'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
              'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
              'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
              'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
              'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
              'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
              'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
              'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
              'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
              'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_selected_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
            'l1tex__data_bank_conflicts_pipe_lsu.sum', 'sm__inst_executed_pipe_tensor.sum',
            'smsp__inst_executed_pipe_xu.sum', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size',
              'smsp__cycles_active.avg']
      for vals in rows[2:]:
          print('----')
          for w in want:
              if w in hdr:
                  print('  ', w, '=', vals[hdr.index(w)])
