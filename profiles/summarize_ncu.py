"""Turn ncu exports brought back in gpurun_out/ into the small summaries committed under profiles/.

  ncu -i gpurun_out/prof_mac.ncu-rep --page raw --csv > /tmp/mac_raw.csv
  python profiles/summarize_ncu.py full /tmp/mac_raw.csv profiles/rNN_k_mac_ncu_full_summary.csv
  python profiles/summarize_ncu.py launches gpurun_out/launches.csv profiles/rNN_launches_summary.csv "<command>"
  python profiles/summarize_ncu.py traffic /tmp/mac_raw.csv profiles/traffic.json <streams> <block> <partitions> [fused=1]
"""
import collections
import csv
import json
import sys

WANT = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'launch__waves_per_multiprocessor', 'sm__inst_executed.sum', 'smsp__inst_executed.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum']
MUL = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}


def full(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write('metric,unit,' + ','.join('launch%d' % i for i in range(len(rows) - 2)) + '\n')
        for k in WANT:
            if k in hdr:
                i = hdr.index(k)
                f.write('%s,%s,%s\n' % (k, units[i], ','.join('"%s"' % r[i] if ',' in r[i] else r[i] for r in rows[2:])))


def traffic(src, dst, streams, block, parts, fused=1):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    ri, wi = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    rd = sum(float(r[ri]) for r in rows[2:]) / (len(rows) - 2) * MUL[units[ri]]
    wr = sum(float(r[wi]) for r in rows[2:]) / (len(rows) - 2) * MUL[units[wi]]
    alg = (streams + 1) * parts * (block + 1) * 8 + (streams * ((block + 1) * 8 + 2 * block * 4) if fused else 0)
    json.dump({"kernel": rows[2][hdr.index('Kernel Name')], "streams": streams, "block": block, "partitions": parts, "dram_bytes_per_launch": rd + wr,
               "fused": bool(fused), "dram_read_bytes": rd, "dram_write_bytes": wr,
               # FDL + shared IR; the fused step also writes the new spectrum and moves the audio block in and out
               "algorithmic_bytes_per_launch": alg, "ratio": (rd + wr) / alg,
               "source": "ncu --set full --clock-control none (%d launches averaged): python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-selfcheck" % (len(rows) - 2)},
              open(dst, 'w'), indent=1)


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(',', '')))
    ours = lambda k: 'at::' not in k and 'cub::' not in k and 'nccl' not in k.lower()      # everything that is not a torch / library kernel
    tot = sum(sum(v) for k, v in agg.items() if ours(k))
    with open(dst, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none : %s\n' % cmd)
        f.write('# per-launch times are cold-cache and serialised: compare SHARES. share = of the time of this library\'s kernels\n')
        f.write('kernel,launches,avg_ns,total_ns,share\n')
        for k, v in agg.items():
            f.write('"%s",%d,%.1f,%.1f,%s\n' % (k[:96], len(v), sum(v) / len(v), sum(v), '%.4f' % (sum(v) / tot) if ours(k) else ''))


if __name__ == '__main__':
    mode = sys.argv[1]
    if mode == 'full':
        full(sys.argv[2], sys.argv[3])
    elif mode == 'traffic':
        traffic(sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7]) if len(sys.argv) > 7 else 1)
    else:
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '')
