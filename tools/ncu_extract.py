"""Selected raw metrics of an .ncu-rep as a small CSV (one column per captured launch), and -- with --blocks -- the SASS source page
of its first launch aggregated into runs of equally often executed instructions (code regions).
  python tools/ncu_extract.py REPORT.ncu-rep OUT.csv [--blocks OUT.txt]"""
import collections, csv, math, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__icc_request_hit_rate.pct', 'sm__icc_requests.sum',
        'idc__request_hit_rate.pct', 'idc__requests.sum', 'gcc__cache_requests_type_instruction.sum', 'gcc__cache_requests_type_constant.sum',
        'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed', 'gcc__cache_requests_type_constant.sum.pct_of_peak_sustained_elapsed']
STALLS = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio'
for k in ('wait', 'short_scoreboard', 'long_scoreboard', 'no_instruction', 'math_pipe_throttle', 'branch_resolving', 'not_selected', 'selected', 'mio_throttle',
          'dispatch_stall', 'barrier', 'lg_throttle', 'membar'):
    KEYS.append(STALLS % k)


def page(rep, name, extra=()):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'] + list(extra), capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = page(rep, 'raw')
    hdr, units = rows[0], rows[1]
    H = {h: i for i, h in enumerate(hdr)}
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit'] + ['launch %d: %s' % (k, r[H['Kernel Name']][:90]) for k, r in enumerate(rows[2:])])
        for k in KEYS:
            if k in H:
                w.writerow([k, units[H[k]]] + [r[H[k]] for r in rows[2:]])
    if '--blocks' in sys.argv:
        bout = sys.argv[sys.argv.index('--blocks') + 1]
        src = page(rep, 'source', ['--print-source', 'sass'])
        i0 = [i for i, r in enumerate(src) if r and r[0] == 'Address'][0]
        h2 = src[i0]
        data = []
        for r in src[i0 + 1:]:
            if r and r[0] == 'Kernel Name':
                break
            data.append(r)
        G = {h: i for i, h in enumerate(h2)}
        sc = [h for h in h2 if h.startswith('stall_') and 'Not Issued' not in h]
        ex = [int(r[G['Instructions Executed']]) for r in data]
        sm = [int(r[G['# Samples']]) for r in data]
        th = [float(r[G['Avg. Predicated-On Threads Executed']]) for r in data]
        tot, tex = sum(sm), sum(ex)
        lvl = lambda e: 0 if e == 0 else int(math.log10(e) * 3)
        with open(bout, 'w') as f:
            agg = collections.Counter()
            for r in data:
                for h in sc:
                    agg[h[6:]] += int(r[G[h]])
            f.write('%d static SASS instructions, %d executed at least once (%.1f kB); warp instructions executed %.3e; stall samples %d\n'
                    % (len(data), sum(1 for e in ex if e > 0), 16 * sum(1 for e in ex if e > 0) / 1024.0, tex, tot))
            f.write('stall reasons, %% of samples: %s\n\n' % ', '.join('%s %.1f' % (k, 100.0 * v / tot) for k, v in agg.most_common(9)))
            f.write('%-13s %6s %12s %8s %8s %6s  %s\n' % ('instructions', 'count', 'executions', '% instr', '% samples', 'lanes', 'top stalls (% of all samples)'))
            i = 0
            while i < len(data):
                j = i
                while j + 1 < len(data) and lvl(ex[j + 1]) == lvl(ex[i]):
                    j += 1
                n = j - i + 1
                e, s = sum(ex[i:j + 1]), sum(sm[i:j + 1])
                if n >= 12 and e > 0:
                    a = collections.Counter()
                    for r in data[i:j + 1]:
                        for h in sc:
                            a[h[6:]] += int(r[G[h]])
                    lanes = sum(t * x for t, x in zip(th[i:j + 1], ex[i:j + 1])) / max(e, 1)
                    f.write('%5d - %5d %6d %12.3e %8.2f %8.2f %6.1f  %s\n' % (i, j, n, e / n, 100.0 * e / tex, 100.0 * s / tot, lanes,
                                                                              ', '.join('%s %.1f' % (k, 100.0 * v / tot) for k, v in a.most_common(3))))
                i = j + 1


if __name__ == '__main__':
    main()
