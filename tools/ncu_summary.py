"""Summarise ncu outputs into profiles/: launch list shares + key metrics of the full capture."""
import collections
import csv
import subprocess
import sys


def launch_shares(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(',', ''))
        v = {'ns': v / 1e6, 'us': v / 1e3, 'ms': v, 's': v * 1e3}.get(r[mu], v)
        name = r[kn].split('(')[0].replace('ssdb::<unnamed>::', '')[:70]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out = ['# launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)',
           'total %.3f ms over %d launches' % (tot, sum(v[0] for v in agg.values()))]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        out.append('%-72s n=%4d %9.3f ms %5.1f%%' % (k, v[0], v[1], 100 * v[1] / tot))
    return '\n'.join(out)


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size']


def full_metrics(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = ['# ncu --set full --clock-control none (%s)' % rep]
    extra = ['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg.per_second', 'lts__t_bytes.sum',
             'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum']
    want = WANT + [w for w in extra if w not in WANT] + sorted(h for h in hdr if 'utchmma' in h and 'pct_of_peak_sustained_elapsed' in h and h not in WANT)
    labels = LABELS or []
    for k, r in enumerate(data):
        out.append('--- ' + r[idx['Kernel Name']].split('(')[0] + ('   [%s]' % labels[k] if k < len(labels) else ''))
        for w in want:
            if w in idx and r[idx[w]] not in ('', 'n/a'):
                out.append('  %-92s %16s %s' % (w, r[idx[w]], units[idx[w]]))
    return '\n'.join(out)


LABELS = None

if __name__ == '__main__':
    kind, src, dst = sys.argv[1:4]
    if len(sys.argv) > 4:
        LABELS = sys.argv[4].split(',')          # one label per captured launch, in launch order
    text = launch_shares(src) if kind == 'launches' else full_metrics(src)
    open(dst, 'w').write(text + '\n')
    print(text[:3000])
