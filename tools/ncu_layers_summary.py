"""profiles/r2_ncu_conv_split_b64.txt from the raw page of the labelled-layer ncu capture (tools/gpu_evidence.sh):
    python tools/ncu_layers_summary.py gpurun_out/prof_layers_raw_<run>.csv profiles/r2_ncu_conv_split_b64.txt"""
import csv
import sys

sys.path.insert(0, 'tools')
import ncu_summary as S

src, dst = sys.argv[1:3]
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
extra = ['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg.per_second', 'lts__t_bytes.sum',
         'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'launch__cluster_dim_x']
want = S.WANT + [w for w in extra if w not in S.WANT] + sorted(
    h for h in hdr if 'utchmma' in h and 'bf16' in h and 'sparsity_off.avg.pct_of_peak_sustained_elapsed' in h and h not in S.WANT)
# layer_bench runs the layers in network order, each kernel kind three times (--once: 2 warm-ups + 1)
labels = ['%s_%s_launch2' % (L, k) for L in ('conv1_2', 'conv2_2', 'conv4_2') for k in ('fprop', 'dgrad', 'wgrad')]
out = ['# ncu --set full --clock-control none -k regex:conv_tc python tools/layer_bench.py vgg300 64 split conv4_2 conv1_2 conv2_2 --once',
       '# (B200, round 2, final tree; each layer: 2 warm-up launches + 1 per kernel kind; the third launch of each kind is listed; split bf16x3 operands;',
       '#  bare kernels through ssdb_op_conv_bench: the fused 2x2 max-pool epilogue of conv1_2 / conv2_2 is NOT in these launches;',
       '#  conv1_2 fprop / dgrad = resident-filter window mode; conv2_2 / conv4_2 fprop / dgrad = pair mode (launch__cluster_dim_x 2);',
       '#  conv4_2 wgrad = the SM-pair kernel conv_tc_wgrad_r2c2_kernel, conv2_2 wgrad (Cout = 128) = single-CTA rw2)',
       '# the .ncu-rep (70 MB) stayed on the GPU box; this is its raw page']
assert len(data) == 3 * len(labels), len(data)
for j, lab in enumerate(labels):
    r = data[3 * j + 2]
    out.append('--- ' + r[idx['Kernel Name']].split('(')[0] + '   [%s]' % lab)
    for w in want:
        if w in idx and r[idx[w]] not in ('', 'n/a'):
            out.append('  %-92s %16s %s' % (w, r[idx[w]], units[idx[w]]))
open(dst, 'w').write('\n'.join(out) + '\n')
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__cycles_elapsed.avg.per_second', 'launch__cluster_dim_x']
for j, lab in enumerate(labels):
    r = data[3 * j + 2]
    print(lab, [r[idx[k]] for k in keys if k in idx])
