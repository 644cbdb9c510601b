"""SASS opcode histogram of the built library: python tools/sass_summary.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'ssd-tensorflow_b200', 'libssd_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
WANT = re.compile(r'\s(UTCHMMA[\w.]*|UTMALDG[\w.]*|UTCBAR[\w.]*|UBLKCP[\w.]*|LDTM[\w.]*|UTCATOMSWS[\w.]*|UCGABAR\w*|HMMA[\w.]*|SYNCS[\w.]*)\s')
cur, k = None, -1
total, ops = collections.Counter(), collections.defaultdict(collections.Counter)
for line in sass.split('\n'):
    if 'Function : ' in line:
        k += 1
        cur = re.sub(r'\((ssdb::)?(\(anonymous namespace\)::|<unnamed>::)?\w+\)(?=\d)', '', names[k].replace('ssdb::(anonymous namespace)::', '').replace('ssdb::<unnamed>::', '').replace('void ', '')).split('(')[0]
        continue
    if cur and re.search(r'/\*[0-9a-f]{4,}\*/\s+\S', line):
        total[cur] += 1
        m = WANT.search(line)
        if m:
            op = m.group(1)
            if op.startswith('SYNCS'):
                op = 'SYNCS'
            op = re.sub(r'^(UTCATOMSWS)(\.2CTA)?.*', r'\1\2', op)
            op = re.sub(r'^(LDTM).*', r'\1', op)
            ops[cur][op] += 1
print('# SASS opcode counts of libssd_b200.so (cuobjdump -sass, sm_100a).  tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2), tcgen05.ld -> LDTM,')
print('# tcgen05.commit -> UTCBAR (.2CTA.MULTICAST = multicast commit of a CTA pair), tcgen05.alloc -> UTCATOMSWS, TMA tensor loads -> UTMALDG.nD,')
print('# TMA bulk copies -> UBLKCP, cluster barrier -> UCGABAR, mbarrier -> SYNCS.  No HMMA / mma.sync instruction anywhere.')
for name in total:
    o = {k_: v for k_, v in ops[name].items() if k_ != 'SYNCS'}
    print('%-64s total %6d  %s' % (name[:64], total[name], '  '.join('%s=%d' % kv for kv in sorted(o.items()))))
