"""SASS opcode summary per kernel of libnerfactor_b200.so (evidence that the hot kernels are
Blackwell-native: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit, SYNCS = mbarrier; HMMA / HGMMA would be the legacy tensor paths).

    python tools/sass_summary.py > profiles/r2_sass_summary.md        (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'nerfactor_b200', 'libnerfactor_b200.so')
OPS = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS',
       'FFMA2', 'FMUL2', 'FADD2', 'FFMA', 'MUFU', 'HMMA', 'HGMMA', 'LDGSTS', 'BAR', 'ATOM', 'RED']


def main():
    txt = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)', line)
        if m:
            op = m.group(1)
            kernels[cur]['_total'] += 1
            for o in OPS:
                if op == o or op.startswith(o + '.') or (o in ('UTCHMMA', 'UTCQMMA', 'UTCBAR') and op.startswith(o)):
                    kernels[cur][o] += 1
    demangle = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
    print('# SASS opcode summary, libnerfactor_b200.so (sm_100a), `cuobjdump -sass`\n')
    print('Counts of static instructions per kernel.  `UTCHMMA` = `tcgen05.mma.kind::f16`, `LDTM` / `STTM` = '
          '`tcgen05.ld` / `tcgen05.st`, `UTCBAR` = `tcgen05.commit`, `UBLKCP` = `cp.async.bulk` (1-D TMA), '
          '`SYNCS` = mbarrier, `FFMA2 / FMUL2 / FADD2` = packed fp32.  No `HMMA` / `HGMMA` (legacy tensor '
          'paths) and no `UTMALDG` (tensor-map TMA: the operands are computed in-kernel or arrive by 1-D bulk '
          'copies) anywhere.\n')
    cols = [o for o in OPS if any(k[o] for k in kernels.values())]
    print('| kernel | instr | ' + ' | '.join(cols) + ' |')
    print('|---|---|' + '---|' * len(cols))
    for (name, c), dm in zip(kernels.items(), demangle):
        short = re.sub(r'\(anonymous namespace\)::', '', dm)
        short = re.sub(r'\(.*', '', short).replace('void ', '')
        print('| `%s` | %d | %s |' % (short, c['_total'], ' | '.join(str(c[o]) if c[o] else '' for o in cols)))


if __name__ == '__main__':
    main()
