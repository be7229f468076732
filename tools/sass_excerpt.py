"""SASS evidence of the shipped library: per tensor-core kernel the count of tcgen05 (UTCHMMA, UTCBAR) / tensor-memory (LDTM, STTM) /
TMA (UTMALDG, UBLKCP) / mbarrier (SYNCS) and related mnemonics, with sample lines.
    python tools/sass_excerpt.py > profiles/sass_excerpt_r2.txt        (runs cuobjdump on the CPU box)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'world_modelz_b200', '_C', 'libwm_b200.so')
WANT = ('UTCHMMA', 'UTCBAR', 'UTCATOMSWS', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'SYNCS', 'ELECT', 'MUFU', 'FFMA2', 'FADD2', 'FMUL2',
        'VIMNMX3', 'F2FP', 'USETMAXREG', 'NANOSLEEP')
SAMPLE = ('UTCHMMA', 'UTCBAR', 'UTMALDG', 'UBLKCP', 'LDTM', 'USETMAXREG')


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(demangle(m.group(1)), [])
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
        if m and cur is not None:
            cur.append(re.sub(r'^@!?U?P\d+\s+', '', m.group(1).strip()))
    total = collections.Counter()
    body = []
    for name, ins in kernels.items():
        ops = collections.Counter()
        for i in ins:
            op = i.split()[0]
            if op.startswith(WANT):
                ops[op if op.startswith(('LDTM', 'STTM', 'SYNCS', 'UTMALDG', 'MUFU', 'UTCATOMSWS')) else op.split('.')[0]] += 1
        if not any(k.startswith(('UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP')) for k in ops):
            continue
        body.append(f'\n## {name}\n  instructions: {len(ins)}')
        for k in sorted(ops):
            body.append(f'  {k:<32} {ops[k]}')
            total[k.split('.')[0]] += ops[k]
        body.append('  e.g.')
        for s in SAMPLE:
            body.extend(f'    {i}' for i in [i for i in ins if i.startswith(s)][:2])
    print('# SASS evidence (round 2, final library): cuobjdump -sass world_modelz_b200/_C/libwm_b200.so, sm_100a (tools/sass_excerpt.py)')
    print('# per kernel: instruction count and the tcgen05 (UTCHMMA, UTCBAR) / TMEM (LDTM, STTM) / TMA (UTMALDG, UBLKCP) / mbarrier (SYNCS)')
    print('# mnemonics it contains, with sample lines')
    print('totals over these kernels: ' + ', '.join(f'{k} {v}' for k, v in sorted(total.items())))
    print('\n'.join(body))


if __name__ == '__main__':
    sys.exit(main())
