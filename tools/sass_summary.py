"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (B200_PROFILING.md): tcgen05 MMA (UTC*MMA),
TMEM loads (LDTM), TMA tiled loads / stores (UTMALDG / UTMASTG), untiled bulk copies (UBLKCP), cluster launch control
(UCGABAR / CLC), from `cuobjdump -sass lib/libvp3d_b200.so`.
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'dynamic-camera-augmented-videopose3d_b200', 'lib', 'libvp3d_b200.so')
PATTERNS = [('UTC*MMA (tcgen05.mma)', re.compile(r'\bUTC\w*MMA')), ('UTCBAR (tcgen05.commit)', re.compile(r'\bUTCBAR')),
            ('LDTM (tcgen05.ld)', re.compile(r'\bLDTM')), ('UTMALDG (TMA load)', re.compile(r'\bUTMALDG')),
            ('UTMASTG (TMA store)', re.compile(r'\bUTMASTG')), ('UBLKCP (bulk copy)', re.compile(r'\bUBLKCP')),
            ('SYNCS (mbarrier)', re.compile(r'\bSYNCS')), ('UGETNEXTWORKID (clusterlaunchcontrol.try_cancel)', re.compile(r'\bUGETNEXTWORKID')),
            ('HMMA (mma.sync, must be 0)', re.compile(r'\bHMMA'))]


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace('void ', '').replace('(anonymous namespace)::', '')
            name = re.sub(r'\((CUtensorMap_st|vp3d::|float|double|unsigned|long|int|void|__nv|__half|uint4|float4)[^)]*(\)|$).*', '', name)
            name = re.sub(r'\(.*\)$', '', name)
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        for label, pat in PATTERNS:
            if pat.search(line):
                counts[name][label] += 1
        counts[name]['instructions'] += 1 if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line) else 0
    labels = [l for l, _ in PATTERNS]
    print('# cuobjdump -sass %s | per-kernel mnemonic counts (tools/sass_summary.py)' % os.path.relpath(LIB, ROOT))
    print('kernel | ' + ' | '.join(labels) + ' | instructions')
    total = collections.Counter()
    for k, c in counts.items():
        if not any(c[l] for l in labels):
            continue
        print(k + ' | ' + ' | '.join(str(c[l]) for l in labels) + ' | %d' % c['instructions'])
        total.update(c)
    print('TOTAL | ' + ' | '.join(str(total[l]) for l in labels) + ' | %d' % total['instructions'])


if __name__ == '__main__':
    sys.exit(main())
