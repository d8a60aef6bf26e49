"""Per-kernel SASS instruction census of libgglb200.so (cuobjdump -sass): which kernels carry the TMA-class bulk copies
(UBLKCP), mbarrier waits (SYNCS), tcgen05 tensor-core / TMEM instructions (UTCHMMA, LDTM, UTCBAR, UTCATOMSWS), vector
reductions (REDG ... F32x4), 64-bit atomics (ATOMG ... MAX.64) and 128-bit shared loads.  Usage:
    python tools/sass_summary.py > profiles/sass_r2_summary.txt"""
import collections
import re
import subprocess
import sys

sys.path.insert(0, '.')
from gammagl_b200 import build  # noqa: E402

PAT = collections.OrderedDict([
    ('UBLKCP', r'\bUBLKCP'), ('UTMALDG', r'\bUTMALDG'), ('SYNCS', r'\bSYNCS'), ('LDS.128', r'\bLDS(\.U)?\.128'),
    ('UTCHMMA', r'\bUTC[A-Z]*MMA'), ('LDTM', r'\bLDTM'), ('UTCBAR', r'\bUTCBAR'), ('UTCATOMSWS', r'\bUTCATOMSWS'),
    ('REDG.F32x4', r'\bREDG\S*F32x4'), ('REDG', r'\bREDG'), ('ATOMG.64', r'\bATOMG\S*\.64'), ('LDG.128', r'\bLDG\S*\.128'),
    ('SHFL', r'\bSHFL'),
])


def strip_params(name):
    """Drop the trailing (parameter list) of a demangled function name; template arguments keep their casts."""
    name = name.strip()
    if not name.endswith(')'):
        return name
    depth = 0
    for i in range(len(name) - 1, -1, -1):
        if name[i] == ')':
            depth += 1
        elif name[i] == '(':
            depth -= 1
            if depth == 0:
                return name[:i]
    return name


def main():
    sass = subprocess.run(['cuobjdump', '-sass', build.LIB], capture_output=True, text=True).stdout
    names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
    blocks = re.split(r'\n\s*Function : ', sass)[1:]
    rows = []
    for name, blk in zip(names, blocks):
        n_inst = len(re.findall(r'^\s+/\*[0-9a-f]{4}\*/', blk, flags=re.M))
        counts = [len(re.findall(p, blk)) for p in PAT.values()]
        short = strip_params(name)
        short = re.sub(r'^(void|int) ', '', short).replace('(int)', '').replace('(bool)', '').replace(' ', '')
        rows.append((short[:110], n_inst, counts))
    print('libgglb200.so: %d kernels; columns = instruction counts in the SASS of each kernel' % len(rows))
    print('%-110s %6s ' % ('kernel', 'insts') + ' '.join('%10s' % k for k in PAT))
    tot = [0] * len(PAT)
    for short, n, c in sorted(rows):
        if any(c[:11]):
            print('%-110s %6d ' % (short, n) + ' '.join('%10d' % v for v in c))
        tot = [a + b for a, b in zip(tot, c)]
    print('%-110s %6s ' % ('TOTAL (all kernels)', '') + ' '.join('%10d' % v for v in tot))


if __name__ == '__main__':
    main()
