"""Instruction-class summary of the in-tree library's SASS, per kernel (cuobjdump -sass): proves which kernels use tcgen05 / TMEM /
TMA / mma.sync / multimem.    python tools/sass_summary.py > profiles/rN_sass_summary.txt"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'invertavatar_b200', 'libinvertavatar_b200.so')
CLASSES = [
    ('UTCHMMA.2CTA', r'\bUTCHMMA\.2CTA'), ('UTCHMMA', r'\bUTCHMMA\b(?!\.2CTA)'), ('UTCBAR(.2CTA.MULTICAST)', r'\bUTCBAR'),
    ('LDTM', r'\bLDTM'), ('UTMALDG(.2CTA)', r'\bUTMALDG'), ('UTMASTG', r'\bUTMASTG'), ('SYNCS', r'\bSYNCS'), ('UCGABAR', r'\bUCGABAR'),
    ('HMMA', r'\bHMMA'), ('MUFU', r'\bMUFU'), ('LDG', r'\bLDG'), ('STG', r'\bSTG'), ('LDS', r'\bLDS'), ('STS', r'\bSTS'),
    ('RED/ATOM', r'\b(RED|ATOMG|ATOMS|ATOM)\b'), ('MULTIMEM', r'MULTIMEM|\.MMA_?ST|UBLKRED|STG\.E\.MC|\bSTMC'), ('SHFL', r'\bSHFL'), ('FFMA', r'\bFFMA'),
]


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    print(f'# {os.path.relpath(LIB, ROOT)}  sha256 {hashlib.sha256(open(LIB, "rb").read()).hexdigest()[:16]}  (cuobjdump -sass, sm_100a)')
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', name)
            cur = kernels.setdefault(name.split('(')[0], collections.Counter())
            continue
        if cur is None or '/*' not in line:
            continue
        body = line.split('*/', 1)[-1]
        if not re.search(r'[A-Z]{3}', body):
            continue
        cur['total'] += 1
        for cname, pat in CLASSES:
            if re.search(pat, body):
                cur[cname] += 1
    cols = [c for c, _ in CLASSES]
    used = [c for c in cols if any(k[c] for k in kernels.values())]
    print('kernel'.ljust(58) + ' '.join(c[:12].rjust(12) for c in ['total'] + used))
    for name, cnt in kernels.items():
        print(name[:57].ljust(58) + ' '.join(str(cnt[c]).rjust(12) for c in ['total'] + used))
    tot = collections.Counter()
    for cnt in kernels.values():
        tot.update(cnt)
    print('ALL'.ljust(58) + ' '.join(str(tot[c]).rjust(12) for c in ['total'] + used))


if __name__ == '__main__':
    sys.exit(main())
