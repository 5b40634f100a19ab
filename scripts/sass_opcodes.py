"""profiles/rN_sass_opcodes.txt: which Blackwell instructions the built library contains.

    python scripts/sass_opcodes.py [out.txt]

Counts SASS mnemonics of `cuobjdump -sass libhopk.so` (no GPU needed): tcgen05 MMA (UTCHMMA / UTCQMMA / UTCIMMA ...), TMEM
traffic (LDTM / STTM / UTCCP), tensor-core barriers (UTCBAR), bulk copies without a tensor map (UBLKCP), remote st.async (STAS), TMA tensor
copies (UTMALDG / UTMASTG / UTMAPF), mbarrier (SYNCS), cluster barriers (UCGABAR), legacy tensor instructions (HMMA / WGMMA:
must be absent), and per-kernel counts of the tensor-core instructions.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'hop-heterogeneous-topology-based-multimodal-entanglement-for-co-speech-gesture-generation_b200', 'libhopk.so')
WATCH = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTCOMMA', 'LDTM', 'STTM', 'UTCCP', 'UTCBAR', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'UTMAPF',
         'UTMACCTL', 'SYNCS', 'UCGABAR', 'STAS', 'HMMA', 'WGMMA', 'BMMA', 'MUFU.TANH', 'MUFU.EX2', 'REDG', 'ATOMG', 'STG.E.ENL2.256',
         'LDG.E.ENL2.256', 'FENCE.VIEW.ASYNC']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    total = collections.Counter()
    per_kernel = collections.defaultdict(collections.Counter)
    kernel = None
    for ln in out.splitlines():
        m = re.search(r'Function : (\S+)', ln)
        if m:
            kernel = m.group(1)
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
        if not m:
            continue
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                total[w] += 1
                per_kernel[kernel][w] += 1
    lines = [f'# cuobjdump -sass {os.path.basename(LIB)} | opcode counts (sm_100a); HMMA / WGMMA / BMMA must be 0',
             f'# kernels in the library: {len(set(re.findall(r"Function : (\S+)", out)))}']
    for w in WATCH:
        lines.append(f'{w:20s} {total[w]:7d}')
    lines.append('')
    lines.append('# per kernel (demangled prefix): UTCHMMA LDTM STTM UBLKCP UTMALDG STAS SYNCS UCGABAR')
    names = subprocess.run(['c++filt'], input='\n'.join(per_kernel), capture_output=True, text=True).stdout.splitlines()
    for k, nm in zip(per_kernel, names):
        c = per_kernel[k]
        if c['UTCHMMA'] or c['UTMALDG'] or c['UBLKCP'] or c['UCGABAR']:
            short = re.sub(r'\(.*', '', nm).replace('hopk::', '')[:110]
            lines.append(f'{c["UTCHMMA"]:5d} {c["LDTM"]:5d} {c["STTM"]:5d} {c["UBLKCP"]:5d} {c["UTMALDG"]:5d} {c["STAS"]:5d} '
                         f'{c["SYNCS"]:5d} {c["UCGABAR"]:5d}  {short}')
    text = '\n'.join(lines) + '\n'
    if len(sys.argv) > 1:
        open(sys.argv[1], 'w').write(text)
    print(text)


if __name__ == '__main__':
    main()
