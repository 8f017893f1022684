"""Joins an `ncu --page source --csv` dump with `nvdisasm -g` line info: executed warp instructions per 1 KB block of
SASS with the source lines that dominate each block, and the opcode mix.  Usage: ncu_blocks.py k.sass <kernel substr> src.csv"""
import collections, csv, re, sys
sass, kname, ncsv = sys.argv[1:4]
lines = open(sass).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('//--------------------- .text.') and kname in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('//--------------------- ')), len(lines))
cur, info = None, {}
for l in lines[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        info[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(ncsv)))
hdr = rows[1]
ia, ii, it, isrc = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('Source')
base, data, ops = None, [], collections.Counter()
for r in rows[2:]:
    if r[0] == 'Kernel Name':
        break
    a = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
    base = a if base is None else base
    n = int(r[ii] or 0)
    data.append((a - base, n, int(r[it] or 0)))
    t = r[isrc].split()
    ops[(t[1] if t[0].startswith('@') else t[0]).split('.')[0]] += n
tot = sum(d[1] for d in data)
print(f"kernel {rows[0][1][:70]}\nexecuted warp instructions: {tot}   thread instructions: {sum(d[2] for d in data)}")
agg = collections.OrderedDict()
for off, n, t in data:
    e = agg.setdefault(off // 0x400, [0, 0, collections.Counter()])
    e[0] += n; e[1] += t
    src = info.get(off, (None, ''))[0]
    if src and src[0] != 'rt_math.cuh':
        e[2][src] += n
print("block    share  lanes/inst  dominant source lines (excluding rt_math.cuh helpers)")
for k, e in agg.items():
    if e[0] / tot < 0.004:
        continue
    top = ', '.join(f"{s[0]}:{s[1]}" for s, _ in e[2].most_common(4))
    print(f"{k * 0x400:05x}  {100 * e[0] / tot:5.1f}%  {e[1] / max(e[0], 1):5.1f}      {top}")
print("opcode mix: " + ', '.join(f"{o} {100 * n / tot:.1f}%" for o, n in ops.most_common(16)))
