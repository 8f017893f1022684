"""Executed warp instructions per source FUNCTION (call chain through the inlined lane code) of one kernel.
Joins `ncu -i X.ncu-rep --page source --csv` with `nvdisasm -gi -c` of the same build: every SASS instruction carries its chain of
inlined frames; frames in rt_math.cuh and the CUDA headers (arithmetic helpers, votes) are folded into their caller.
Usage: ncu_funcs.py <nvdisasm -gi listing> <kernel substr> <src.csv> [depth=3]"""
import bisect, collections, csv, os, re, sys
sass, kname, ncsv = sys.argv[1:4]
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 3
CSRC = os.environ.get("NCU_FUNCS_CSRC") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "se-195-project-ray-tracer_b200", "csrc")
funcs = {}
for f in os.listdir(CSRC):
    starts = []
    for i, l in enumerate(open(os.path.join(CSRC, f), errors="replace"), 1):
        m = re.match(r'\s*(?:template\s*<[^>]*>\s*)?(?:RT_HD(?:_COLD)?|__global__ void(?: __launch_bounds__\([^)]*\))?|__device__ __forceinline__|static)\s+[\w:<> \*&]*?\b(\w+)\s*\(', l)
        if m and not l.strip().startswith('//'):
            starts.append((i, m.group(1)))
    funcs[f] = starts
def fn(f, n):
    st = funcs.get(f)
    if not st: return None
    k = bisect.bisect_right([a for a, _ in st], n) - 1
    return st[k][1] if k >= 0 else None
lines = open(sass).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('//--------------------- .text.') and kname in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('//--------------------- ')), len(lines))
info, chain, fresh = {}, [], True
for l in lines[start:end]:
    ms = re.findall(r'File "([^"]+)", line (\d+)', l)
    if l.lstrip().startswith('//## File') and ms:
        if fresh: chain, fresh = [], False
        for f, n in ms:
            chain.append((os.path.basename(f), int(n)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        info[int(m.group(1), 16)] = list(chain)
        fresh = True
rows = list(csv.reader(open(ncsv)))
hdr = rows[1]
ia, ii, it = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
base = None
agg = collections.OrderedDict()
tot = tott = 0
for r in rows[2:]:
    if r[0] == 'Kernel Name': break
    a = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
    base = a if base is None else base
    n, t = int(r[ii] or 0), int(r[it] or 0)
    names = []
    for f, ln in reversed(info.get(a - base, [])):          # outermost first
        if f == 'rt_math.cuh' or f not in funcs: continue
        name = fn(f, ln)
        if name and (not names or names[-1] != name): names.append(name)
    key = ' > '.join(names[:depth]) or '?'
    e = agg.setdefault(key, [0, 0])
    e[0] += n; e[1] += t
    tot += n; tott += t
print(f"kernel {rows[0][1][:90]}\nexecuted warp instructions {tot}, lanes per instruction {tott / max(tot, 1):.2f}")
for k, e in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if e[0] / tot < 0.002: continue
    print(f"{100 * e[0] / tot:6.2f}%  {e[1] / max(e[0], 1):5.1f} lanes  {k}")
