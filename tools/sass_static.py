"""Static SASS instruction counts per source function of one kernel.
usage: sass_static.py <nvdisasm -g -c listing of one kernel> <source file> ...
An instruction whose line info points into rt_math.cuh (the inlined arithmetic helpers) is attributed to the most recent
line of one of the given source files, i.e. to its caller."""
import re, collections, sys, bisect, os
srcs = sys.argv[2:]
funcs = {}
for s in srcs:
    starts = []
    for i, l in enumerate(open(s), 1):
        m = re.match(r'\s*(?:RT_HD(?:_COLD)?|__global__|__device__ __forceinline__|static)\s+[\w:<> ]*?\b(\w+)\s*\(', l)
        if m and not l.strip().startswith('//'):
            starts.append((i, m.group(1)))
    funcs[os.path.basename(s)] = starts
def fn(f, n):
    st = funcs.get(f)
    if not st: return f
    k = bisect.bisect_right([a for a, _ in st], n) - 1
    return st[k][1] if k >= 0 else f
ctx = None
cnt = collections.Counter()
for l in open(sys.argv[1]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f = os.path.basename(m.group(1))
        if f in funcs: ctx = fn(f, int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,5}\*/', l):
        cnt[ctx] += 1
print('total', sum(cnt.values()))
for k, c in cnt.most_common(): print('%5d  %s' % (c, k))
