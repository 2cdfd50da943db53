"""Static SASS instruction count of one refine kernel instance by source function (needs nvdisasm;
run in the build container):  python profiles/tools/static_code.py [mangled-name-substring]"""
import bisect, collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
want = sys.argv[1] if len(sys.argv) > 1 else "ConfigIfLi2ELb1ELi0ELb0ELb0ELb0ELb0E"
obj = os.path.join(ROOT, "clustertracking_b200/csrc/_build/inst_float_0_0.o")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
text = subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=tmp, capture_output=True, text=True).stdout
lines = text.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and want in l and "refine_kernel" in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith(".text.") or lines[i].startswith(".section")), len(lines))
src = open(os.path.join(ROOT, "clustertracking_b200/csrc/ctk_solver.cuh")).read().splitlines()
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r'\s*(?:template <[^>]*>\s*)?(?:CTK_DEV_BIG|CTK_DEV|CTK_COLD)\s+[\w:<>\*& ]+?\s+(\w+)\s*\(', l)
    if m:
        funcs.append((i, m.group(1)))
starts = [f[0] for f in funcs]
cur, count, byline = None, collections.Counter(), collections.Counter()
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l) and cur:
        f, ln = cur
        name = f
        if f == 'ctk_solver.cuh':
            k = bisect.bisect_right(starts, ln) - 1
            name = funcs[k][1] if k >= 0 else '?'
        count[name] += 1
        byline[(f, ln)] += 1
tot = sum(count.values())
print("total", tot, "instructions (%.0f KB)" % (tot * 16 / 1024.))
for n, c in count.most_common(30):
    print("%6d %5.1f%% %s" % (c, 100 * c / tot, n))
print("top lines")
for (f, ln), c in byline.most_common(20):
    print(c, f, ln, src[ln - 1].strip()[:100] if f == 'ctk_solver.cuh' else '')
