"""Attribute an ncu per-SASS-instruction profile to CUDA source lines / inlined functions (development aid).

    python tools/ncu_lines.py <report.ncu-rep> <library.so> <kernel-mangled-substring> [top]

ncu's `--page source --csv` lists every SASS instruction of the kernel with executed counts and stall samples;
`nvdisasm -g` of the same cubin lists the same instructions with `//## File "...", line N` markers.  The two are
joined by instruction order (the counts must agree, so profile and library have to be the same build).
"""
import csv, collections, os, re, subprocess, sys, tempfile

rep, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
lines = None
for f in sorted(os.listdir(tmp)):
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    m = re.search(r"^\.text\.[^\n]*%s[^\n]*:\n" % re.escape(kname), txt, re.M)
    if not m:
        continue
    body = txt[m.end():]
    end = re.search(r"^//-+ \.", body, re.M)
    body = body[:end.start()] if end else body
    cur = ("?", 0); lines = []
    for l in body.splitlines():
        mm = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2))); continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    break
assert lines, "kernel not found in " + lib
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(out) if l.startswith('"Address"'))
rows = list(csv.DictReader(out[start:]))
assert len(rows) == len(lines), (len(rows), len(lines))
def num(x):
    try: return float(x)
    except ValueError: return 0.0
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
fagg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
tot = [0.0, 0.0, 0.0]
for r, loc in zip(rows, lines):
    v = (num(r["Instructions Executed"]), num(r["Thread Instructions Executed"]), num(r["# Samples"]))
    for a in (agg[loc], fagg[loc[0]]):
        a[0] += v[0]; a[1] += v[1]; a[2] += v[2]; a[3] += 1
    for i in range(3): tot[i] += v[i]
print("total warp-inst %.3g  thread-inst %.3g  samples %d  sass lines %d" % (tot[0], tot[1], tot[2], len(rows)))
print("-- by file: share of warp-inst | share of stall samples | avg active lanes | sass lines")
for k, a in sorted(fagg.items(), key=lambda kv: -kv[1][0]):
    print("  %-22s %5.1f%%  %5.1f%%  %5.1f  %6d" % (k, 100 * a[0] / tot[0], 100 * a[2] / max(tot[2], 1), a[1] / max(a[0], 1), a[3]))
print("-- top lines by stall samples")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print("  %-22s:%-5d inst %5.1f%%  samples %5.1f%%  lanes %5.1f  sass %5d" % (k[0], k[1], 100 * a[0] / tot[0], 100 * a[2] / max(tot[2], 1), a[1] / max(a[0], 1), a[3]))
