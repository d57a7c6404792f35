"""Per-phase breakdown of the dense shade kernel's ncu profile (development aid).
   python tools/ncu_phases.py <report> <lib> <kernel> <line of 'if (s >= 0) {' in surface_round> <line of regen 'if ((int)threadIdx.x < take)'>"""
import sys
rep, lib, kn, l_surf, l_regen = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
sys.argv = ['x', rep, lib, kn, '0']
import os
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ncu_lines.py')).read().split('print("total warp-inst')[0]
exec(src)
surf = [i for i, loc in enumerate(lines) if loc == ('kernels.cu', l_surf)]
reg = [i for i, loc in enumerate(lines) if loc == ('kernels.cu', l_regen)]
dstart = surf[0]; sstart = [i for i in surf if i > dstart + 2000][0]; rstart = reg[0]
stall_cols = [c for c in rows[0].keys() if c.startswith('stall_') and 'Not Issued' not in c]
N = len(rows)
marks = sorted([0, dstart, sstart, rstart, N]); names = {0: 'pre/A', dstart: 'D', sstart: 'S', rstart: 'R+tail'}
for a, b in zip(marks[:-1], marks[1:]):
    ie = sum(num(r["Instructions Executed"]) for r in rows[a:b]); te = sum(num(r["Thread Instructions Executed"]) for r in rows[a:b]); sm = sum(num(r["# Samples"]) for r in rows[a:b])
    st = {c: sum(num(r[c]) for r in rows[a:b]) for c in stall_cols}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
    print("%-7s sass %5d inst %5.1f%% samples %5.1f%% lanes %4.1f | %s" % (names[a], b - a, 100 * ie / tot[0], 100 * sm / tot[2], te / max(ie, 1), " ".join("%s %.0f%%" % (k[6:], 100 * v / max(sm, 1)) for k, v in top)))
