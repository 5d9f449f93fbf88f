"""Join an ncu SASS source page (csv) with nvdisasm -g line info: executed instructions and stall samples per source line.
usage: ncu_lines.py <ncu source csv> <cubin> <kernel substring> [top N]"""
import csv, re, subprocess, sys
src_csv, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, infun = [], None, False
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        infun = kern in ln
        continue
    if not infun:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
h = rows[1]
iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
body = rows[2:]
assert len(body) == len(lines), (len(body), len(lines))
agg = {}
tot_e = tot_s = 0
for r, l in zip(body, lines):
    e, s = int(r[iE] or 0), int(r[iN] or 0)
    a = agg.setdefault(l, [0, 0]); a[0] += e; a[1] += s
    tot_e += e; tot_s += s
print("total instructions %d samples %d" % (tot_e, tot_s))
srcfiles = {}
for (f, n), (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = ""
    try:
        if f not in srcfiles:
            import glob
            c = glob.glob("/root/repo/**/" + f, recursive=True)
            srcfiles[f] = open(c[0]).read().splitlines() if c else []
        text = srcfiles[f][n - 1].strip()[:90]
    except Exception:
        pass
    print("%-12s %4d  inst %5.1f%%  samples %5.1f%%  %s" % (f, n, 100.0 * e / tot_e, 100.0 * s / tot_s, text))
