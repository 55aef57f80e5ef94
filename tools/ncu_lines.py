"""Join an ncu SASS source page (csv) with nvdisasm line info: instructions executed and stall samples per source line.

    cuobjdump -xelf all smrf_b200/libsmrfb.so ; nvdisasm -g -c dk.sm_100a.cubin > dk.sass
    ncu -i report.ncu-rep --page source --csv > page.csv
    python tools/ncu_lines.py page.csv dk.sass <mangled kernel substring> [top]
"""
import csv, re, sys
page, sass, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(page)))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
h = rows[hi]
ia, ie, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
ins = []
for r in rows[hi + 1:]:
    if len(r) > ie and r[ia].startswith("0x"):
        ins.append((int(r[ia], 16), int(r[ie]), int(r[isamp])))
base = ins[0][0]
# nvdisasm: sections per function; //## File "...", line N  markers precede instructions  /*0010*/
line_of = {}
cur = None
infn = False
for l in open(sass):
    if l.startswith("\t.section") or l.startswith(".section"):
        infn = kern in l and ".text." in l
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
agg = {}
tot = tots = 0
for a, e, s in ins:
    k = line_of.get(a - base, ("?", 0))
    x = agg.setdefault(k, [0, 0])
    x[0] += e
    x[1] += s
    tot += e
    tots += s
print("total warp instructions %d, samples %d" % (tot, tots))
srcs = {}
for (f, ln), (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    try:
        if f not in srcs:
            import glob
            cand = glob.glob("/root/repo/smrf_b200/csrc/" + f)
            srcs[f] = open(cand[0]).read().split("\n") if cand else []
        text = srcs[f][ln - 1].strip()[:100] if srcs[f] else ""
    except Exception:
        pass
    print("%5.1f%% instr %5.1f%% samples  %s:%d  %s" % (100.0 * e / tot, 100.0 * s / max(1, tots), f, ln, text))
