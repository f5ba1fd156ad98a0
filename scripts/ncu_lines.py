#!/usr/bin/env python
"""Per-source-line instruction and stall-sample counts of one kernel: joins the SASS page of an .ncu-rep
(`ncu -i rep --page source --csv`) with `nvdisasm -g -c` line info of the cubin the kernel came from.

    python scripts/ncu_lines.py rep.ncu-rep eic-opticks_b200/csrc/phox_engine.sm_100a.cubin '<mangled kernel name>' [top]
"""
import csv, io, re, subprocess, sys, collections

def main():
    rep, cubin, fun = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[1]; col = {n: i for i, n in enumerate(hdr)}
    data = rows[2:]
    base = int(data[0][col["Address"]], 16)
    inst = {}
    for r in data:
        off = int(r[col["Address"]], 16) - base
        inst[off] = (int(r[col["Instructions Executed"]]), int(r[col["# Samples"]]), int(r[col["Thread Instructions Executed"]]), r[col["Source"]].strip())
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(dis) if l.startswith("//---") and (".text." + fun + " ") in l)
    cur = ("?", 0)
    by_line = collections.defaultdict(lambda: [0, 0, 0])
    for l in dis[start + 1:]:
        if l.startswith("//---"): break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
        if m:
            off = int(m.group(1), 16)
            if off in inst:
                a = by_line[cur]; a[0] += inst[off][0]; a[1] += inst[off][1]; a[2] += inst[off][2]
    tot_i = sum(v[0] for v in by_line.values()); tot_s = sum(v[1] for v in by_line.values())
    print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
    by_file = collections.defaultdict(lambda: [0, 0])
    for (f, ln), v in by_line.items(): by_file[f][0] += v[0]; by_file[f][1] += v[1]
    for f, v in sorted(by_file.items(), key=lambda kv: -kv[1][0]): print("%-24s inst %5.1f %%  samples %5.1f %%" % (f, 100. * v[0] / tot_i, 100. * v[1] / tot_s))
    print("--- by line (top %d by instructions)" % top)
    for (f, ln), v in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-22s %5d  inst %5.2f %%  samples %5.2f %%  lanes %4.1f" % (f, ln, 100. * v[0] / tot_i, 100. * v[1] / tot_s, v[2] / max(v[0], 1)))
    print("--- by line (top %d by samples)" % top)
    for (f, ln), v in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-22s %5d  inst %5.2f %%  samples %5.2f %%  lanes %4.1f" % (f, ln, 100. * v[0] / tot_i, 100. * v[1] / tot_s, v[2] / max(v[0], 1)))

if __name__ == "__main__":
    main()
