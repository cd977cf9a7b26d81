"""Join an ncu SASS source-page CSV with nvdisasm -g line info: per-CUDA-line instruction counts and
stall samples.  usage: ncu_lines.py <report.ncu-rep> <kernel-mangled-substring> <cubin.dis>"""
import csv, re, subprocess, sys, collections
rep, kname, dis = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
sass = [(r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0), float(r[ix["Avg. Threads Executed"]] or 0)) for r in rows[2:] if len(r) == len(hdr)]
# nvdisasm: walk the function, track current line
lines = open(dis).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
cur = ("?", 0); seq = []
for l in lines[start + 1:]:
    if l.startswith("//--------------------- .text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m: seq.append(cur)
print("sass instrs ncu", len(sass), "nvdisasm", len(seq))
agg = collections.defaultdict(lambda: [0, 0, 0.0])
for (src, ie, smp, thr), loc in zip(sass, seq):
    a = agg[loc]; a[0] += ie; a[1] += smp; a[2] += ie * thr
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total inst", ti, "samples", ts)
srcs = {}
def srcline(f, n):
    import glob
    if f not in srcs:
        g = glob.glob(f"/root/repo/distributed_crawler_b200/csrc/{f}")
        srcs[f] = open(g[0]).read().splitlines() if g else []
    return srcs[f][n - 1].strip()[:100] if 0 < n <= len(srcs[f]) else ""
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    print(f"{a[0]/ti*100:5.1f}% inst {a[1]/ts*100:5.1f}% smp thr {a[2]/max(a[0],1):4.1f}  {loc[0]}:{loc[1]}  {srcline(*loc)}")

# bucket view: instructions per (file, 25-line bucket)
print("---- buckets (>=1.5% of instructions)")
bk = collections.defaultdict(lambda: [0, 0])
for loc, a in agg.items():
    k = (loc[0], loc[1] // 25 * 25)
    bk[k][0] += a[0]; bk[k][1] += a[1]
for k, a in sorted(bk.items(), key=lambda kv: -kv[1][0]):
    if a[0] / ti > 0.015:
        print(f"{a[0]/ti*100:5.1f}% inst {a[1]/ts*100:5.1f}% smp  {k[0]}:{k[1]}-{k[1]+24}  {srcline(k[0], k[1]+1)[:60]}")
