"""like ncu_lines.py but for a report with several kernels: usage ncu_lines2.py rep kernel-substr dis [top]"""
import csv, re, subprocess, sys, collections
rep, kname, dis = sys.argv[1], sys.argv[2], sys.argv[3]
# find the kernel's id in the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h = rr[0]
kid = [r[h.index('ID')] for r in rr[2:] if kname in r[h.index('Kernel Name')]][0]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{int(kid)+1}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
hdr = rows[hi]; ix = {k: i for i, k in enumerate(hdr)}
sass = [(r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0), float(r[ix["Avg. Threads Executed"]] or 0)) for r in rows[hi + 1:] if len(r) == len(hdr) and r[ix["Address"]] != 'Address']
lines = open(dis).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
cur = ("?", 0); seq = []
for l in lines[start + 1:]:
    if l.startswith("//--------------------- .text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l): seq.append(cur)
print("sass instrs ncu", len(sass), "nvdisasm", len(seq))
agg = collections.defaultdict(lambda: [0, 0])
for (src, ie, smp, thr), loc in zip(sass, seq):
    a = agg[loc]; a[0] += ie; a[1] += smp
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
import glob
srcs = {}
def srcline(f, n):
    if f not in srcs:
        g = glob.glob(f"/root/repo/distributed_crawler_b200/csrc/{f}")
        srcs[f] = open(g[0]).read().splitlines() if g else []
    return srcs[f][n - 1].strip()[:90] if 0 < n <= len(srcs[f]) else ""
bk = collections.defaultdict(lambda: [0, 0])
for loc, a in agg.items():
    k = (loc[0], loc[1] // 10 * 10); bk[k][0] += a[0]; bk[k][1] += a[1]
print("total inst", ti)
for k, a in sorted(bk.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]:
    print(f"{a[0]/ti*100:5.1f}% inst {a[1]/max(ts,1)*100:5.1f}% smp  {k[0]}:{k[1]}  {srcline(k[0], k[1]+1)[:70]}")
