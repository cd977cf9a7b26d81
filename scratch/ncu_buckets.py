"""per-kernel 20-line source buckets from a (multi-kernel) ncu report + nvdisasm -g listing"""
import subprocess, csv, re, collections, sys
rep, dis = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
lines=open(dis).read().splitlines()
def disasm_locs(kname):
    start=next(i for i,l in enumerate(lines) if l.startswith(".text.") and kname in l)
    cur=("?",0); seq=[]
    for l in lines[start+1:]:
        if l.startswith("//--------------------- .text."): break
        m=re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur=(m.group(1).split("/")[-1], int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", l): seq.append(cur)
    return seq
srcs={}
def srcline(f,n):
    import glob
    if f not in srcs:
        g=glob.glob(f"/root/repo/distributed_crawler_b200/csrc/{f}")
        srcs[f]=open(g[0]).read().splitlines() if g else []
    return srcs[f][n-1].strip()[:90] if 0<n<=len(srcs[f]) else ""
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
secs=[]; cur=None
for r in rows:
    if r and r[0]=="Kernel Name": cur={"name":r[1],"rows":[]}; secs.append(cur); continue
    if cur is not None: cur["rows"].append(r)
for sec in secs:
    hdr=sec["rows"][0]; ix={h:i for i,h in enumerate(hdr)}
    sass=[(int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)) for r in sec["rows"][1:] if len(r)==len(hdr)]
    kname=re.match(r"(?:tgi::)?(\w+)", sec["name"]).group(1)
    seq=disasm_locs(kname+"E")
    agg=collections.defaultdict(lambda:[0,0])
    for (ie,smp),loc in zip(sass,seq):
        agg[loc][0]+=ie; agg[loc][1]+=smp
    ti=sum(a[0] for a in agg.values())
    print(f"== {kname}: {len(sass)} SASS, {ti/1e6:.1f}M warp-instr")
    bk=collections.defaultdict(int)
    for loc,a in agg.items(): bk[(loc[0],loc[1]//20*20)]+=a[0]
    for k,v in sorted(bk.items(), key=lambda kv:-kv[1])[:top]:
        print(f"  {v/ti*100:5.1f}% {v/1e6:7.1f}M  {k[0]}:{k[1]}  {srcline(k[0],k[1]+1)[:70]}")
