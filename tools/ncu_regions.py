"""Instruction and stall-sample shares per source region of pfac_kernels.cu for one kernel of an ncu report
(python tools/ncu_regions.py report.ncu-rep <kernel substring>); see tools/ncu_lines.py."""
import sys, re, csv, io, subprocess, collections, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines
rep, ksub = sys.argv[1], sys.argv[2]
lib=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'pfac_b200', 'lib', 'libpfac.so')
table=ncu_lines.line_table(lib, ksub)
off2line={o:ln for o,ln,_ in table}
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hdr=None;recs=[];base=None
for r in rows:
    if r and r[0]=="Address": hdr=r; continue
    if hdr and len(r)==len(hdr) and r[0].startswith("0x"):
        d=dict(zip(hdr,r)); a=int(d["Address"],16)
        if base is None: base=a
        d["_off"]=a-base; recs.append(d)
src=open(os.path.join(os.path.dirname(lib), '..', 'csrc', 'pfac_kernels.cu')).read().splitlines()
def find(s):
    for i,l in enumerate(src):
        if s in l: return i+1
marks=[("helpers(mbar/tma/ld/st)",1),("comm",find("comm_exchange_scan(const KParams")),("probe_hot/cold",find("__device__ __forceinline__ uint32_t home_bucket")),("stage_tables",find("__device__ __forceinline__ Tables stage_tables")),("prefilter16",find("struct FilterView")),("text_word_slow",find("__device__ __noinline__ uint32_t text_word_slow")),("walk_batch",find("__device__ __forceinline__ int walk_batch")),("elect/clip",find("__device__ __forceinline__ bool elect_one")),("reduce helpers",find("// Reduce kernel: fused match")),("reduce kernel",find("__global__ void __launch_bounds__(kRedThreads, 1) pfac_reduce_kernel")),("matcher",find("// ============================ matcher warps")),("dense kernel",find("// Dense kernel: one persistent CTA of 32 autonomous warps"))]
marks=[m for m in marks if m[1]]; marks.sort(key=lambda m:m[1])
def num(d,k):
    try: return float((d.get(k) or "0").replace(",",""))
    except: return 0.0
agg=collections.defaultdict(lambda:[0.0,0.0,0])
for d in recs:
    ln=off2line.get(d["_off"],-1)
    name=[n for n,s in marks if s<=ln][-1] if ln>0 else "?"
    a=agg[name]; a[0]+=num(d,"Instructions Executed"); a[1]+=num(d,"# Samples"); a[2]+=1
ti=sum(a[0] for a in agg.values()); ts=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print("%-26s inst %5.1f%%  samples %5.1f%%  static %d" % (k, 100*a[0]/ti, 100*a[1]/ts, a[2]))
