"""static look at the hot kernel: registers / stack and the opcode histogram of the walk loop of lic_sample_kernel<1,1,0,0>
(python scripts/sass_loop_stats.py [path/to/libvv_b200.so]); needs only cuobjdump, no GPU"""
import re, collections, subprocess, sys
lib=sys.argv[1] if len(sys.argv) > 1 else __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))), 'vectorvisualization_b200', 'libvv_b200.so')
name="_ZN6vvb20017lic_sample_kernelILi1ELi1ELb0ELb0EEEvNS_9DevParamsE"
res=subprocess.run(["cuobjdump","-res-usage",lib],capture_output=True,text=True).stdout.splitlines()
i=next(k for k,l in enumerate(res) if name in l); print(res[i+1].strip())
sass=subprocess.run(["cuobjdump","-sass","-fun",name,lib],capture_output=True,text=True).stdout
ins=[]
for l in sass.splitlines():
    m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1),16), m.group(2).strip()))
loops=[]
for a,t in ins:
    m=re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(?:P\d,\s*)?(0x[0-9a-f]+)', t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<a: loops.append(((a-tgt)//16+1,tgt,a))
loops.sort(reverse=True)
def op(t):
    t=re.sub(r'^@!?U?P\d\s+','',t)
    return t.split()[0].split('.')[0]
n,lo,hi=loops[1]
body=[t for ad,t in ins if lo<=ad<=hi]
c=collections.Counter(op(t) for t in body)
print("total", len(ins), "loop", n, dict(c.most_common(30)))
