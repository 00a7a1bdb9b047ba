import re, collections, sys
lines = open('/tmp/mc_dis.txt').read().split('\n')
want = sys.argv[1] if len(sys.argv) > 1 else 'mc_kernelILb1'
func=None; cur=None
cnt=collections.Counter(); tot=0
for ln in lines:
    m = re.match(r'\s*\.text\.(\S+):', ln)
    if m: func=m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', ln) and func and want in func:
        cnt[cur]+=1; tot+=1
print("total", tot)
src = {}
for (f,l),c in sorted(cnt.items(), key=lambda x:-x[1])[:int(sys.argv[2]) if len(sys.argv)>2 else 40]:
    txt=''
    for base in ('dolfinx-external-operator_b200/csrc/','/usr/local/cuda/include/','/usr/local/cuda/include/crt/'):
        try:
            txt = open(base+f).read().split('\n')[l-1].strip()[:90]; break
        except Exception: pass
    print(f"{c:5d} {f}:{l}  {txt}")
