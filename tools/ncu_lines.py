"""Map ncu per-instruction samples (--page source --csv) onto source lines using nvdisasm -g line info of the
same cubin (instructions are matched by order within the kernel)."""
import csv, re, sys, collections, subprocess
rep, cubin, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iSmp, iEx = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ins = [(r[iS].strip(), int(r[iSmp] or 0), int(r[iEx] or 0)) for r in rows[2:] if len(r) > iEx]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
func = None; cur = None; lines = []
for ln in dis:
    m = re.match(r'\s*\.text\.(\S+):', ln)
    if m: func = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', ln) and func and kern in func:
        lines.append(cur)
print("ncu instr", len(ins), "nvdisasm instr", len(lines))
n = min(len(ins), len(lines))
smp = collections.Counter(); ex = collections.Counter()
for k in range(n):
    smp[lines[k]] += ins[k][1]; ex[lines[k]] += ins[k][2]
tot = sum(smp.values()); totex = sum(ex.values())
print("total samples", tot, "total warp-instr", totex)
for (f, l), c in smp.most_common(top):
    txt = ''
    for base in ('dolfinx-external-operator_b200/csrc/', '/usr/local/cuda/include/', '/usr/local/cuda/include/crt/'):
        try: txt = open(base + f).read().split('\n')[l - 1].strip()[:80]; break
        except Exception: pass
    print(f"{100*c/tot:5.1f}% smp {100*ex[(f,l)]/totex:5.1f}% ins  {f}:{l}  {txt}")
