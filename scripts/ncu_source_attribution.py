"""Per-source-line and per-opcode share of the warp instructions one profiled kernel executed.

  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_source.csv
  python scripts/ncu_source_attribution.py X_source.csv [top_lines] [git-rev | -]

The report must come from a build with -lineinfo and a capture with --import-source on.  Source text is looked up in
pharmsol_b200/csrc/device/ by file name — at the given git revision (the commit the profiled build was made from), in the
current tree when none is given, not at all with `-`.
"""
import csv,collections,re,sys
rows=list(csv.reader(open(sys.argv[1])))
import os
srcdir=os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'pharmsol_b200', 'csrc', 'device') + os.sep
files=[(i,r[1]) for i,r in enumerate(rows) if r and r[0]=='File Path']+[(len(rows),None)]
tot=0; perline=collections.Counter(); opc=collections.Counter()
lineops=collections.defaultdict(collections.Counter)
def num(x):
    try: return int(x)
    except: return 0
for (a,f),(b,_) in zip(files,files[1:]):
    hdr=rows[a+2]
    ie=hdr.index('Instructions Executed')
    fn=f.split('/')[-1]
    cur=None
    for r in rows[a+3:b]:
        if len(r)<=ie: continue
        if r[0]!='':
            cur=(fn,int(r[0]))
        else:
            sass=r[3]; n=num(r[ie])
            if not n: continue
            tot+=n; perline[cur]+=n
            m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)',sass)
            op=m.group(2) if m else sass
            op='.'.join(op.split('.')[:2]) if op.startswith(('MUFU','F2F','I2F','F2I')) else op.split('.')[0]
            opc[op]+=n; lineops[cur][op]+=n
print('total',tot)
rev = sys.argv[3] if len(sys.argv) > 3 else None
_cache = {}
def srcline(k):
    if rev == '-': return ''
    try:
        if k[0] not in _cache:
            if rev:
                import subprocess
                _cache[k[0]] = subprocess.run(['git', 'show', rev + ':pharmsol_b200/csrc/device/' + k[0]], capture_output=True, text=True, check=True).stdout.split('\n')
            else:
                _cache[k[0]] = open(srcdir + k[0]).read().split('\n')
        return _cache[k[0]][k[1]-1].strip()[:100]
    except Exception as e: return ''
for k,v in perline.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 25):
    print(f"{100*v/tot:5.1f}% {k[0]}:{k[1]} {srcline(k)} | {dict(lineops[k].most_common(6))}")
print()
for k,v in opc.most_common(32): print(f"{100*v/tot:5.1f}% {k}")
