"""Per-pass trace of a pooled run (HEROAM_TRACE=1 stderr) at a few passes: time, populations."""
import re, sys
rows=[]
for line in open(sys.argv[1]):
    m=re.match(r'pass (\d+) live (\d+) (?:behind (\d+) )?bins ((?:\d+ )+) ?([\d.]+) ms  psteps (\d+) free (\d+)',line)
    if m: rows.append((int(m.group(1)),int(m.group(2)),[int(x) for x in m.group(4).split()],float(m.group(5)),int(m.group(6)),int(m.group(3) or 0)))
starts=[i for i,r in enumerate(rows) if r[0]==1]
rows=rows[starts[-1]:]
print(len(rows),'passes', sum(r[3] for r in rows)/1e3,'s')
every=int(sys.argv[2]) if len(sys.argv)>2 else 100
cum=0
for i,r in enumerate(rows):
    cum+=r[3]
    if i%every==0 or i==len(rows)-1:
        b=r[2]
        print(f"  pass {r[0]:5d} t={cum/1e3:6.2f}s {r[3]:7.2f} ms  n1..4={b[0]}/{b[1]}/{b[2]}/{b[3]} wider={sum(b[4:16])} fl={b[16]}/{b[17]}/{b[18]} lon={b[19]} cores={sum(b[20:29])} drain={b[30:]} psteps={r[4]:.4e} behind={r[5]}")
