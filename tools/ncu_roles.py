"""Per-role summary of an `ncu --page source --csv` dump of vq_assign_tc_kernel: instructions and stall samples between
user-given SASS line boundaries, plus the hottest instructions.   python tools/ncu_roles.py src.csv [min_frac]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed'); iSrc = hdr.index('Source')
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iS]) for r in data)
print('total samples', tot, 'total warp instructions', sum(int(r[iI]) for r in data))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
# role boundaries: first wait of each role is a SYNCS.PHASECHK...TRYWAIT; print markers so ranges can be read off
for n, r in enumerate(data):
    s = int(r[iS]); src = r[iSrc].strip()
    mark = any(k in src for k in ('UTMALDG', 'UTCHMMA', 'LDTM', 'BAR.SYNC', 'EXIT', 'ATOM', 'RED.', 'STG', 'SHFL', 'VOTE', 'REDUX')) or ('SYNCS' in src and 'EXCH' not in src)
    if s >= tot * frac or mark:
        st = sorted([(int(r[i]), h[6:]) for i, h in stalls], reverse=True)[:2]
        print(n, s, r[iI], src[:80], st if s else '')
