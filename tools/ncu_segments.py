"""Aggregate an `ncu --page source --csv --print-source sass` export per barrier-delimited code segment.
The step kernel is a sequence of __syncthreads()-separated phases, so segments map to phases in program order."""
import csv, sys
path = sys.argv[1]; nblk = float(sys.argv[2]) if len(sys.argv) > 2 else 296.0; nsub = float(sys.argv[3]) if len(sys.argv) > 3 else 100.0
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
ia = hdr.index("Source"); isamp = hdr.index("# Samples"); iinst = hdr.index("Instructions Executed"); ithr = hdr.index("Thread Instructions Executed")
ibar = hdr.index("stall_barrier")
segs = []; cur = dict(samp=0, inst=0, thr=0, n=0, ldl=0, ldg=0, bar=0)
ts = ti = 0
for r in data:
    src = r[ia]; s = int(r[isamp]); i = int(r[iinst]); t = int(r[ithr])
    ts += s; ti += i
    cur['samp'] += s; cur['inst'] += i; cur['thr'] += t; cur['n'] += 1; cur['bar'] += int(r[ibar])
    if 'LDL' in src or 'STL' in src: cur['ldl'] += i
    if 'LDG' in src or ' LD.' in src or 'LD.E' in src: cur['ldg'] += i
    if 'BAR.SYNC' in src:
        segs.append(cur); cur = dict(samp=0, inst=0, thr=0, n=0, ldl=0, ldg=0, bar=0)
segs.append(cur)
print(f"total samples {ts} total warp-inst {ti} ({ti/nblk/nsub:.0f} per sub-step per CTA)")
for k, s in enumerate(segs):
    if s['samp'] > 0.003 * ts or s['inst'] > 0.003 * ti:
        print(f"seg {k:3d} sass {s['n']:6d} samp {100*s['samp']/ts:5.1f}% (barrier {100*s['bar']/max(1,s['samp']):3.0f}%) inst {100*s['inst']/ti:5.1f}% ({s['inst']/nblk/nsub:7.0f}/substep/CTA) thr/inst {s['thr']/max(1,s['inst']):5.1f} local {100*s['ldl']/max(1,s['inst']):4.1f}% glob {100*s['ldg']/max(1,s['inst']):4.1f}%")

# Lock-step blocks: samples with stall_barrier sit on the instruction AFTER a BAR.SYNC, i.e. they are the wait for the slowest
# lane of the PREVIOUS interval.  Second table: per interval, own work samples vs the wait that follows it.
print("\nwork vs barrier wait (wait = stall_barrier samples at the start of the next segment, attributed back):")
seg2 = []; cur = dict(work=0, wait=0, inst=0)
for r in data:
    b = int(r[ibar]); cur['work'] += int(r[isamp]) - b; cur['wait'] += b; cur['inst'] += int(r[iinst])
    if 'BAR.SYNC' in r[ia]:
        seg2.append(cur); cur = dict(work=0, wait=0, inst=0)
seg2.append(cur)
tw = sum(x['wait'] for x in seg2)
print(f"barrier wait total {100*tw/ts:.1f}% of samples")
for k in range(len(seg2) - 1):
    w_after = seg2[k + 1]['wait']
    if seg2[k]['work'] + w_after > 0.004 * ts:
        print(f"seg {k:3d} work {100*seg2[k]['work']/ts:5.1f}%  wait for its slowest lane {100*w_after/ts:5.1f}%  inst {seg2[k]['inst']/nblk/nsub:7.0f}/substep/CTA")
