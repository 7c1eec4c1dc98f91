"""Join an `ncu --page source --csv --print-source sass` export with `nvdisasm -g -c` line info of the same cubin:
hot static code footprint (SASS lines executed by >= a fraction of env-sub-steps) per CUDA source line / function.
usage: ncu_lines.py <src.csv> <dis.txt> <kernel substring e.g. ILi2ELi1E> <n_ctas> <n_substeps> [frac]"""
import csv, re, sys, collections
srcp, disp, kname, nblk, nsub = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), float(sys.argv[5])
frac = float(sys.argv[6]) if len(sys.argv) > 6 else 0.1
# --- disassembly: instruction index -> (file, line)
lines = open(disp).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
loc = []; cur = ("?", 0)
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): loc.append(cur)
rows = list(csv.reader(open(srcp))); hdr = rows[1]; data = rows[2:]
ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
print("sass lines: csv", len(data), "disasm", len(loc))
n = min(len(data), len(loc))
base = nblk * nsub
hot = collections.Counter(); dyn = collections.Counter(); smp = collections.Counter(); stat = collections.Counter()
for k in range(n):
    e = int(data[k][ii]); key = loc[k]
    stat[key] += 1
    if e >= frac * base: hot[key] += 1
    dyn[key] += e; smp[key] += int(data[k][isamp])
tot_hot = sum(hot.values())
print("hot static lines", tot_hot, "=", tot_hot * 16 // 1024, "KB")
# aggregate per file in buckets of source lines
byfile = collections.Counter()
for (f, ln), c in hot.items(): byfile[f] += c
print(dict(byfile))
print("%-22s %6s %6s %9s %7s" % ("source line", "hot", "static", "dyn/substep", "samp%"))
ts = sum(smp.values())
for key, c in sorted(hot.items(), key=lambda x: -x[1])[:70]:
    print("%-16s:%5d %6d %6d %9.1f %6.2f%%" % (key[0], key[1], c, stat[key], dyn[key] / base, 100.0 * smp[key] / ts))
