"""Dynamic instruction / stall-sample counts per CUDA source line: joins `ncu --page source --csv` (per-SASS-instruction
executed counts) with `nvdisasm --print-line-info` of the same cubin on the instruction offset.
usage: python scripts/ncu_by_line.py <ncu_source.csv> <nvdisasm.sass> <kernel-name-substring> <units (frames) per launch> [top]"""
import csv, re, sys, collections

src_csv, sass, kname, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
# offset -> (outermost file:line in the kernel's own file, innermost file:line)
loc, cur, on = {}, None, False
for ln in open(sass):
    if ln.startswith("//---") and ".text." in ln:
        on = kname in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        chain = [(m.group(1).split("/")[-1], int(m.group(2)))]
        for mm in re.finditer(r'inlined at "([^"]+)", line (\d+)', m.group(3)):
            chain.append((mm.group(1).split("/")[-1], int(mm.group(2))))
        cur = chain
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and cur:
        loc[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isrc, ie, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[2][ia], 16)
outer, inner = collections.Counter(), collections.Counter()
osamp = collections.Counter()
tot = 0
for r in rows[2:]:
    try:
        n = int(r[ie]); off = int(r[ia], 16) - base
    except Exception:
        continue
    tot += n
    ch = loc.get(off, ([("?", 0)], ""))[0]
    own = [c for c in ch if c[0].startswith("stft") or c[0].startswith("yin") or c[0].startswith("dtw")]
    key = own[-1] if own else ch[-1]   # outermost location in the kernel's own file
    outer[key] += n
    osamp[key] += int(r[isamp] or 0)
    inner[ch[0]] += n
print(f"total {tot / units:.1f} inst/unit")
print("-- by outermost own-file line (inst/unit, stall samples)")
for k, v in outer.most_common(top):
    print(f"{k[0]}:{k[1]:<5d} {v / units:8.1f}  {osamp[k]}")
