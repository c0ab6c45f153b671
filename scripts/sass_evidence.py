"""Counts and short excerpts of the Blackwell-specific SASS of the main kernels (from the objects build() leaves in build/).
usage: python scripts/sass_evidence.py > profiles/r02_sass_evidence.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJS = [("stft_v5.o", ["stft_v5_transform_kernelILi10ELi8", "stft_v5_scan_kernelILi10ELi8"]), ("yin32.o", ["yin32_kernel"]),
        ("timedomain.o", ["frame_walk_multi_kernelILi12ELb1ELb1ELb1ELi4"]), ("dtw.o", ["dtw_fill_warp_kernelILi4ELi0"])]
KEYS = ["UBLKCP", "UTMA", "SYNCS", "FADD2", "FMUL2", "FFMA2", "LDGSTS", "LDGDEPBAR", "DEPBAR", "SHFL", "REDUX", "VIMNMX3", "DADD", "DMUL", "DFMA",
        "LDS.128", "STS.128", "LDS.64", "STS.64", "LDG", "STG", "BAR", "WARPSYNC"]
print("# SASS evidence (cuobjdump -sass of the sm_100a objects of this tree)\n")
print("Counts are static instructions of the named kernel.  `UBLKCP.S.G` / `UBLKCP.G.S` = `cp.async.bulk` (TMA bulk copy global -> "
      "shared / shared -> global), `SYNCS.*` = mbarrier arrive / expect_tx / try_wait, `LDGSTS` = `cp.async` (LSU path), "
      "`FADD2 / FMUL2 / FFMA2` = the packed two-wide FP32 instructions of sm_100.\n")
for obj, kernels in OBJS:
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "build", obj)], capture_output=True, text=True).stdout
    for fn in re.split(r"\n\s*Function : ", txt)[1:]:
        name = fn.split("\n")[0]
        if not any(k in name for k in kernels):
            continue
        ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", fn)
        cnt = collections.Counter()
        first = {}
        for a, t in ins:
            op = t.split()[1] if t.startswith("@") else t.split()[0]
            for k in KEYS:
                if op.startswith(k):
                    cnt[k] += 1
                    if k in ("UBLKCP", "UTMA", "SYNCS", "FFMA2", "LDGSTS") and len(first.setdefault(k, [])) < 3:
                        first[k].append(f"/*{a}*/ {t}")
        short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"sonar::\(anonymous namespace\)::", "", short)
        short = re.sub(r"\(.*", "", short).replace("void ", "")
        print(f"## `{short}` ({obj.replace('.o', '.cu')}): {len(ins)} instructions\n")
        print("| " + " | ".join(k for k in KEYS if cnt[k]) + " |")
        print("|" + "---|" * sum(1 for k in KEYS if cnt[k]))
        print("| " + " | ".join(str(cnt[k]) for k in KEYS if cnt[k]) + " |\n")
        for k, lines in first.items():
            print("```")
            for l in lines:
                print(l)
            print("```")
        print()
