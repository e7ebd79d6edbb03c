#!/usr/bin/env python
"""Committed SASS evidence (north_star: "with committed SASS listings"): per-kernel counts of the Blackwell
tensor-core / TMEM / TMA instructions in librovr_b200.so and an excerpt of the MMA-issue loop and of the
epilogue of the main igemm instance. Runs on the CPU box (cuobjdump only reads the cubin).

    python scripts/dump_sass.py > profiles/sass_igemm.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "reinformcement-optimized-video-reconstruction_b200", "librovr_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "UTCCP", "ELECT", "HMMA", "IMMA",
             "UCGABAR", "ACQBULK", "UBLKCP"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    arch = re.search(r"arch = (sm_[0-9a-z]+)", sass)
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass), arch: {arch.group(1) if arch else '?'}")
    print("# columns: instructions | " + " ".join(MNEMONICS))
    rows = []
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/", l)]
        counts = [sum(1 for l in lines if re.search(r"\b" + m, l)) for m in MNEMONICS]
        rows.append((demangled, len(lines), counts, lines))
    rows.sort(key=lambda r: -r[2][0])
    for name, n, counts, _ in rows:
        if counts[0] or counts[2] or counts[4] or counts[5]:
            short = re.sub(r"\(.*", "", name)
            print(f"{short[:110]:110s} {n:6d} | " + " ".join(f"{c:5d}" for c in counts))
    tot = [sum(r[2][i] for r in rows) for i in range(len(MNEMONICS))]
    print(f"{'TOTAL (all ' + str(len(rows)) + ' kernels)':110s} {sum(r[1] for r in rows):6d} | " + " ".join(f"{c:5d}" for c in tot))
    assert tot[MNEMONICS.index("HMMA")] == 0 and tot[MNEMONICS.index("IMMA")] == 0, "legacy mma.sync found"
    # excerpt: the plain single-CTA igemm instance
    for want, title in (("igemm_kernel<true, false, false, false>", "igemm_kernel<plain epilogue, single CTA>"),
                        ("igemm_kernel<true, false, false, true>", "igemm_kernel<plain epilogue, CTA pair (cta_group::2)>"),
                        ("wgrad_halo_kernel", "wgrad_halo_kernel")):
        cand = [r for r in rows if want.replace(" ", "") in r[0].replace(" ", "").replace("(bool)1", "true").replace("(bool)0", "false")]
        if not cand:
            continue
        name, n, counts, lines = cand[0]
        idx = [i for i, l in enumerate(lines) if "UTCHMMA" in l]
        print(f"\n## {title}: {n} instructions, first UTCHMMA group (MMA issue, one elected thread)")
        lo, hi = max(0, idx[0] - 12), min(len(lines), idx[0] + 28)
        for l in lines[lo:hi]:
            print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())
        ld = [i for i, l in enumerate(lines) if "LDTM" in l]
        if ld:
            print(f"## {title}: first LDTM group (epilogue: TMEM -> registers)")
            for l in lines[max(0, ld[0] - 4):ld[0] + 14]:
                print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())
        tm = [i for i, l in enumerate(lines) if "UTMALDG" in l]
        if tm:
            print(f"## {title}: first UTMALDG group (TMA producer)")
            for l in lines[max(0, tm[0] - 6):tm[0] + 6]:
                print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())
        st = [i for i, l in enumerate(lines) if "UTMASTG" in l]
        if st:
            print(f"## {title}: first UTMASTG (TMA store of the output tile)")
            for l in lines[max(0, st[0] - 3):st[0] + 3]:
                print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())


if __name__ == "__main__":
    main()
