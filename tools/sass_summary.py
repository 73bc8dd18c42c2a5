#!/usr/bin/env python
"""SASS evidence per object file: counts of the Blackwell-specific mnemonics (tcgen05 MMA = UTCHMMA/UTCQMMA..., TMA tensor loads =
UTMALDG, bulk copies = UBLKCP, TMEM loads/stores = LDTM/STTM, tcgen05 barriers = UTCBAR) next to the legacy warp-level HMMA, per kernel
object of libemmax.so. Runs here (no GPU needed):  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""

import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "LDTM", "STTM", "UTCCP", "HMMA", "LDSM", "SYNCS", "LDGSTS", "ATOM", "RED", "MUFU"]


def main() -> None:
    objs = sorted(glob.glob(os.path.join(ROOT, "build", "csrc", "*.o")))
    if not objs:
        sys.exit("no objects under build/csrc: run `make -C emmax_b200/csrc` first")
    print("# cuobjdump -sass of build/csrc/*.o (sm_100a), instruction counts per mnemonic prefix; one row per kernel")
    for o in objs:
        sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True, check=True).stdout
        print(f"\n== {os.path.basename(o)}")
        kernels = re.split(r"\n\s*Function : ", sass)[1:]
        for k in kernels:
            name = k.split("\n", 1)[0].strip()
            try:
                name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
            except FileNotFoundError:
                pass
            name = re.sub(r"\(.*", "", name)
            n_instr = len(re.findall(r"^\s+/\*[0-9a-f]{4,}\*/", k, flags=re.M))
            counts = {m: len(re.findall(r"\b" + m + r"[.\s;]", k)) for m in MNEMONICS}
            shown = " ".join(f"{m}={c}" for m, c in counts.items() if c)
            print(f"  {name:<70s} instr={n_instr:<6d} {shown}")


if __name__ == "__main__":
    main()
