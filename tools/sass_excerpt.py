"""SASS excerpt of the windowed tick kernel for profiles/: instruction mix, the window staging and the probe group.
    python tools/sass_excerpt.py > profiles/r02_k1w_sass_excerpt.txt      (needs cuobjdump; no GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FN = "_Z17tick_check_kernelILb0ELb1EEv8TickArgs9TickPeers"


def main():
    so = os.path.join(ROOT, "rtpp_b200", "librtpp_gpu.so")
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", FN, so], capture_output=True, text=True).stdout
    ins = []
    for line in txt.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((m.group(1), m.group(2).strip()))
    if not ins:
        sys.exit("no SASS found (is the library built?)")
    mix = collections.Counter()
    for _, t in ins:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        mix[op.split(".")[0]] += 1
    print(f"SASS of tick_check_kernel<kPeers=0,kBulk=1> (cuobjdump -sass rtpp_b200/librtpp_gpu.so, sm_100a), {len(ins)} instructions")
    print("static instruction mix: " + ", ".join("%s %d" % kv for kv in mix.most_common(28)))
    ops = [t for _, t in ins]
    first = next(i for i, t in enumerate(ops) if "SYNCS.ARRIVE.TRANS64" in t)
    last = next(i for i, t in enumerate(ops) if "SYNCS.PHASECHK" in t and i > first)
    print("\n--- window staging: 16 lanes issue one bulk copy each (UBLKCP), the warp waits on its own mbarrier (SYNCS)")
    for a, t in ins[max(0, first - 4): last + 6]:
        print(f"/*{a}*/ {t}")
    p0 = next(i for i, t in enumerate(ops) if "FADD2.RZ" in t)
    print("\n--- probe group (source probe, then the first unrolled step of the k loop): packed f32x2 arithmetic, truncation by")
    print("--- FADD2.RZ + 2^23, one LDS of the shared-memory window word, funnel shift to the probe's bit; LDG.E.64 = the ring of")
    print("--- neighbour rows fetched five steps ahead")
    for a, t in ins[max(0, p0 - 4): p0 + 110]:
        print(f"/*{a}*/ {t}")


if __name__ == "__main__":
    main()
