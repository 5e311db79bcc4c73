"""SASS evidence: per-kernel instruction counts of the mnemonics that matter (cuobjdump -sass of the built library).
   python profiles/tools/sass_mnemonics.py slam-sensor-fusion_b200/csrc/libssf_gpu.so > profiles/r2/sass_mnemonics_v8.txt"""
import collections, re, subprocess, sys

lib = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
want = ["UBLKCP", "SYNCS", "DMMA", "LDG.E.128", "LDG.E.64", "STG.E.128", "ATOMS", "ATOMG", "RED.", "BAR.SYNC", "SHFL", "MUFU",
        "REDUX", "VOTE", "MEMBAR", "ERRBAR", "FENCE"]
print(f"SASS evidence for {lib} (cuobjdump -sass, sm_100a). Per kernel: instruction count and the mnemonics that show")
print("TMA bulk copies (UBLKCP), mbarrier operations (SYNCS), the FP64 tensor-core MMA (DMMA), 128-bit global loads.\n")
tot = collections.Counter()
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", f)
    c = collections.Counter()
    for i in ins:
        for w in want:
            if i.startswith(w):
                c[w] += 1
                tot[w] += 1
    print(f"{name[:120]:120s} n={len(ins):6d} " + " ".join(f"{k}={v}" for k, v in c.items() if v))
print("\nTOTAL " + " ".join(f"{k}={v}" for k, v in tot.items()))
