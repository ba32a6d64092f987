"""SASS evidence file: python profiles/sass_excerpt.py [libtqsim.so] > profiles/sass_r02_excerpt.txt   (needs cuobjdump, no GPU)
Per kernel: counts of the TMA / mbarrier / tensor-core mnemonics and the instructions themselves (first DMMAs only)."""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "tensorrl_qas_b200/libtqsim.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
print("# SASS evidence, round 2: `cuobjdump -sass tensorrl_qas_b200/libtqsim.so` (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -lineinfo).")
print("# tile_stream_kernel<MODE> (tq_stream.cu): tile loads are UTMALDG.5D (cp.async.bulk.tensor, completion on an mbarrier:")
print("# SYNCS.ARRIVE.TRANS64 = arrive.expect_tx, SYNCS.PHASECHK.TRANS64.TRYWAIT = try_wait.parity), tile stores UTMASTG.5D")
print("# (+ UTMACMDFLUSH = commit_group, DEPBAR-style wait on the bulk group), the dense blocks DMMA.8x8x4 (mma.sync.m8n8k4.f64;")
print("# tcgen05 has no FP64 kind).  Counts per kernel, then the instructions themselves with their addresses.")
KEYS = ["UTMALDG", "UTMASTG", "UTMACMDFLUSH", "SYNCS", "DMMA", "DFMA", "DMUL", "LDG", "STG", "LDS", "STS", "BAR", "SHFL"]
name, body = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        body[name] = []
    elif name and re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+\S", line):
        body[name].append(line.rstrip())
for name, lines in body.items():
    if not re.search(r"tile_stream_kernel|tile_pass_mma_kernel|expect_direct_kernel", name):
        continue
    cnt = {k: sum(1 for l in lines if re.search(r"\b" + k + r"\b|\b" + k + r"\.", l)) for k in KEYS}
    print(f"\n== {name}")
    print("   instruction counts: " + ", ".join(f"{k} {v}" for k, v in cnt.items()))
    if "tile_stream_kernel" not in name:
        continue
    n_dmma = 0
    for l in lines:
        if re.search(r"UTMALDG|UTMASTG|UTMACMDFLUSH|SYNCS|FENCE\.VIEW\.ASYNC", l):
            print(l)
        elif "DMMA" in l and n_dmma < 4:
            n_dmma += 1
            print(l)
