"""SASS instruction counts per kernel of the built library -> profiles/<name>: python scripts/sass_summary.py profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "nllssolver.jl_b200", "libnlls_b200.so")
KEYS = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "SHFL", "BAR"]
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
out, name, cnt = [], None, None
def flush():
    if name is None: return
    dm = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dm = re.sub(r"\(.*", "", dm)
    tot = sum(cnt.values())
    out.append(f"{dm[:70]:70s} total {tot:6d} " + " ".join(f"{k} {cnt[k]}" for k in KEYS if cnt[k]))
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush(); name, cnt = m.group(1), collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cnt is not None:
        op = m.group(1)
        cnt["_all"] += 0
        cnt[op if op in KEYS else "_other"] += 1
flush()
arch = subprocess.run(["cuobjdump", "-lelf", SO], capture_output=True, text=True).stdout
with open(sys.argv[1], "w") as f:
    f.write("# SASS instruction counts per kernel of nllssolver.jl_b200/libnlls_b200.so (cuobjdump -sass; ELF: " + ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))) + "); names demangled\n")
    f.write("# DMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64), UBLKCP = TMA bulk copy (cp.async.bulk), SYNCS = mbarrier ops, LDGSTS = cp.async\n")
    f.write("# there is no UTC*MMA / TMEM: tcgen05.mma has no f64 kind, so DMMA + bulk copies are the Blackwell-native path for an FP64 solver\n\n")
    f.write("\n".join(out) + "\n")
print(len(out), "kernels")
