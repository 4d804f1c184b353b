"""Static census of the built libwmk.so (no GPU needed): per kernel family the register / stack / static shared-memory
footprint (`cuobjdump -res-usage`) and the Blackwell instruction mix of its SASS (`cuobjdump -sass`): tcgen05 MMAs
(UTCHMMA), TMA loads / stores (UTMALDG / UTMASTG), TMEM loads (LDTM), tcgen05 commits (UTCBAR), Ampere-style HMMA,
local-memory spills (LDL / STL), packed FMAs.  Template instances of one kernel are merged into one row (ranges shown).

    python tools/sass_census.py > profiles/r02_sass_census.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "image-in-speech-watermarking_b200", "csrc", "libwmk.so")
MNEMONICS = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "HMMA", "FFMA2", "FHFMA", "MUFU", "LDL", "STL", "ATOMG", "REDG", "RED")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def family(demangled):
    """`wmk::(anonymous namespace)::gemm_kernel<128, true>(Args)` -> `gemm_kernel`."""
    s = re.sub(r"\(anonymous namespace\)::", "", demangled)
    s = re.sub(r"^void\s+", "", s)
    s = s.split("(")[0]
    s = re.sub(r"<.*$", "", s)
    return s.replace("wmk::", "")


def rng(vals):
    lo, hi = min(vals), max(vals)
    return str(lo) if lo == hi else "%d-%d" % (lo, hi)


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line)}
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.defaultdict(collections.Counter)
    n_inst = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if cur and m:
            op = m.group(1)
            n_inst[cur] += 1
            for k in MNEMONICS:
                if op == k or (k == "HMMA" and op.startswith("HMMA")):
                    counts[cur][k] += 1
    names = sorted(usage)
    dm = demangle(names)
    fam = collections.defaultdict(list)
    for n in names:
        fam[family(dm[n])].append(n)
    print("# Static census of `libwmk.so` (sm_100a SASS; `python tools/sass_census.py`)\n")
    print("%d kernels (template instances) in %d families.  REG = registers per thread, STACK = per-thread stack bytes "
          "(address-taken locals), SPILL = LDL + STL instructions in the SASS, SMEM = static shared memory (dynamic shared memory "
          "is set at launch and not shown); instruction columns are static counts summed over the family's instances.\n"
          % (len(names), len(fam)))
    tot = collections.Counter()
    print("| kernel family | instances | REG | STACK | SMEM (static) | SASS instr. | UTCHMMA | UTMALDG | UTMASTG | LDTM | UTCBAR | HMMA | FFMA2 | FHFMA | MUFU | SPILL |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for f in sorted(fam, key=lambda f: -sum(n_inst[n] for n in fam[f])):
        ns = fam[f]
        c = collections.Counter()
        for n in ns:
            c.update(counts[n])
        tot.update(c)
        cell = lambda k: str(c[k]) if c[k] else ""
        print("| `%s` | %d | %s | %s | %s | %d | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
            f, len(ns), rng([usage[n]["REG"] for n in ns]), rng([usage[n]["STACK"] for n in ns]),
            rng([usage[n]["SHARED"] for n in ns]), sum(n_inst[n] for n in ns), cell("UTCHMMA"), cell("UTMALDG"), cell("UTMASTG"),
            cell("LDTM"), cell("UTCBAR"), cell("HMMA"), cell("FFMA2"), cell("FHFMA"), cell("MUFU"),
            str(c["LDL"] + c["STL"]) if c["LDL"] + c["STL"] else ""))
    print("\nTotals: " + ", ".join("%s %d" % (k, tot[k]) for k in MNEMONICS if tot[k]) + ".")


if __name__ == "__main__":
    sys.exit(main())
