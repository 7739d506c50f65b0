"""SASS evidence for the Blackwell-native claims: per-kernel counts of the mnemonics that prove bulk async copies (UBLKCP), mbarrier
traffic (SYNCS), the async-proxy fence, cp.async (LDGSTS), FP64 arithmetic and 16-byte shared-memory accesses.
    python tools/sass_counts.py > profiles/r2_sass_evidence.md          (cuobjdump from the CUDA toolkit; no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "topopteval.jl_b200", "libtopopt_b200.so")
COLS = ("UBLKCP", "SYNCS", "FENCE.VIEW.ASYNC", "LDGSTS", "BAR.SYNC", "DFMA", "DMUL", "DADD", "LDS.128", "STS.128", "SHFL", "RED.E.ADD.F64")
PAT = re.compile(r"\b(" + "|".join(re.escape(c) for c in COLS) + r")\b")
WANT = ("k_spmv_bsr_pipe", "k_ebe_tile", "k_ebe_pipe", "k_ebe_nodes", "k_asm_rows_tet", "k_asm_offdiag", "k_asm_diag", "k_asm_atomic", "k_cgcg_vec", "k_cg_xr", "k_cg_p",
        "k_xchg", "k_elem_energy", "k_tl_")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1); counts[cur] = collections.Counter(); continue
        if cur:
            for t in PAT.findall(ln):
                counts[cur][t] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonic counts per kernel (`cuobjdump -sass topopteval.jl_b200/libtopopt_b200.so`, sm_100a)\n")
    print("UBLKCP = `cp.async.bulk` (bulk async copy global → shared, completes on an mbarrier), SYNCS = mbarrier init / arrive / expect_tx / try_wait,")
    print("FENCE.VIEW.ASYNC = `fence.proxy.async`, LDGSTS = `cp.async`, RED.E.ADD.F64 = `red.global.add.f64` (ATOMIC assembly only).\n")
    print("| kernel | " + " | ".join(COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for mangled, name in zip(counts, names):
        short = name.split("(")[0].replace("void ", "")
        if any(w in short for w in WANT):
            c = counts[mangled]
            print("| `%s` | %s |" % (short, " | ".join(str(c.get(t, 0)) for t in COLS)))


if __name__ == "__main__":
    sys.exit(main())
