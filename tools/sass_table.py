"""Developer tool: static SASS mnemonic counts per kernel of libctcb200.so (cuobjdump -sass), as a markdown table.

    python tools/sass_table.py > profiles/rNN_sass_mnemonics.md
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "asr_chinese_e2e_b200", "libctcb200.so")
COLS = ["UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "USETMAXREG", "DFMA", "DMUL", "DADD", "MUFU", "REDG", "SHFL",
        "STG", "LDG", "BAR"]
txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        counts[kern][m.group(1)] += 1
        counts[kern]["total"] += 1
print("# SASS mnemonic counts per kernel of libctcb200.so (cuobjdump -sass; static counts, sm_100a)\n")
print("| kernel | " + " | ".join(COLS) + " | total |")
print("|---|" + "---|" * (len(COLS) + 1))
rows = []
for k, c in counts.items():
    name = demangle(k)
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("ctcb200::", "")
    rows.append((name, c))
# one representative instantiation per family: the ones the C2 / C4 configurations launch
KEEP = [r"^k0_prep", r"^k4_rescale", r"^k1_lse_gather<128, 9, true, true, false>", r"^k1_lse_gather<64, 17, true, false, false>",
        r"^k1d_sweep<17, true>", r"^k1p_sweep<128, 17, true, (true|false)>", r"^k2_lattice<(4|8), (true|false)>", r"^k3_grad<128, 9, true>", r"^k3p_patch", r"^k5_greedy_cer<4>",
        r"^kce_rows<128, 9, true", r"^k_head", r"^k_gemm3", r"^k_split_tf32", r"^k_sum_partials", r"^k6_edit"]
rows = [r for r in rows if any(re.search(k, r[0]) for k in KEEP)]
for name, c in sorted(rows):
    print(f"| `{name}` | " + " | ".join(str(c.get(col, 0)) for col in COLS) + f" | {c['total']} |")
print("""
`UTCHMMA` = `tcgen05.mma` (kind::tf32: the fused CTC head `k_head` and the parameter-gradient GEMM `k_gemm3`), `UTMALDG` = tensor-map
TMA loads (`cp.async.bulk.tensor.2d`), `LDTM` = `tcgen05.ld` (TMEM -> registers in the epilogues), `UTCBAR` = `tcgen05.commit`,
`USETMAXREG` = `setmaxnreg` (k_gemm3 moves registers from the producer warpgroup to the epilogue warpgroups),
`UBLKCP` = 1-D bulk TMA copy (`cp.async.bulk`: logits frame groups in, and in `k1p_sweep<..,BULKST>` gradient groups out), `SYNCS` = mbarrier operations,
`REDG` = reductions to global memory (the sparse occupancy corrections of `k3p_patch`; the few in the sweep kernels are the
status-word `atomicOr`).  The lattice kernel holds both recursions (float64 `DFMA/DMUL/DADD` fast path, `MUFU` log-space fallback).
Tensor-core instructions appear only in the two GEMM-shaped kernels of section 8f-1; the core CTC path is HBM- and latency-bound.""")
