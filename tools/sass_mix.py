"""Instruction mix and proof-of-architecture mnemonics per kernel from the built library's SASS
(cuobjdump -sass) -> markdown (profiles/r02_sass.md).   python tools/sass_mix.py > profiles/r02_sass.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "amcontrast3d_b200", "lib", "libamc3d_sm100a.so")
WANT = ["knn_wq_kernel<1, false>", "knn_tq_kernel<3, false>", "ball_wq_kernel<1, false>", "amloss_forward_kernel<16, 1, 0>",
        "fps_cluster_kernel<16, 12, 4, false>", "fps_culled_kernel<false>", "group_fwd_tma_kernel<4>", "group_bwd_tma_kernel<true>",
        "fused_sa_fwd_kernel<32, false>", "fused_sa_fwd_kernel<32, true>", "fused_sa_bwd_scatter_kernel", "ffma_probe_kernel"]
GROUPS = [("tensor core (tcgen05.mma / ld / alloc / commit)", r"^(UTC\w*MMA|LDTM|STTM|UTCBAR|UTCATOMSWS|UTCCP)"),
          ("TMA / bulk async / cp.async", r"^(UTMALDG|UTMASTG|UBLKCP|LDGSTS|UBLKRED)"),
          ("mbarrier / cluster (SYNCS, UCGABAR, STAS)", r"^(SYNCS|UCGABAR|STAS|ARRIVES)"),
          ("FP32 math (FFMA FADD FMUL FMNMX FSETP FSEL)", r"^(FFMA|FADD|FMUL|FMNMX|FSETP|FSEL|FCHK|MUFU)"),
          ("integer / logic", r"^(IMAD|IADD3|LOP3|SHF|LEA|ISETP|SEL|VIADD|VIMNMX|PRMT|POPC|FLO|BREV|IABS|I2F|F2I|UIADD3|ULOP3|UISETP|USHF|UMOV|UIMAD|ULEA|USEL|MOV|S2R|S2UR|CS2R|R2UR|LDC|ULDC|LDCU)"),
          ("warp exchange (SHFL VOTE REDUX MATCH)", r"^(SHFL|VOTE|REDUX|MATCH|WARPSYNC)"),
          ("global loads / stores (LDG STG)", r"^(LDG|STG|LD\b|ST\b)"),
          ("global reductions / atomics (REDG RED ATOMG)", r"^(REDG|RED\b|ATOMG|ATOM\b)"),
          ("shared memory (LDS STS ATOMS LDSM)", r"^(LDS|STS|ATOMS|LDSM)"),
          ("control (BRA BSSY BSYNC BAR EXIT CALL)", r"^(BRA|BSSY|BSYNC|BAR|EXIT|CALL|RET|YIELD|NANOSLEEP|DEPBAR|ERRBAR|MEMBAR|FENCE|CCTL)")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        funcs[cur].append(m.group(1))
names = demangle(list(funcs))
print("# SASS of the shipped library (`cuobjdump -sass amcontrast3d_b200/lib/libamc3d_sm100a.so`), per kernel\n")
print("Counts are static instructions in the kernel's SASS, grouped; the mnemonics that prove the Blackwell paths are listed "
      "verbatim (B200_PROFILING.md: `tcgen05.mma` -> `UTC*MMA`, `tcgen05.ld` -> `LDTM`, bulk async copies -> `UBLKCP`, "
      "`cp.async` -> `LDGSTS`, `st.async` -> `STAS`, mbarrier -> `SYNCS`).\n")
for want in WANT:
    hit = [m for m, d in names.items() if re.sub(r"\(.*", "", d).replace("void amc3d::", "").replace("void ", "").replace("(bool)", "").replace("(int)", "") .replace("amc3d::", "") == want]
    if not hit:
        hit = [m for m, d in names.items() if want.split("<")[0] in d and all(tok in d.replace("(int)", "").replace("(bool)", "") for tok in [want])]
    if not hit:
        print(f"## `{want}` — not found\n")
        continue
    ins = funcs[hit[0]]
    c = collections.Counter(i.split(".")[0] for i in ins)
    full = collections.Counter(ins)
    print(f"## `{want}` — {len(ins)} instructions\n")
    print("| group | count | mnemonics |")
    print("|---|---:|---|")
    used = set()
    for title, pat in GROUPS:
        sel = {k: v for k, v in c.items() if re.match(pat, k)}
        used |= set(sel)
        if sel:
            print(f"| {title} | {sum(sel.values())} | " + ", ".join(f"{k} {v}" for k, v in sorted(sel.items(), key=lambda kv: -kv[1])[:8]) + " |")
    rest = {k: v for k, v in c.items() if k not in used}
    if rest:
        print(f"| other | {sum(rest.values())} | " + ", ".join(f"{k} {v}" for k, v in sorted(rest.items(), key=lambda kv: -kv[1])[:8]) + " |")
    proof = [k for k in full if re.match(r"^(UTC|LDTM|UBLKCP|LDGSTS|STAS|UTMA|SYNCS|REDG|UCGABAR|REDUX)", k)]
    if proof:
        print("\nverbatim: " + ", ".join(f"`{k}` x{full[k]}" for k in sorted(proof)) + "\n")
    else:
        print()
