"""SASS evidence (cuobjdump of the built objects, no GPU needed): per kernel, the counts of the Blackwell-native opcodes
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
LDGSTS = cp.async, HMMA = mma.sync, MUFU, FFMA2 ...) and the instruction total.
    python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJS = ["gemm_tcgen05.o", "scan_fwd_bf16.o", "scan_fwd_wp_bf16.o", "scan_fwd_wp2_bf16.o", "scan_fwd_wph_bf16.o", "conv1d.o", "norm.o"]
WANT = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "LDGSTS", "ARRIVES", "HMMA", "LDSM", "MUFU", "FFMA2", "FMUL2", "FADD2", "BAR"]
KEEP = re.compile(r"gemm_bf16_tn_kernel|scan_fwd_tma_kernel|scan_fwd_tpc2_kernel|scan_fwd_wp_kernel|scan_fwd_wp2_kernel|scan_fwd_wph_kernel|conv_fwd_tok4_kernel|block_tail_kernel|block_tail_row4_kernel")
for o in OBJS:
    path = os.path.join(ROOT, "build", "obj", o)
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, counts, total = None, None, 0
    out = []
    def flush():
        if name and KEEP.search(name):
            dm = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            out.append((dm, total, dict(counts)))
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            name, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            total += 1
            op = m.group(1)
            for w in WANT:
                if op.startswith(w):
                    counts[w] += 1
    flush()
    print(f"== {o}")
    for dm, total, c in out:
        print(f"  {dm[:150]}")
        print(f"      {total} instructions: " + ", ".join(f"{k} {c[k]}" for k in WANT if c.get(k)))
