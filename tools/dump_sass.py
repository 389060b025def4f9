#!/usr/bin/env python
"""SASS listings of the production kernels for profiles/ (the north star asks for a committed listing per kernel).

    python tools/dump_sass.py [TAG]      ->  profiles/TAG_sass/<kernel>.sass  +  profiles/TAG_sass/README.md

Each file starts with the kernel's instruction mix (opcode counts), so the claims of DESIGN.md can be checked
without reading the listing: FFMA2 = packed FP32 FMA, UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld,
LDGSTS = cp.async, UTCBAR = tcgen05.commit, SYNCS = mbarrier, REDUX / IDP = warp reduction / dp4a."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fm_radio_b200", "libfmgpu.so")
WANT = [  # (file stem, regex on the demangled name)
    ("k1_toeplitz_i8", r"k1_toeplitz_i8<\(int\)2, \(int\)1>"), ("k1_fir4_discrim_u8", r"k1_fir4_discrim_u8"),
    ("k1_fir4_discrim_cf32", r"k1_fir4_discrim_cf32"),
    ("k2_mpx_sparse_hilbert", r"k2_mpx<\(bool\)1>"), ("k3_pll_fast", r"k3_pll<\(bool\)0, \(int\)0, \(bool\)1>"),
    ("k3_pll_exact", r"k3_pll<\(bool\)0, \(int\)0, \(bool\)0>"), ("k4_mix_fir", r"k4_mix_fir"), ("k4b_lmr_phase", r"k4b_lmr_phase"),
    ("k5_bpsk_symbolwise", r"k5_bpsk<\(bool\)0, \(bool\)0>"), ("k6_rds", r"k6_rds\("), ("k7_audio_pcm", r"k7_audio_pcm"),
    ("chan_mma_i8_pipelined", r"chan_mma_i8_pipelined"), ("chan_mma_i8_v1", r"chan_mma_i8\("), ("chan_fir_fp32", r"chan_fir_fp32"),
]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    out_dir = os.path.join(ROOT, "profiles", f"{tag}_sass")
    os.makedirs(out_dir, exist_ok=True)
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    chunks = re.split(r"\n\s*Function : ", sass)[1:]
    names = subprocess.run(["cu++filt"] + [c.split("\n", 1)[0].strip() for c in chunks], capture_output=True, text=True).stdout.splitlines()
    rows = []
    for stem, pat in WANT:
        hit = [(n, c) for n, c in zip(names, chunks) if re.search(pat, n)]
        if not hit:
            print("missing:", stem)
            continue
        name, body = hit[0]
        ops = collections.Counter()
        lines = []
        for ln in body.split("\n")[1:]:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", ln)
            if m:
                txt = m.group(2).strip()
                op = re.sub(r"^@!?U?P\d+\s+", "", txt).split()[0].split(".")[0]
                ops[op] += 1
                lines.append(f"/*{m.group(1)}*/  {txt}")
        mix = ", ".join(f"{k} {v}" for k, v in ops.most_common(18))
        with open(os.path.join(out_dir, stem + ".sass"), "w") as f:
            f.write(f"// {name}\n// {sum(ops.values())} instructions; mix: {mix}\n")
            f.write("\n".join(lines) + "\n")
        rows.append((stem, sum(ops.values()), ops))
    keys = ["FFMA2", "FFMA", "FADD2", "FMUL2", "LDS", "STS", "LDG", "STG", "LDGSTS", "UTCIMMA", "LDTM", "UTCBAR", "SYNCS", "MUFU", "IDP", "REDUX", "BAR", "SHFL"]
    with open(os.path.join(out_dir, "README.md"), "w") as f:
        f.write(f"# SASS listings ({tag}), `cuobjdump -sass fm_radio_b200/libfmgpu.so`, one file per production kernel\n\n")
        f.write("Static opcode counts (whole kernel, all paths):\n\n| kernel | instr | " + " | ".join(keys) + " |\n|---|---|" + "---|" * len(keys) + "\n")
        for stem, n, ops in rows:
            f.write(f"| {stem} | {n} | " + " | ".join(str(ops.get(k, 0)) for k in keys) + " |\n")
    print(open(os.path.join(out_dir, "README.md")).read())


if __name__ == "__main__":
    main()
