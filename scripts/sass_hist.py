"""SASS opcode histograms of the hot kernels in the built library (cuobjdump -sass; runs on the CPU box).
    python scripts/sass_hist.py > profiles/r2_sass_histograms.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "neural-volume-super-resolution_b200", "libnvsr_b200.so")
WANT = ["mlp_chain_tc_kernelILb1ELi4ELi1ELi0ELb0", "mlp_chain_tc_kernelILb1ELi4ELi3ELi1ELb0", "mlp_chain_tc_kernelILb1ELi4ELi3ELi1ELb1",
        "mlp_chain_tc_kernelILb1ELin6", "gather_tile_16ILb1ELi6", "gather_rows_16ILb1ELi6", "composite_kernelILb1ELb1",
        "composite_kernelILb0ELb0", "composite_bwd_warp_kernel", "dgrad_chain_kernel", "wgrad_kernel", "nonzero_rows_kernel",
        "ray_sum_rows_kernel", "gather_bwd_rows_kernel", "pack_weights_kernelI6__half", "sr_finalize_kernelILb1"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            funcs[cur][m.group(1)] += 1
    for w in WANT:
        for name, c in funcs.items():
            if w in name:
                dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
                total = sum(c.values())
                print(f"== {dem[:150]}\n   {total} SASS instructions")
                groups = collections.Counter()
                for op, n in c.items():
                    groups[op.split(".")[0]] += n
                print("   by mnemonic: " + ", ".join(f"{k} {v}" for k, v in groups.most_common(28)))
                keys = [k for k in c if re.match(r"UTC|LDTM|STTM|UBLKCP|UTMA|SYNCS|LDG\.E\.(ENL2\.)?256|LDG\.E\.128|STG\.E\.(EF\.)?128|F2FP|FFMA2|HADD2\.F32|RED|MUFU|DADD|DMUL|DFMA|SHFL", k)]
                print("   notable: " + ", ".join(f"{k} {c[k]}" for k in sorted(keys)))
                print()


if __name__ == "__main__":
    main()
