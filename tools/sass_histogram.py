"""dev tool: per-kernel SASS opcode histograms (evidence for profiles/): which memory / math instructions a kernel uses.

    python tools/sass_histogram.py > profiles/r2_sass_opcodes.md

Reads the objects under build/csrc (cuobjdump -sass), demangles the kernel names and prints, for the kernels listed in
KERNELS, the instruction count and the opcodes that prove the data path: UBLKCP / SYNCS (TMA bulk copies + mbarriers),
LDGSTS (cp.async), DFMA / DMUL / DADD (fp64 pipe), DMMA (FP64 tensor cores), LDS / STS, LDG / STG, SHFL, BAR, MUFU."""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = [  # (object glob, demangled-name regex, label)
    ("kf_p1.o", r"kf_p1_forward_kernel<2, true, true, true>", "headline forward: k_endog = 1, k_states = 2, loglik + tape (Z = e0, H = 0 promised)"),
    ("kf_p1.o", r"kf_p1_adjoint_kernel<2, false, false, false, true, true>", "headline adjoint: TMA tape ring (Z = e0, H = 0 promised)"),
    ("kf_thread_m2.o", r"kf_thread_kernel<2, 1, 0, 0, false>", "generic thread-per-unit forward (A/B reference)"),
    ("kf_thread_m2.o", r"kf_thread_kernel<2, 1, 0, 2, false>", "generic thread-per-unit adjoint (cp.async ring)"),
    ("kf_thread_m2.o", r"kf_thread_kernel<2, 1, 0, 1, false>", "full-output forward (filtered / predicted moments)"),
    ("kf_coopT_m6.o", r"kf_rows_kernel<6, 3, 4, 0, false, false>", "config 3 forward: fused rows, k_states = 6"),
    ("kf_coopT_m6.o", r"kf_rows_kernel<6, 3, 4, 0, true, false>", "config 3 adjoint: fused rows"),
    ("kf_coopT_m30.o", r"kf_rowsD_kernel<30, 1, 0, false, false>", "config 4 forward: DMMA tile products"),
    ("kf_coopT_m30.o", r"kf_rowsD_kernel<30, 1, 0, true, false>", "config 4 adjoint: DMMA tile products"),
    ("kf_coopT_m30.o", r"kf_rowsD_kernel<30, 1, 0, true, true>", "config 4 adjoint with T-bar: T-bar accumulated on the tensor cores"),
    ("kf_coopT_m30.o", r"kf_dareD_kernel<30, 1>", "steady-state covariance (DARE) on the tensor cores, k_states = 30"),
    ("kf_coopT_m14.o", r"kf_rowsD_kernel<14, 1, 0, false, false>", "k_states 14 forward: 16 x 16 tiles"),
    ("kf_coopT_m14.o", r"kf_rowsD_kernel<14, 1, 0, true, false>", "k_states 14 adjoint: 16 x 16 tiles"),
]
SHOW = ("UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DMUL", "DADD", "DSETP", "DMMA", "MUFU", "LDS", "STS", "LDG", "LD", "STG",
        "SHFL", "BAR", "WARPSYNC", "BRA", "FSEL", "IMAD", "IADD3")


def kernels_of(obj):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    out, name = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and name:
            out[name][m.group(1)] += 1
    return out


def main():
    print("# SASS opcode histograms of the kernels on BASELINE.json's configs (round 2)\n")
    print("`python tools/sass_histogram.py` over `build/csrc/*.o` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a`).")
    print("Whole-kernel static counts (prologue + loop bodies + epilogue), opcode suffixes folded.  tcgen05 has no FP64")
    print("kind, so no `UTC*MMA` / `LDTM` is expected on this path; `UBLKCP` + `SYNCS` are the TMA bulk copies and their")
    print("mbarriers, `LDGSTS` is `cp.async`, `DMMA` the FP64 tensor-core instruction.\n")
    cache = {}
    for objname, rx, label in KERNELS:
        obj = os.path.join(ROOT, "build", "csrc", objname)
        if obj not in cache:
            ks = kernels_of(obj)
            names = list(ks)
            dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
            cache[obj] = list(zip(dem, (ks[n] for n in names)))
        hit = [(d, c) for d, c in cache[obj] if re.search(rx, d)]
        if not hit:
            print(f"## {label}\n\n(not found: {rx})\n")
            continue
        d, c = hit[0]
        total = sum(c.values())
        print(f"## {label}\n\n`{d.split('(')[0]}` - {total} instructions\n")
        print("| opcode | count |\n|---|---|")
        for op in SHOW:
            if c.get(op):
                print(f"| {op} | {c[op]} |")
        rest = sum(v for k, v in c.items() if k not in SHOW)
        print(f"| (other) | {rest} |\n")


if __name__ == "__main__":
    main()
