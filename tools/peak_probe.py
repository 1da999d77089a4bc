import sys; sys.path.insert(0, ".")
from pymc_statespace_b200 import fp64_peak_tflops
for w in (1, 2, 4, 8, 16):
    print(w, "warps/SMSP: reuse", round(fp64_peak_tflops(warps_per_smsp=w), 2), " distinct", round(fp64_peak_tflops(distinct_operands=True, warps_per_smsp=w), 2))
