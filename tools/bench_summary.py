import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N=%d value %.4e ms %.4f e2e %.4e ms %.4f"%(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
for k in [k for k in d if k.startswith("roofline")]:
    r=d[k]; print(k, r["kernel"][:28], "frac %.3f ms %.4f moved_frac %.3f achieved %.0f traffic %s"%(r["frac"], r["ms_per_launch"], r["moved_frac"], r["achieved"], r["traffic"]))
print("c5 %.4e e2e %.4e ms %.3f"%(d["c5"]["value"], d["c5"]["e2e"]["value"], d["c5"]["ms_per_step"]))
if "cpu_baseline" in d: print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "ratio e2e/cpu %.0f" % (d["e2e"]["value"] / d["cpu_baseline"]["value"]))
print(d["clocks"], d["gpu_launches"], d["fp64"]["algorithmic_frac"])
