import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("cfg2", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["parity_on_sample"], "frac", d["roofline"]["frac"], d["roofline"].get("probe_frac"))
if "locate" in d:
    l=d["locate"]; print("locate", l["value"], l["ms_per_step"], "e2e", l["e2e"]["value"], l["e2e"]["ms_per_step"], l.get("cpu_baseline",{}).get("parity_on_sample"))
    if "find" in l: print("  find 64-mers", l["find"]["ms_per_step"], l["find"]["value"])
    for w in l.get("wide_ranges",[]): print("  wide", w["pattern_length"], w["ranges"], round(w["path_nodes_per_range"],1), w["ms_per_step"], w["value"], w.get("cpu_baseline",{}).get("parity_on_sample"))
if "cfg5" in d:
    c=d["cfg5"]; print("cfg5", c["value"], c["ms_per_step"], c.get("cpu_baseline",{}).get("parity_on_sample"), c.get("roofline",{}).get("frac"), c.get("roofline",{}).get("dram_frac"))
if "cfg4" in d:
    c=d["cfg4"]; print("cfg4", c["value"], c["ms_per_pass"], c.get("cpu_baseline",{}).get("parity_on_sample"), c["config"]["index"].get("jump_k"), c["config"]["index"]["device_bytes"], c.get("roofline",{}).get("probe_frac"))
print("clocks", d["clocks"], "launches", d["gpu_launches"], "n_gpus", d["n_gpus"])
