import sys,json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print(d["config"].get("name"), "n_gpus=%d"%d["n_gpus"], "edges/s=%.3g"%d["value"], "ms fwd/bwd=%.3f/%.3f"%(d.get("ms_fwd",0),d.get("ms_bwd",0)), "frac=%.3f"%d["roofline"]["frac"] if "roofline" in d else "", "e2e=%.3g"%d["e2e"]["value"])
