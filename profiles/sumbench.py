import json,sys
l=[x for x in open(sys.argv[1]).read().splitlines() if x.startswith("{")][-1]
d=json.loads(l)
print("N=%d value %.1f M step %.2f us frac %.4f frac_timed %s"%(d["n_gpus"],d["value"]/1e6,d["ms_per_step"]*1e3,d["roofline"]["frac"],d["roofline"].get("frac_over_timed_steps")), d["clocks"])
print(" e2e %.2f M frac %.3f ceil %.1f GB/s"%(d["e2e"]["value"]/1e6, d["e2e"].get("frac_of_d2h_ceiling") or 0, d["e2e"].get("d2h_ceiling_GBps") or 0))
print(" gather", d.get("gather"))
for k,v in (d.get("extra") or {}).items(): print(" ",k, "%.2f M"%(v["value"]/1e6), "ms %.4f"%v["ms_per_step"], v.get("verified"))
