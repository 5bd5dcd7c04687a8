import sys, json
for l in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench.json"):
    if l.startswith('{"metric"'):
        d = json.loads(l)
        o = d.get("other_workloads", {})
        print(d["config"]["workload"][:4], round(d["value"]), "fps; ms/step", {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, "frac", round(d["roofline"]["frac"], 3),
              "e2e", round(d.get("e2e", {}).get("value", 0)), {k: round(v["value"]) for k, v in d.get("e2e_variants", {}).items()},
              "|", {k: (round(v["value"]), round(v["roofline"]["frac"], 3)) for k, v in o.items()}, "| parity_ranks", d.get("parity_ranks"), "| clocks", d.get("clocks"))
