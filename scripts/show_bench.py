import sys, json
for l in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench.json"):
    if l.startswith('{"metric"'):
        d = json.loads(l)
        o = d.get("other_workloads", {})
        print("cfg3", round(d["value"]), "fps; ms/step", {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, "frac", round(d["roofline"]["frac"], 3),
              "e2e", round(d["e2e"]["value"]), "e2e_rgb8", round(d.get("e2e_rgb8_sink", {}).get("value", 0)), "| cfg2", round(o.get("cfg2", {}).get("frames_per_s", 0)), "cfg5", round(o.get("cfg5", {}).get("frames_per_s", 0)), "| clocks", d.get("clocks"))
