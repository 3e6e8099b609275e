"""Print one line per bench JSON file (helper for profiles/ notes)."""
import json
import sys

for f in sys.argv[1:]:
    txt = open(f).read().strip().splitlines()
    try:
        d = json.loads(txt[-1])
        r = d["roofline"]
        print("%-34s %.3e agent-steps/s  %.2f us/launch  %.0f GB/s  frac %.3f  e2e %.3e  sm %s MHz %s" % (
            f.split("/")[-1], d["value"], r["avg_launch_us"], r["achieved"], r["frac"], d["e2e"]["value"],
            d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e, txt[-2:])
