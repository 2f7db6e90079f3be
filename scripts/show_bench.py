#!/usr/bin/env python
"""Prints the headline numbers of bench.py JSON lines: show_bench.py FILE..."""
import json
import sys


def show(f):
    line = [l for l in open(f) if l.startswith("{")][-1]
    j = json.loads(line)
    r = lambda d: {k: round(v, 2) for k, v in d.items()}
    print("==", f, "N=%d value %.3e ms %.2f" % (j["n_gpus"], j["value"], j["ms_per_step"]), r(j["stages_ms"]), "exchange:", j["config"].get("exchange"))
    s = j.get("strong")
    if s:
        print("  strong value %.3e ms %.2f" % (s["value"], s["ms_per_step"]), r(s["stages_ms"]))
    if j.get("parity"):
        print("  parity", {k: v for k, v in j["parity"].items() if not isinstance(v, dict) and "cpu" not in k and "oracle" not in k})
    if j.get("e2e"):
        e = j["e2e"]
        print("  e2e %.3e ms %.1f" % (e["value"], e["ms_per_step"]), r(e["calls_ms"]), "ceiling GB/s/GPU %.1f frac %.2f" % (e["h2d_ceiling"]["gb_per_s_per_gpu_all_ranks_concurrent"], e["h2d_ceiling"]["gridding_call_frac_of_ceiling"]))
    c = j.get("config5")
    if c:
        print("  c5 %.3e ms %.1f" % (c["value"], c["ms_per_step"]), r(c["stages_ms"]), "routing share %.3f" % c["routing_share_of_step"],
              "checksum %.1e adjoint %.1e" % (c["parity"]["checksum_rel_err"], c["parity"]["adjoint_rel_err"]), "routed", c["config"]["routed_records"])
    if j.get("aw"):
        for k, v in j["aw"].items():
            if isinstance(v, dict):
                print("  aw", k, "device ms %.2f e2e vis/s %.3e pinned %.3e" % (v["device_ms"], v["vis_per_s_e2e"], v["vis_per_s_e2e_pinned"]), {a: b for a, b in v.items() if "parity" in a})
    print("  clocks", j["clocks"], "roofline l2 frac %.3f fp64 frac %.3f kernel ms %.2f" % (j["roofline"]["l2_taps"]["frac"], j["roofline"]["fp64"]["frac"], j["roofline"]["kernel_ms"]))


for f in sys.argv[1:]:
    show(f)
