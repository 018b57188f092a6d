import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for l in sys.stdin:
    try:
        d = json.loads(l)
    except Exception:
        if l.strip(): print(l.strip()[:300])
        continue
    r = d.get("roofline", {})
    o = d.get("other_mode") or {}
    print(tag, d["config"]["workload"].split(",")[-1].strip(), "Gcell/s %.3f ms/step %.2f pred %.2f corr %.2f frac %.3f e2e %.3f | other(%s) %.3f G %.2f ms pred %.2f corr %.2f" % (
        d["value"] / 1e9, d["ms_per_step"], r.get("predictor_ms_per_step", 0), r.get("corrector_ms_per_step", 0), r.get("frac") or 0,
        ((d.get("e2e") or {}).get("value") or 0) / 1e9, o.get("mode"), (o.get("value") or 0) / 1e9, o.get("ms_per_step") or 0,
        o.get("predictor_ms_per_step") or 0, o.get("corrector_ms_per_step") or 0))
