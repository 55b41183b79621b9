import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f, "value=%.4e ms_per_step=%.3f kernel_ms=%s digest_ms=%s" % (d["value"], d["ms_per_step"], r.get("kernel_ms"), r.get("digest_kernel_ms")),
              d["config"].get("collective"), d.get("e2e", {}) and d["e2e"].get("value"))
    except Exception as e:
        print(f, "FAILED", e)
