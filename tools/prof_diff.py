"""Compare two per-op profiles written by `bench.py --profile-out` (fused ops are matched to the sum of their parts)."""
import json
import sys

a = json.load(open(sys.argv[1]))["ops"]
b = json.load(open(sys.argv[2]))["ops"]
kinds = set(sys.argv[3].split(",")) if len(sys.argv) > 3 else None
da = {o["name"]: o["ms"] for o in a}
print("total %.3f -> %.3f ms" % (sum(o["ms"] for o in a), sum(o["ms"] for o in b)))
for o in b:
    pre = o["name"].split("/")[0] + "/"
    names = [n if n.startswith(pre) else pre + n for n in o["name"].split("+")]
    was = sum(da.get(n, 0.0) for n in names)
    if kinds and o["kind"] not in kinds:
        continue
    if abs(was - o["ms"]) > 0.05 * max(was, 0.05):
        print("%-44s %8.3f -> %8.3f  %s" % (o["name"][:44], was, o["ms"],
                                            ("%.0f TF/s" % (o["flops"] / o["ms"] / 1e9)) if o["flops"] else ""))
