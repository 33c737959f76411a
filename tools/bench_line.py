"""print the key fields of a bench.py JSON line read from stdin (experiment helper)"""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(tag, "value", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), {k: round(v, 2) for k, v in d["breakdown_ms_per_step"].items()},
      "coef", round(d["coefficient_path"]["value"], 2), "frac", round(d["roofline"]["frac"], 3))
