"""Reads one bench.py JSON line on stdin and prints the numbers one compares between two runs."""
import json
import sys

line = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = line["roofline"]
print(sys.argv[1] if len(sys.argv) > 1 else "", "value", round(line["value"], 1), "e2e", round(line["e2e"]["value"], 1),
      "ms/step", round(line["ms_per_step"], 1), "profiled ms", round(r.get("profiled_ms_per_step", 0.0), 1),
      {k: v["stream_ms_per_step"] for k, v in r.get("families", {}).items()})
