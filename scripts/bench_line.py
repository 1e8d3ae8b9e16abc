#!/usr/bin/env python
"""Print the headline numbers of a bench.py JSON line read from stdin (label = argv[1])."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(sys.argv[1] if len(sys.argv) > 1 else "", "fps %.0f  e2e %.0f  ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
