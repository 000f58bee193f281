"""Scratch: from an ncu launch-list CSV pick the launch-skip count that lands `ncu -k regex:<pat> -s N` on the hot attempt
kernel with the largest grid (a full-size round): python tools/pick_skip.py launches.csv 'k_attempt_hot|k_commit_coop|k_spheres|k_attempt_slow'"""
import csv, re, sys
pat = re.compile(sys.argv[2])
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
best, bestg, k = 0, -1, 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    if not pat.search(name):
        continue
    if "k_attempt_hot" in name:
        g = int(row["Grid Size"].strip("()").split(",")[0])
        if g > bestg:
            best, bestg = k, g
    k += 1
print(best)
