"""Side-by-side table of per-launch CUDA-event timings written by `bench.py --dump-ops`."""
import json
import sys


def main():
    tabs = [json.load(open(p)) for p in sys.argv[1:]]
    names = [o["op"] for o in tabs[0]]
    maps = [{o["op"]: o for o in t} for t in tabs]
    tot = [0.0] * len(tabs)
    for n in names:
        row = [m.get(n, {}).get("ms", float("nan")) for m in maps]
        for i, v in enumerate(row):
            tot[i] += v if v == v else 0.0
        print("%-40s" % n + "".join(" %8.4f" % v for v in row) +
              ("  %+6.1f%%" % (100.0 * (row[-1] / row[0] - 1.0)) if len(row) > 1 and row[0] > 0 else ""))
    print("%-40s" % "TOTAL" + "".join(" %8.4f" % v for v in tot))


if __name__ == "__main__":
    main()
