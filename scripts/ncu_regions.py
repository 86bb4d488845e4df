"""Developer tool: stall samples of one ncu capture aggregated over buckets of consecutive SASS instructions, with the dominant stall
reasons and opcodes of each bucket (where in the kernel the time goes).    python scripts/ncu_regions.py file.ncu-rep [bucket]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, bucket = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 200
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[start - 1][:2])
    hdr, data = rows[start], [r for r in rows[start + 1:] if len(r) > 10]
    ix = {h: i for i, h in enumerate(hdr)}
    S, src, ex = ix["# Samples"], ix["Source"], ix["Instructions Executed"]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[S] or 0) for r in data)
    print("total samples", tot, "instructions", len(data))
    for b in range(0, len(data), bucket):
        chunk = data[b:b + bucket]
        s = sum(int(r[S] or 0) for r in chunk)
        execd = sum(int(r[ex] or 0) for r in chunk)
        agg = collections.Counter()
        for r in chunk:
            for h in stalls:
                if r[ix[h]]:
                    agg[h] += int(r[ix[h]])
        ops = collections.Counter((r[src].split()[1] if r[src].startswith("@") else r[src].split()[0]) for r in chunk if r[src].split())
        top = ", ".join(f"{k[6:]}:{v}" for k, v in agg.most_common(4))
        print(f"{b:5d} {100 * s / max(tot, 1):5.1f}% exec {execd / 1e6:7.1f}M  {top} | {ops.most_common(4)}")


if __name__ == "__main__":
    main()
