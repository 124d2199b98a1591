"""Hot SASS regions of one kernel in an .ncu-rep: instructions executed and stall
samples, bucketed in runs of N SASS lines.  python tools/ncu_hot.py rep regex [N]"""
import csv
import io
import subprocess
import sys


def main(rep, pattern, n=40, top=14):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name",
                          "regex:" + pattern], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # first block only (first matching launch)
    hdr = None
    body = []
    for r in rows:
        if r and r[0] == "Address":
            if hdr is not None:
                break
            hdr = r
            continue
        if hdr is not None and r and r[0].startswith("0x"):
            body.append(r)
    ie = hdr.index("Instructions Executed")
    ss = hdr.index("# Samples")
    tot_i = sum(int(r[ie]) for r in body)
    tot_s = sum(int(r[ss]) for r in body)
    print("total warp instructions %d, samples %d, SASS lines %d" % (tot_i, tot_s, len(body)))
    buckets = []
    for b in range(0, len(body), n):
        chunk = body[b:b + n]
        buckets.append((sum(int(r[ie]) for r in chunk), sum(int(r[ss]) for r in chunk), b))
    buckets.sort(reverse=True)
    for inst, samp, b in buckets[:top]:
        print("-- lines %d..%d: %.1f%% of instructions, %.1f%% of samples" %
              (b, b + n, 100.0 * inst / tot_i, 100.0 * samp / max(tot_s, 1)))
        ops = {}
        for r in body[b:b + n]:
            op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
            ops[op] = ops.get(op, 0) + int(r[ie])
        print("   " + ", ".join("%s %.1fM" % (k, v / 1e6) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:12]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
